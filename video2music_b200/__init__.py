"""video2music_b200 -- B200-native (sm_100a) kernels behind the PyTorch module API of the Affective
Multimodal Transformer hot path of khangklj/Video2Music.

Importing the package never touches the GPU; the compute modules load csrc/libv2m_b200.so on first
use and fail loudly if it is missing (no CPU or PyTorch fallback).
"""
from .video_music_transformer import VideoMusicTransformer  # noqa: F401
from .rpr import MultiheadAttentionRPR, TransformerDecoderLayerRPR, TransformerDecoderRPR  # noqa: F401
from .pscan import pscan  # noqa: F401
from .moe import GLUExpert, MoELayer, SharedMoELayer, TopKScheduler, TemperatureScheduler  # noqa: F401
from .grouped_query_attention import MultiheadGQA, scaled_dot_product_gqa  # noqa: F401
from .custom_transformer import (TransformerEncoderLayer, TransformerDecoderLayer, TransformerEncoder,  # noqa: F401
                                 TransformerDecoder, TransformerEncoderShorter, TransformerDecoderShorter,
                                 CustomMultiheadAttention, DifferentialMultiheadAttention, RotaryPositionalEmbeddings)
from .video_music_transformer_v2 import VideoMusicTransformer_V1, VideoMusicTransformer_V2, VideoMusicTransformer_V3  # noqa: F401
from .mamba import MambaConfig, MambaBlock, ResidualBlock, Mamba, RMSNorm, BiMambaEncoderLayer, BiMambaEncoderLayer_V1, BiMambaEncoder  # noqa: F401

from .video_regression import VideoRegression  # noqa: F401

__version__ = "0.1.0"
