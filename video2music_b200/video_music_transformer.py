"""Drop-in for model/video_music_transformer.py:910-1132 of the reference (base AMT class).

Same constructor signature, same parameter / buffer names and shapes (a reference checkpoint loads
with `load_state_dict`, train.py:180 / video2music.py:649), same `forward` and `generate`
signatures and tensor layouts.  All arithmetic runs in the sm_100a kernels of libv2m_b200.so via
`engine.py`; there is no PyTorch fallback path.

Extensions over the reference (additive, the reference call patterns keep their meaning):
  * `generate` accepts a batch of videos (features with a leading batch dim B >= 1, primer (P,) or
    (B, P)) and decodes them together with a KV cache; the reference is hard-wired to batch 1 and
    re-runs the full model every step (:1059-1071).
  * `compute_dtype` (torch.float32 | torch.bfloat16) selects the exact SIMT path or the tensor-core path.
"""
import copy
from typing import Optional

import torch
import torch.nn as nn

from . import engine
from .positional_encoding import PositionalEncoding
from .rpr import MultiheadAttentionRPR, TransformerDecoderLayerRPR, TransformerDecoderRPR, _get_clones

# utilities/constants.py:50-62,66
CHORD_END, CHORD_PAD, CHORD_SIZE = 157, 158, 159
CHORD_ROOT_SIZE, CHORD_ATTR_SIZE = 15, 16
SCENE_OFFSET_MAX = 300
IS_SEPERATED = False                    # utilities/constants.py:11


class _EncoderLayer(nn.Module):
    """Parameter container with the key names of the stock nn.TransformerEncoderLayer that
    nn.Transformer builds at video_music_transformer.py:967-971 (post-norm, ReLU)."""

    def __init__(self, d_model, nhead, dim_feedforward, dropout):
        super().__init__()
        self.self_attn = MultiheadAttentionRPR(d_model, nhead, dropout=dropout, er_len=None)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)


class _Encoder(nn.Module):
    def __init__(self, layer, num_layers, norm):
        super().__init__()
        self.layers = _get_clones(layer, num_layers)
        self.num_layers = num_layers
        self.norm = norm


class AMTTransformer(nn.Module):
    """Shape of torch.nn.Transformer(d_model, nhead, 6, 6, dim_feedforward, dropout, custom_decoder=...):
    `.encoder.layers[i]`, `.encoder.norm`, `.decoder` and `generate_square_subsequent_mask`."""

    def __init__(self, d_model, nhead, num_encoder_layers, num_decoder_layers, dim_feedforward, dropout,
                 custom_decoder=None, er_len=None):
        super().__init__()
        self.d_model, self.nhead = d_model, nhead
        self.encoder = _Encoder(_EncoderLayer(d_model, nhead, dim_feedforward, dropout), num_encoder_layers,
                                nn.LayerNorm(d_model))
        if custom_decoder is not None:
            self.decoder = custom_decoder
        else:
            self.decoder = TransformerDecoderRPR(
                TransformerDecoderLayerRPR(d_model, nhead, dim_feedforward, dropout, er_len=er_len), num_decoder_layers,
                nn.LayerNorm(d_model))
        self._reset_parameters()

    def _reset_parameters(self):
        """nn.Transformer._reset_parameters: xavier-uniform on every parameter with dim > 1 (this
        includes Er)."""
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    @staticmethod
    def generate_square_subsequent_mask(sz, device=None, dtype=None):
        return torch.triu(torch.full((sz, sz), float("-inf"), device=device, dtype=dtype or torch.float32), diagonal=1)


class VideoMusicTransformer(nn.Module):
    def __init__(self, n_layers=6, num_heads=8, d_model=512, dim_feedforward=1024,
                 dropout=0.1, max_sequence_midi=2048, max_sequence_video=300,
                 max_sequence_chord=300, total_vf_dim=0, rpr=False, scene_embed=False,
                 chord_embed=False, chord_embedding_weights: Optional[torch.Tensor] = None):
        super().__init__()
        self.nlayers = n_layers
        self.nhead = num_heads
        self.d_model = d_model
        self.d_ff = dim_feedforward
        self.dropout = dropout
        self.max_seq_midi = max_sequence_midi
        self.max_seq_video = max_sequence_video
        self.max_seq_chord = max_sequence_chord
        self.rpr = rpr
        self.scene_embed = scene_embed
        self.chord_embed = chord_embed
        if scene_embed:
            raise NotImplementedError("scene_embed=True is not used by any shipped configuration (train.py:136-168)")
        if chord_embed:
            # The reference loads word2vec_filled.bin through gensim (video_music_transformer.py:933-937); gensim is
            # not a dependency here, so the (159, 512) table is passed in or arrives through load_state_dict.
            w = chord_embedding_weights if chord_embedding_weights is not None else torch.zeros(CHORD_SIZE, d_model)
            self.chord_embedding_model = nn.Embedding.from_pretrained(w.clone().float(), freeze=True)
        self.embedding = nn.Embedding(CHORD_SIZE, d_model)
        self.embedding_root = nn.Embedding(CHORD_ROOT_SIZE, d_model)
        self.embedding_attr = nn.Embedding(CHORD_ATTR_SIZE, d_model)
        self.total_vf_dim = total_vf_dim
        self.Linear_vis = nn.Linear(total_vf_dim, d_model)
        self.Linear_chord = nn.Linear(d_model + 1, d_model)
        self.positional_encoding = PositionalEncoding(d_model, dropout, max_sequence_chord)
        self.positional_encoding_video = PositionalEncoding(d_model, dropout, max_sequence_video)
        self.condition_linear = nn.Linear(1, d_model)
        self.transformer = AMTTransformer(d_model, num_heads, n_layers, n_layers, dim_feedforward, dropout,
                                          er_len=max_sequence_chord if rpr else None)
        self.Wout_root = nn.Linear(d_model, CHORD_ROOT_SIZE)
        self.Wout_attr = nn.Linear(d_model, CHORD_ATTR_SIZE)
        self.Wout = nn.Linear(d_model, CHORD_SIZE)
        self.softmax = nn.Softmax(dim=-1)
        self.compute_dtype = torch.float32
        self._weights = {}

    # ------------------------------------------------------------------ plumbing
    def __deepcopy__(self, memo):
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = {} if k == "_weights" else copy.deepcopy(v, memo)
        return new

    def set_compute_dtype(self, dtype: torch.dtype) -> "VideoMusicTransformer":
        assert dtype in (torch.float32, torch.bfloat16)
        self.compute_dtype = dtype
        return self

    def _cfg(self):
        return dict(d_model=self.d_model, nhead=self.nhead, d_ff=self.d_ff, n_layers=self.nlayers,
                    chord_embed=self.chord_embed)

    def _w(self) -> engine.AMTWeights:
        w = self._weights.get(self.compute_dtype)
        if w is None:
            w = self._weights[self.compute_dtype] = engine.AMTWeights(self, self.compute_dtype)
        return w

    def _device(self):
        return self.Wout.weight.device

    # ------------------------------------------------------------------ forward
    def forward(self, x, x_root, x_attr, feature_semantic_list, feature_key, feature_scene_offset, feature_motion,
                feature_emotion, mask=True):
        drop = float(self.dropout) if self.training else 0.0
        grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        dev = self._device()
        args = [t.to(dev) for t in (x, x_root, x_attr, feature_semantic_list, feature_key, feature_scene_offset,
                                    feature_motion, feature_emotion)]
        if grad or drop > 0:                  # training mode with dropout draws masks with or without gradients
            from .autograd import amt_forward_autograd
            seed = 0
            if drop > 0:                      # a fresh mask per forward call, reproducible under torch.manual_seed
                self._drop_calls = getattr(self, "_drop_calls", 0) + 1
                rank = torch.distributed.get_rank() if torch.distributed.is_available() and torch.distributed.is_initialized() else 0
                seed = (torch.initial_seed() * 1000003 + self._drop_calls * 7919 + rank * 104729) & 0xFFFFFFFF
            y = amt_forward_autograd(self, *args, mask=mask is True, dropout_p=drop, seed=seed)
        else:
            y = engine.amt_forward(self._w(), self._cfg(), *args, mask=mask is True)
        if IS_SEPERATED:
            raise NotImplementedError("IS_SEPERATED=True is a compile-time switch that is off in the reference")
        return y

    # ------------------------------------------------------------------ generation
    @torch.no_grad()
    def generate(self, feature_semantic_list=[], feature_key=None, feature_scene_offset=None, feature_motion=None,
                 feature_emotion=None, primer=None, primer_root=None, primer_attr=None, target_seq_length=300, beam=0,
                 beam_chance=1.0, max_conseq_N=0, max_conseq_chord=2, use_graph=True, return_logits=False, decode_mode="auto",
                 uniforms=None):
        assert (not self.training), "Cannot generate while in training mode"
        # beam >= 1 with beam_chance >= 1: the deterministic branch (:1078-1084) == greedy arg-max for beam == 1;
        # beam == 0: the sampling branch with its constraints (:1085-1128).  The random mix of both (0 < beam_chance < 1)
        # and beam > 1 (which only rewrites gen_seq rows of a batch of one) are not reproduced.
        if beam == 0:
            sample = True
        elif beam == 1 and beam_chance >= 1.0:
            sample = False
        else:
            raise NotImplementedError("generate(): beam=0 (sampling branch) and beam=1 with beam_chance=1.0 (greedy) run on "
                                      "device; beam > 1 / a random mix of branches (video_music_transformer.py:1073-1084) do not")
        dev = self._device()
        sem = feature_semantic_list.to(dev).float()
        B = sem.shape[0]
        key = torch.as_tensor(feature_key, dtype=torch.float32).reshape(-1)
        if key.numel() == 1 and B > 1:
            key = key.expand(B)
        st = engine.build_decode(self._w(), self._cfg(), sem, key.to(dev), feature_scene_offset.to(dev).float(),
                                 feature_motion.to(dev).float(), feature_emotion.to(dev).float(),
                                 primer.long(), primer_root.long(), primer_attr.long(), target_seq_length,
                                 want_logits=return_logits, mode=decode_mode, sample=sample, max_conseq_N=max_conseq_N,
                                 max_conseq_chord=max_conseq_chord, uniforms=uniforms)
        engine.run_decode(st, target_seq_length - 1, use_graph=use_graph)
        gen = st.gen[:, :target_seq_length]
        if return_logits:
            return gen, st.logits_all
        return gen
