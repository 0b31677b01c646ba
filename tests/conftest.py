import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (B200, sm_100a) device")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


def same_checksum(a: float, b: float) -> bool:
    """Weight fingerprints are float64 sums; the reduction order of torch.sum differs between CPUs (SIMD width)."""
    return abs(a - b) <= 1e-9 * max(abs(a), abs(b), 1.0)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| relative to the largest magnitude of the reference tensor b."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def amt_state_dict(vf_dim, seed, chord_embed=False, wout_gain=1.0, n_layers=6, d_model=512, d_ff=1024, er_len=300,
                   max_video=300):
    """The seeded weights every AMT golden was generated with (same shapes as the reference constructor)."""
    from video2music_b200 import synthetic as syn
    from video2music_b200 import VideoMusicTransformer
    m = VideoMusicTransformer(n_layers=n_layers, d_model=d_model, dim_feedforward=d_ff, total_vf_dim=vf_dim, rpr=True,
                              chord_embed=chord_embed, max_sequence_chord=er_len, max_sequence_video=max_video)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = syn.fill_like_reference_init(shapes, seed=seed, wout_gain=wout_gain)
    # (chord_embedding_model.weight, when present, is seeded like every other table: make_golden.py overwrites the
    #  stand-in Word2Vec vectors of oracle/ref_shim.py through load_state_dict)
    m.load_state_dict(sd, strict=False)
    return m, {k: v.clone() for k, v in m.state_dict().items()}
