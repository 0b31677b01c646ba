"""Deterministic synthetic inputs and weights.

There is no dataset and no checkpoint in the reference repository
(saved_models/AMT/README.md:1), so tests, goldens and benchmarks all use the
seeded generators below.  Only `torch.rand` / `torch.randint` on the CPU
generator are used (a plain mt19937 stream) so that the GPU box regenerates
bit-identical tensors from the seeds stored in tests/golden/*.
"""
import zlib
from typing import Dict, Tuple

import torch

# vocabulary constants, utilities/constants.py:50-62 of the reference
CHORD_END, CHORD_PAD, CHORD_SIZE = 157, 158, 159
CHORD_ROOT_PAD, CHORD_ROOT_SIZE = 14, 15
CHORD_ATTR_PAD, CHORD_ATTR_SIZE = 15, 16

SQRT12 = 12.0 ** 0.5


def _gen(seed: int, name: str = "") -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
    return g


def unit_uniform(shape, g) -> torch.Tensor:
    """zero mean, unit variance uniform."""
    return (torch.rand(shape, generator=g, dtype=torch.float32) - 0.5) * SQRT12


def vf_dim(motion_type: int = 0) -> int:
    """total_vf_dim as computed by train.py:110-130 (768 semantic + 1 scene offset
    + {1 | 512 | 768} motion + 6 emotion)."""
    return 768 + 1 + {0: 1, 1: 512, 2: 768}[motion_type] + 6


def make_inputs(batch: int, seed: int = 1234, tgt_len: int = 299, src_len: int = 300,
                motion_type: int = 0) -> Dict[str, torch.Tensor]:
    """One synthetic batch with the tensor layouts of run_model_vevo.py:31-45."""
    g = _gen(seed, "inputs")
    d = {}
    d["x"] = torch.randint(0, CHORD_END, (batch, tgt_len), generator=g, dtype=torch.int64)
    d["x_root"] = torch.randint(0, 13, (batch, tgt_len), generator=g, dtype=torch.int64)
    d["x_attr"] = torch.randint(0, 14, (batch, tgt_len), generator=g, dtype=torch.int64)
    d["feature_semantic_list"] = unit_uniform((batch, src_len, 768), g)
    d["feature_key"] = torch.randint(0, 2, (batch, 1), generator=g).float()
    d["feature_scene_offset"] = torch.randint(0, 10, (batch, src_len), generator=g).float()
    if motion_type == 0:
        d["feature_motion"] = torch.rand((batch, src_len), generator=g)
    else:
        d["feature_motion"] = unit_uniform((batch, src_len, {1: 512, 2: 768}[motion_type]), g)
    d["feature_emotion"] = torch.softmax(unit_uniform((batch, src_len, 6), g), dim=-1)
    d["tgt"] = torch.randint(0, CHORD_END, (batch, tgt_len), generator=g, dtype=torch.int64)
    d["tgt_emotion"] = (torch.rand((batch, tgt_len, CHORD_SIZE), generator=g) < 0.2).float()
    return d


def pad_targets(tgt: torch.Tensor, seed: int, max_pad: int = 150) -> torch.Tensor:
    """Ragged targets: the last U(0, max_pad) positions of every row become CHORD_PAD, as the
    data loader pads short chord sequences (dataset/vevo_dataset.py) and CrossEntropyLoss(ignore_index=158) drops them."""
    g = _gen(seed, "pad")
    n = torch.randint(0, min(max_pad, tgt.shape[1] - 1) + 1, (tgt.shape[0],), generator=g)
    out = tgt.clone()
    for b in range(tgt.shape[0]):
        if int(n[b]):
            out[b, tgt.shape[1] - int(n[b]):] = CHORD_PAD
    return out


def fill_like_reference_init(shapes: Dict[str, Tuple[int, ...]], seed: int = 0,
                             wout_gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Seeded values for every entry of a state_dict, given only names/shapes.

    Scales follow what the reference constructor produces (xavier-uniform for
    matrices inside nn.Transformer, video_music_transformer.py:964-971;
    U(-1/sqrt(fan_in), ..) for nn.Linear; unit-variance embeddings) but, unlike
    the constructor, biases and LayerNorm affine terms are non-trivial so that
    parity tests exercise them.  Buffers named '*.pe' are skipped (they are
    deterministic sinusoids built by the module itself).
    """
    out = {}
    for name in sorted(shapes):
        shp = tuple(shapes[name])
        if name.endswith(".pe"):
            continue
        g = _gen(seed, name)
        leaf = name.split(".")[-1]
        if "norm" in name and leaf == "weight" and len(shp) == 1:
            t = 1.0 + 0.1 * unit_uniform(shp, g) / SQRT12 * 2
        elif "norm" in name and leaf == "bias":
            t = 0.1 * unit_uniform(shp, g) / SQRT12 * 2
        elif "embedding" in name and len(shp) == 2:
            t = unit_uniform(shp, g)
        elif leaf == "Er":
            bound = (6.0 / (shp[0] + shp[1])) ** 0.5
            t = unit_uniform(shp, g) / (SQRT12 / 2) * bound * 4.0
        elif len(shp) >= 2:
            fan_out, fan_in = shp[0], shp[1]
            bound = (6.0 / (fan_in + fan_out)) ** 0.5
            t = unit_uniform(shp, g) / (SQRT12 / 2) * bound
            if name.startswith("Wout"):
                t = t * wout_gain
        else:
            t = 0.05 * unit_uniform(shp, g) / (SQRT12 / 2)
        out[name] = t.contiguous()
    return out


def checksum(sd: Dict[str, torch.Tensor]) -> float:
    """Order-independent fingerprint used by the goldens to prove that the
    weights regenerated on another machine are the ones the reference saw."""
    tot = 0.0
    for k in sorted(sd):
        v = sd[k]
        if v.is_floating_point():
            tot += float(v.double().abs().sum()) + 3.0 * float(v.double().sum())
    return tot
