"""CPU: the C-ABI shared library loads and exports every symbol include/v2m_b200.h declares; the
product package never imports the oracle; ops refuse CPU tensors (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "v2m_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(v2m_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from video2music_b200 import _lib
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), "libv2m_b200.so does not export %s" % s
    assert set(syms) == set(_lib.EXPORTS)
    assert lib.v2m_abi_version() == 1


def test_struct_sizes_match_c_layout():
    from video2music_b200 import _lib
    lib = _lib.load()
    for which, cls in enumerate((_lib.Epilogue, _lib.Attn, _lib.DecLayer, _lib.Decode)):
        assert lib.v2m_struct_size(which) == ctypes.sizeof(cls), cls.__name__
    assert ctypes.sizeof(_lib.DecLayer) == 24 * 8


def test_no_cpu_fallback():
    from video2music_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.linear(torch.zeros(4, 8), torch.zeros(3, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.pscan_fwd(torch.zeros(1, 2, 2, 16), torch.zeros(1, 2, 2, 16))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "video2music_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in re.sub(r"#.*", "", src).replace("oracle/ref_shim.py", ""), f


def test_state_dict_layout():
    from video2music_b200 import VideoMusicTransformer
    m = VideoMusicTransformer(total_vf_dim=776, rpr=True)
    sd = m.state_dict()
    assert len(sd) == 207                                  # SURVEY.md section 5 (checkpoint compatibility)
    assert sum(p.numel() for p in m.parameters()) == 32517310
    assert tuple(sd["transformer.decoder.layers.0.self_attn.Er"].shape) == (300, 64)
    import copy
    copy.deepcopy(m)                                       # modules must be deepcopy-safe (rpr.py:20)
