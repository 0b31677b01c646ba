"""CPU: properties of the generated sm_100a SASS that the measurements of round 2 depend on and that the CUDA source cannot
show (profiles/README.md, session 5) -- checked with cuobjdump on the built library:

* the dense contractions really run on tcgen05 / TMEM / TMA (UTCHMMA, LDTM / STTM, UTMALDG / UTMASTG), the CTA-pair GEMM on
  the 2-CTA forms (UTCHMMA.2CTA, UTMALDG.2D.2CTA, UTCBAR.2CTA.MULTICAST);
* the issuer warps of those kernels are free of ELECT / R2UR "waterfall" loops (BRA.U.ANY): inside `if (lane == 0)` with C++
  spin loops ptxas wrapped every tcgen05 / TMA instruction in one (~25 instructions each), which made the issuing thread --
  not the tensor pipe -- the limiter of the GEMM and both attention kernels;
* the shared-memory scratch of the attention forward is accessed with LDS / STS, not with generic LD.E / ST.E (the re-aligned
  dynamic shared-memory base had cost the compiler the address space).
"""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "video2music_b200", "csrc", "libv2m_b200.so")
PAT = {
    "mma": re.compile(r"\bUTCHMMA\b"), "mma2": re.compile(r"\bUTCHMMA\.2CTA\b"), "ldtm": re.compile(r"\bLDTM\b"),
    "sttm": re.compile(r"\bSTTM\b"), "tma_ld": re.compile(r"\bUTMALDG\b"), "tma_ld2": re.compile(r"\bUTMALDG\.\dD\.2CTA\b"),
    "tma_st": re.compile(r"\bUTMASTG\b"), "commit_mc": re.compile(r"\bUTCBAR\.2CTA\.MULTICAST\b"),
    "waterfall": re.compile(r"\bBRA\.U\.ANY\b"), "generic": re.compile(r"(?<![A-Z])(LD|ST)\.E\b"),
    "lds": re.compile(r"\bLDS\b"), "hopper": re.compile(r"\b(HGMMA|WGMMA)\b"),
}


@pytest.fixture(scope="module")
def sass():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    if not os.path.exists(LIB):
        pytest.skip("library not built")
    txt = subprocess.run([exe, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = {}
    fn = None
    for line in txt.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
            per[fn] = {k: 0 for k in PAT}
        elif fn is not None and any(t in line for t in ("UTC", "UTMA", "LDTM", "STTM", "BRA.U", "LD.E", "ST.E", "LDS", "GMMA")):
            for k, p in PAT.items():
                if p.search(line):
                    per[fn][k] += 1
    assert per, "no kernels found in the library"
    return per


def _kernels(sass, needle):
    ks = {f: c for f, c in sass.items() if needle in f}
    assert ks, "no kernel matching %s" % needle
    return ks


def test_library_is_sm100_only_code(sass):
    assert all(c["hopper"] == 0 for c in sass.values())


def test_gemm_runs_on_tcgen05_with_a_lean_issue_loop(sass):
    ks = _kernels(sass, "gemm_bf16_tc_kernel")
    assert len(ks) >= 10                                   # BN x A / B major x pair
    for f, c in ks.items():
        assert c["mma"] > 0 and c["tma_ld"] > 0 and c["ldtm"] > 0 and c["tma_st"] > 0, f
        assert c["waterfall"] <= 2, (f, c)                 # (the epilogue's TMA store keeps one)
        assert c["generic"] <= 2, (f, c)


def test_gemm_pair_kernel_uses_the_2cta_forms(sass):
    pair = {f: c for f, c in _kernels(sass, "gemm_bf16_tc_kernel").items() if c["mma2"] > 0}
    assert len(pair) == 2                                  # K-major and MN-major B
    for f, c in pair.items():
        assert c["mma2"] == c["mma"], (f, c)               # every MMA of the pair kernel is the 2-CTA form
        assert c["tma_ld2"] > 0 and c["commit_mc"] > 0, (f, c)


def test_attention_forward_tcgen05_lean_and_shared_space(sass):
    ks = _kernels(sass, "attn_bf16_tc_kernel")
    assert len(ks) == 6
    for f, c in ks.items():
        assert c["mma"] > 0 and c["tma_ld"] > 0 and c["ldtm"] > 0 and c["sttm"] > 0, f      # P goes back to TMEM
        assert c["waterfall"] == 0, (f, c)
        assert c["generic"] <= 1, (f, c)
    rpr = [c for f, c in ks.items() if "ILi2ELb1E" in f]   # <2 softmax warps, HAS_ER>
    assert len(rpr) == 2 and all(c["lds"] >= 32 for c in rpr)     # the skew is a shifted read of shared memory


def test_attention_backward_tcgen05_lean(sass):
    ks = _kernels(sass, "attn_bwd_tc5_kernel")
    for f, c in ks.items():
        assert c["mma"] > 0 and c["tma_ld"] > 0 and c["ldtm"] > 0, f
        assert c["waterfall"] == 0, (f, c)


def test_pscan_streams_through_tma(sass):
    for f, c in _kernels(sass, "pscan_tma_kernel").items():
        assert c["tma_ld"] > 0 and c["tma_st"] > 0, f
