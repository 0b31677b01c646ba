"""A/B timing of the pscan kernels (old register / two-pass kernels vs the TMA streaming kernel), CUDA events, L2 flushed
between launches.  Env: V2M_PSCAN_TMA=0 selects the old kernels, V2M_PSCAN_CPI / V2M_PSCAN_LC the item width / chunk."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from video2music_b200 import ops

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timed(fn, n=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]

tag = "tma=%s cpi=%s lc=%s" % (os.environ.get("V2M_PSCAN_TMA", "1"), os.environ.get("V2M_PSCAN_CPI", "auto"), os.environ.get("V2M_PSCAN_LC", "auto"))
for (b, l) in ((64, 300), (8, 4096)):
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.rand(b, l, 256, 16, device="cuda", generator=g) * 0.5 + 0.45
    X = torch.rand(b, l, 256, 16, device="cuda", generator=g) - 0.5
    ms = timed(lambda: ops.pscan_fwd(A, X))
    H = ops.pscan_fwd(A, X)
    print("%s  fwd (%d,%d): %.3f ms  %.0f GB/s" % (tag, b, l, ms, 12.0 * A.numel() / ms / 1e6))
    ms = timed(lambda: ops.pscan_bwd(A, H, X))
    print("%s  bwd (%d,%d): %.3f ms  %.0f GB/s" % (tag, b, l, ms, 20.0 * A.numel() / ms / 1e6))
    # correctness against a float64 sequential scan on a slice
    Hs = torch.zeros(l, 4, 16, dtype=torch.float64, device="cuda")
    a64, x64 = A[b - 1, :, 100:104].double(), X[b - 1, :, 100:104].double()
    h = torch.zeros(4, 16, dtype=torch.float64, device="cuda")
    for t in range(l):
        h = a64[t] * h + x64[t]
        Hs[t] = h
    print("   fwd max err vs float64 slice: %.3e" % (H[b - 1, :, 100:104].double() - Hs).abs().max().item())
