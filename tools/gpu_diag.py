"""Diagnostics for the first bring-up on a real B200 (run under gpurun; writes to stdout)."""
import sys
import os
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from video2music_b200 import ops, synthetic as syn

DEV = "cuda:0"


def section(name, fn):
    print("=" * 20, name, flush=True)
    try:
        fn()
    except Exception:
        traceback.print_exc()
    torch.cuda.synchronize()
    print(flush=True)


def gemm_tc_patterns():
    """Where does the tcgen05 GEMM go wrong, if it does: per 128x{128,256} tile and per k-block."""
    for (M, N, K) in [(128, 128, 64), (128, 128, 128), (256, 256, 64), (128, 128, 16), (300, 512, 776)]:
        g = syn._gen(1, "diag")
        a = torch.randint(-2, 3, (M, K), generator=g).float()
        w = torch.randint(-2, 3, (N, K), generator=g).float()
        y = ops.linear(a.to(DEV).bfloat16(), w.to(DEV).bfloat16(), None, out_dtype=torch.float32).cpu()
        ref = a @ w.T
        bad = (y != ref)
        print("M%d N%d K%d: mismatches %d / %d, max abs err %.3f" % (M, N, K, int(bad.sum()), bad.numel(),
                                                                      float((y - ref).abs().max())))
        if bad.any():
            rows = bad.any(1).nonzero().flatten()
            cols = bad.any(0).nonzero().flatten()
            print("  bad rows: n=%d first %s last %s" % (len(rows), rows[:8].tolist(), rows[-4:].tolist()))
            print("  bad cols: n=%d first %s last %s" % (len(cols), cols[:8].tolist(), cols[-4:].tolist()))
            print("  y[0,:8]  ", y[0, :8].tolist())
            print("  ref[0,:8]", ref[0, :8].tolist())
            # does y match a partial-K product?
            for kk in range(16, K + 1, 16):
                if torch.equal(y, a[:, :kk] @ w[:, :kk].T):
                    print("  == product over the first %d of %d k" % (kk, K))


def timing():
    """Rough kernel timings with CUDA events (not the bench)."""
    def t(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    for (M, N, K) in [(19200, 1536, 512), (19200, 512, 512), (19200, 1024, 512), (19200, 512, 1024), (8192, 8192, 8192)]:
        a = torch.randn(M, K, device=DEV).bfloat16()
        w = torch.randn(N, K, device=DEV).bfloat16()
        ms = t(lambda: ops.linear(a, w, None, out_dtype=torch.bfloat16))
        ms_t = t(lambda: torch.nn.functional.linear(a, w))
        print("gemm bf16 M%d N%d K%d: ours %.3f ms (%.1f TF/s)  cuBLAS %.3f ms (%.1f TF/s)" % (
            M, N, K, ms, 2 * M * N * K / ms / 1e9, ms_t, 2 * M * N * K / ms_t / 1e9))
    for (M, N, K) in [(19200, 1536, 512)]:
        a, w = torch.randn(M, K, device=DEV), torch.randn(N, K, device=DEV)
        ms = t(lambda: ops.linear(a, w, None), n=5)
        print("gemm f32 M%d N%d K%d: %.3f ms (%.1f TF/s)" % (M, N, K, ms, 2 * M * N * K / ms / 1e9))
    for (B, L, D, N) in [(64, 300, 256, 16), (8, 4096, 256, 16)]:
        A = torch.rand(B, L, D, N, device=DEV)
        X = torch.randn(B, L, D, N, device=DEV)
        ms = t(lambda: ops.pscan_fwd(A, X))
        print("pscan fwd (%d,%d,%d,%d): %.3f ms  %.0f GB/s" % (B, L, D, N, ms, 12 * A.numel() / ms / 1e6))


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), torch.version.cuda)
    section("tcgen05 gemm patterns", gemm_tc_patterns)
    section("timing", timing)
