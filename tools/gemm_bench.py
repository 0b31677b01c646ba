"""tcgen05 GEMM vs cuBLAS (torch.matmul) on the AMT shapes: python tools/gemm_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video2music_b200 import ops
shapes = [(19200, 1536, 512), (19200, 512, 512), (19200, 1024, 512), (19200, 512, 1024), (153600, 1536, 512), (153600, 512, 1024),
          (19200, 776, 776)]
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for M, N, K in shapes:
    Kp = (K + 7) // 8 * 8
    a = torch.randn(M, Kp, device="cuda").bfloat16()
    w = torch.randn(N, Kp, device="cuda").bfloat16()
    b = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda").bfloat16()
    t_plain = timeit(lambda: ops.linear(a, w, None, k=K, out_dtype=torch.bfloat16))
    t_bias = timeit(lambda: ops.linear(a, w, b, k=K, out_dtype=torch.bfloat16))
    t_res = timeit(lambda: ops.linear(a, w, b, k=K, residual=res, out_dtype=torch.bfloat16))
    t_f32 = timeit(lambda: ops.linear(a, w, b, k=K, out_dtype=torch.float32))
    t_cublas = timeit(lambda: torch.matmul(a[:, :K], w[:, :K].t()))
    gf = 2.0 * M * N * K / 1e9
    print("M=%6d N=%4d K=%4d  ours plain %6.1f us (%5.0f TF/s)  +bias %6.1f  +bias+res %6.1f  f32-out %6.1f | cuBLAS %6.1f us (%5.0f TF/s)" %
          (M, N, K, t_plain, gf / t_plain * 1e3, t_bias, t_res, t_f32, t_cublas, gf / t_cublas * 1e3))
