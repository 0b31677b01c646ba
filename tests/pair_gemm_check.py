"""CTA-pair GEMM (V2M_GEMM_PAIR=2 forces it for every eligible shape; the switch is read once per process, hence a script that
tests/test_gpu_kernels.py runs in a subprocess) against torch.matmul: plain / bias+relu / residual / fp32 out / MN-major B (dX),
including M tails where the second CTA of the last pair owns no valid row."""
import os, sys
os.environ.setdefault("V2M_GEMM_PAIR", "2")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video2music_b200 import ops
torch.manual_seed(0)
dev = "cuda"
def check(name, y, ref, tol=2e-2):
    err = float((y.float() - ref).abs().max() / ref.abs().max())
    print("%-60s rel err %.2e %s" % (name, err, "ok" if err < tol else "FAIL"), flush=True)
    return err < tol
ok = True
for (M, N, K) in [(256, 256, 64), (512, 512, 512), (300, 256, 128), (19136, 512, 512), (19136, 1536, 512), (4000, 1024, 1024), (129, 768, 512)]:
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev).bfloat16()
    ref = a.float() @ w.float().t()
    ok &= check("plain bf16 out M=%d N=%d K=%d" % (M, N, K), ops.linear(a, w, None, out_dtype=torch.bfloat16), ref)
    ok &= check("bias+relu", ops.linear(a, w, b, relu=True, out_dtype=torch.bfloat16), torch.relu(ref + b))
    ok &= check("bias+residual", ops.linear(a, w, b, residual=res, out_dtype=torch.bfloat16), ref + b + res.float())
    ok &= check("fp32 out", ops.linear(a, w, b, out_dtype=torch.float32), ref + b, 1e-2)
    # dX = dY W : a [M, N'] x b stored [N', K'] row-major (MN-major B): C[M, K'] = a @ b
    wt = (torch.randn(K, N, device=dev) * 0.05).bfloat16()      # [K (reduction), N (output cols)]
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    y = ops.linear_general(a, wt, a_mn=False, b_mn=True, M=M, N=N, K=K, out=out)
    ok &= check("MN-major B (dX)", y if y is not None else out, a.float() @ wt.float())
torch.cuda.synchronize()
print("ALL OK" if ok else "SOME FAILED")
