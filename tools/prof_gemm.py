"""One tcgen05 GEMM shape a few times (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video2music_b200 import ops
M, N, K = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (19200, 1536, 512)))
a = torch.randn(M, K, device="cuda").bfloat16()
w = torch.randn(N, K, device="cuda").bfloat16()
b = torch.randn(N, device="cuda")
for _ in range(4):
    y = ops.linear(a, w, b, out_dtype=torch.bfloat16)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
