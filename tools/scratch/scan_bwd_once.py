"""One selective_scan_bwd call at (64,300,256,16) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from video2music_b200 import ops
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(5)
B, L, ED, N, R = 64, 300, 256, 16, 8
M = B * L
xz, xc = torch.randn(M, 2 * ED, generator=g).to(dev), torch.randn(M, ED, generator=g).to(dev)
dr, dbc = (torch.randn(M, ED, generator=g) - 1.0).to(dev), torch.randn(M, R + 2 * N, generator=g).to(dev)
A_log = torch.log(torch.arange(1, N + 1).float()).repeat(ED, 1).to(dev)
D, dtb, dout = torch.ones(ED, device=dev), torch.zeros(ED, device=dev), torch.randn(M, ED, generator=g).to(dev)
ddbc, dxz = torch.zeros_like(dbc), torch.empty_like(xz)
out = torch.empty(M, ED, device=dev)
for _ in range(2):
    ops.selective_scan_bwd(xc, dr, dtb, A_log, dbc[:, R:R + N], dbc[:, R + N:], D, xz[:, ED:], dout, ddbc[:, R:R + N], ddbc[:, R + N:],
                           dxz[:, ED:], B, L)
    ops.selective_scan(xc, dr, dtb, A_log, dbc[:, R:R + N], dbc[:, R + N:], D, xz[:, ED:], B, L)
torch.cuda.synchronize()
print("done")
