"""One launch of each generation-step kernel at the config-4 shapes (for `ncu --set full`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from video2music_b200 import ops
dev = torch.device("cuda", 0)
torch.manual_seed(0)
x = torch.randn(64, 512, device=dev); w = torch.randn(1536, 512, device=dev) * 0.05; b = torch.randn(1536, device=dev)
K = torch.randn(64, 300, 128, device=dev); V = torch.randn(64, 300, 128, device=dev); q = torch.randn(64, 512, device=dev)
nd = torch.tensor([300], dtype=torch.int32, device=dev)
for _ in range(3):
    ops.step_linear(x, w, b)
    ops.step_attention(q, K, V, Hq=8, Hkv=2, dh=64, n_max=300, kv_strides=(K.stride(0), K.stride(1)), n_dev=nd, q_scale=0.125)
torch.cuda.synchronize()
