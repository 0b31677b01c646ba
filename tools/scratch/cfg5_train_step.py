"""Config-5 (VideoRegression, 6 Bi-Mamba+ layers) training step: live timing and kernels for an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from video2music_b200 import VideoRegression
dev = torch.device("cuda", 0)
torch.manual_seed(0)
g = torch.Generator().manual_seed(11)
reg = VideoRegression(n_layers=6, d_model=128, d_hidden=256, dropout=0.0, total_vf_dim=774, regModel="bimamba+").to(dev).train()
sem, emo = torch.randn(64, 300, 768, generator=g).to(dev), torch.softmax(torch.randn(64, 300, 6, generator=g), -1).to(dev)
zz = torch.zeros(64, 300, device=dev)
import sys
if len(sys.argv) > 1 and sys.argv[1] == 'bf16':
    from video2music_b200.mamba import MambaBlock, _FFN
    for mod in reg.modules():
        if isinstance(mod, (MambaBlock, _FFN)): mod.compute_dtype = torch.bfloat16
def f():
    reg.zero_grad(set_to_none=True)
    ln, inst = reg(sem, zz, zz, emo)
    (ln.sum() + inst.sum()).backward()
for _ in range(2): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): f()
e1.record(); e1.synchronize()
print("%.2f ms per step" % (e0.elapsed_time(e1) / 3))
