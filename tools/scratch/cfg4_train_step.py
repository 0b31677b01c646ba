"""Config-4 (GQA + MoE encoder, bf16 experts) training step: a few steps for an ncu launch list / live timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn as nn
from video2music_b200 import (GLUExpert, MoELayer, MultiheadGQA, TransformerEncoder, TransformerEncoderLayer)
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
layer = TransformerEncoderLayer(MultiheadGQA(512, 8, 2, dropout=0.0), MoELayer(GLUExpert(512, 1024, 0.0), 512, n_experts=6,
                                n_experts_per_token=2, dropout=0.0), pre_norm=False, norm=nn.LayerNorm(512), dropout=0.0)
enc = TransformerEncoder(layer, 6, nn.LayerNorm(512)).to(dev).train()
for mod in enc.modules():
    if isinstance(mod, (MoELayer, MultiheadGQA)) and (len(sys.argv) < 3 or sys.argv[2] == 'bf16'):
        mod.compute_dtype = torch.bfloat16
src = torch.randn(300, B, 512, generator=torch.Generator().manual_seed(11)).to(dev)
def f():
    enc.zero_grad(set_to_none=True)
    enc(src).sum().backward()
for _ in range(2): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): f()
e1.record(); e1.synchronize()
print("B=%d: %.2f ms per step" % (B, e0.elapsed_time(e1) / 3))
