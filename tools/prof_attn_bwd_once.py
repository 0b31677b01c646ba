"""One launch of the attention backward (cross-attention shape, B=512) for ncu: python tools/prof_attn_bwd_once.py [rpr]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video2music_b200 import ops
rpr = len(sys.argv) > 1 and sys.argv[1] == "rpr"
B, L, S, H, dh, E = 512, 299, 299 if rpr else 300, 8, 64, 512
dev = torch.device("cuda", 0)
g = torch.Generator(device="cpu").manual_seed(3)
qkv = (torch.randn(B, L, 3 * E, generator=g) * 0.3).to(dev).bfloat16()
kv = (torch.randn(B, S, 2 * E, generator=g) * 0.3).to(dev).bfloat16()
dO = (torch.randn(B, L, E, generator=g) * 0.3).to(dev).bfloat16()
Er = (torch.randn(300, dh, generator=g) * 0.3).to(dev).bfloat16() if rpr else None
out = torch.empty(B, L, E, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B * H, L, device=dev, dtype=torch.float32)
st, sk = (L * 3 * E, 3 * E), (S * 2 * E, 2 * E)
if rpr:
    k_, v_, ks = qkv[:, :, E:], qkv[:, :, 2 * E:], st
else:
    k_, v_, ks = kv, kv[:, :, E:], sk
for _ in range(2):
    ops.attention(qkv, k_, v_, out, B=B, Hq=H, Hkv=H, Lq=L, Lk=S, dh=dh, q_strides=st, k_strides=ks, v_strides=ks, o_strides=(L * E, E),
                  causal=rpr, Er=Er, lse=lse)
    dq = torch.empty_like(qkv)
    dkv = torch.empty_like(qkv if rpr else kv)
    der = torch.zeros(300, dh, device=dev) if rpr else None
    ops.attention_bwd(qkv, k_, v_, out, dO, lse, Er, dq, dkv[:, :, E:] if rpr else dkv, dkv[:, :, 2 * E:] if rpr else dkv[:, :, E:], der, B=B, Hq=H, Hkv=H,
                      Lq=L, Lk=S, dh=dh, q_strides=st, k_strides=ks, v_strides=ks, o_strides=(L * E, E), do_strides=(L * E, E),
                      dq_strides=st, dkv_strides=ks, causal=rpr, tensor_core=True)
torch.cuda.synchronize()
print("ok")
