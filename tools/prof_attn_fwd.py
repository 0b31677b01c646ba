"""Attention forward alone (RPR causal self-attention, with / without dropout, and cross-attention) at the BASELINE shape,
CUDA events, L2 flushed between launches.  usage: [V2M_ATTN_SHIFT=0] python tools/prof_attn_fwd.py [batch] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video2music_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 9
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(name, fn, flops):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(REPS):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print("%-52s %8.3f ms  %.1f TFLOP/s dense-equivalent  (min %.3f)" % (name, ms, flops / ms / 1e9, min(ts)), flush=True)


g = torch.Generator(device="cpu").manual_seed(3)
L, S, H, dh, E = 299, 300, 8, 64, 512
qkv = (torch.randn(B, L, 3 * E, generator=g) * 0.3).to(dev).bfloat16()
Er = (torch.randn(300, dh, generator=g) * 0.3).to(dev).bfloat16()
out = torch.empty(B, L, E, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B * H, L, device=dev, dtype=torch.float32)
st = (L * 3 * E, 3 * E)
print("V2M_ATTN_SHIFT=%s" % os.environ.get("V2M_ATTN_SHIFT", "(default 1)"))
for drop in (None, (0.2, 99)):
    timed("attn_bf16_tc RPR causal self B=%d L=%d%s" % (B, L, " dropout 0.2" if drop else ""),
          lambda: ops.attention(qkv, qkv[:, :, E:], qkv[:, :, 2 * E:], out, B=B, Hq=H, Hkv=H, Lq=L, Lk=L, dh=dh, q_strides=st, k_strides=st,
                                v_strides=st, o_strides=(L * E, E), causal=True, Er=Er, lse=lse, dropout=drop), flops=6.0 * L * L * dh * B * H)
kv = (torch.randn(B, S, 2 * E, generator=g) * 0.3).to(dev).bfloat16()
sk = (S * 2 * E, 2 * E)
timed("attn_bf16_tc cross B=%d L=%d S=%d" % (B, L, S),
      lambda: ops.attention(qkv, kv, kv[:, :, E:], out, B=B, Hq=H, Hkv=H, Lq=L, Lk=S, dh=dh, q_strides=st, k_strides=sk, v_strides=sk,
                            o_strides=(L * E, E), causal=False, lse=lse), flops=4.0 * L * S * dh * B * H)
timed("attn_bf16_tc causal self, no Er B=%d L=%d" % (B, L),
      lambda: ops.attention(qkv, qkv[:, :, E:], qkv[:, :, 2 * E:], out, B=B, Hq=H, Hkv=H, Lq=L, Lk=L, dh=dh, q_strides=st, k_strides=st,
                            v_strides=st, o_strides=(L * E, E), causal=True, lse=lse), flops=4.0 * L * L * dh * B * H)
torch.cuda.synchronize()
