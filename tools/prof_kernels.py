"""The tensor-core and scan kernels at the BASELINE shapes, each a few times, with CUDA-event timings
(run alone) -- and the same process under ncu for the tensor-pipe / DRAM counters.
usage: python tools/prof_kernels.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video2music_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(name, fn, flops=None, bytes_=None, reps=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()                                   # L2 flush between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    extra = ""
    if flops:
        extra += "  %.1f TFLOP/s" % (flops / ms / 1e9)
    if bytes_:
        extra += "  %.0f GB/s" % (bytes_ / ms / 1e6)
    print("%-46s %8.3f ms%s" % (name, ms, extra), flush=True)


g = torch.Generator(device="cpu").manual_seed(3)
L, S, H, dh, E, FF = 299, 300, 8, 64, 512, 1024
# ---- GEMMs of one layer at M = B*L tokens
M = B * L
a = torch.randn(M, E, generator=g).to(dev).bfloat16()
for (N, K, nm) in [(3 * E, E, "in_proj"), (FF, E, "linear1+relu"), (E, FF, "linear2"), (E, E, "out_proj")]:
    x = a if K == E else torch.randn(M, K, generator=g).to(dev).bfloat16()
    w = (torch.randn(N, K, generator=g) * 0.05).to(dev).bfloat16()
    b = torch.randn(N, generator=g).to(dev)
    timed("gemm_bf16_tc %s M=%d N=%d K=%d" % (nm, M, N, K), lambda: ops.linear(x, w, b, out_dtype=torch.bfloat16, relu=(nm == "linear1+relu")),
          flops=2.0 * M * N * K)
# ---- attention forward: RPR causal self-attention and cross-attention, (B, L, 3E) packed projections
qkv = (torch.randn(B, L, 3 * E, generator=g) * 0.3).to(dev).bfloat16()
Er = (torch.randn(300, dh, generator=g) * 0.3).to(dev).bfloat16()
out = torch.empty(B, L, E, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B * H, L, device=dev, dtype=torch.float32)
st = (L * 3 * E, 3 * E)
timed("attn_bf16_tc RPR causal self B=%d L=%d" % (B, L),
      lambda: ops.attention(qkv, qkv[:, :, E:], qkv[:, :, 2 * E:], out, B=B, Hq=H, Hkv=H, Lq=L, Lk=L, dh=dh, q_strides=st, k_strides=st,
                            v_strides=st, o_strides=(L * E, E), causal=True, Er=Er, lse=lse), flops=6.0 * L * L * dh * B * H)
kv = (torch.randn(B, S, 2 * E, generator=g) * 0.3).to(dev).bfloat16()
sk = (S * 2 * E, 2 * E)
timed("attn_bf16_tc cross B=%d L=%d S=%d" % (B, L, S),
      lambda: ops.attention(qkv, kv, kv[:, :, E:], out, B=B, Hq=H, Hkv=H, Lq=L, Lk=S, dh=dh, q_strides=st, k_strides=sk, v_strides=sk,
                            o_strides=(L * E, E), causal=False, lse=lse), flops=4.0 * L * S * dh * B * H)
# ---- attention backward (tensor-core paths): RPR causal self-attention, encoder self-attention, cross-attention
dO = (torch.randn(B, L, E, generator=g) * 0.3).to(dev).bfloat16()
dqkv = torch.empty_like(qkv)
der = torch.zeros(300, dh, device=dev)
ops.attention(qkv, qkv[:, :, E:], qkv[:, :, 2 * E:], out, B=B, Hq=H, Hkv=H, Lq=L, Lk=L, dh=dh, q_strides=st, k_strides=st, v_strides=st,
              o_strides=(L * E, E), causal=True, Er=Er, lse=lse)
timed("attn_bwd RPR causal self B=%d L=%d" % (B, L),
      lambda: ops.attention_bwd(qkv, qkv[:, :, E:], qkv[:, :, 2 * E:], out, dO, lse, Er, dqkv, dqkv[:, :, E:], dqkv[:, :, 2 * E:], der,
                                B=B, Hq=H, Hkv=H, Lq=L, Lk=L, dh=dh, q_strides=st, k_strides=st, v_strides=st, o_strides=(L * E, E),
                                do_strides=(L * E, E), dq_strides=st, dkv_strides=st, causal=True, tensor_core=True),
      flops=10.0 * L * L * dh * B * H)
ops.attention(qkv, kv, kv[:, :, E:], out, B=B, Hq=H, Hkv=H, Lq=L, Lk=S, dh=dh, q_strides=st, k_strides=sk, v_strides=sk,
              o_strides=(L * E, E), causal=False, lse=lse)
dkv = torch.empty_like(kv)
for drop in (None, (0.2, 99)):
    timed("attn_bwd cross B=%d L=%d S=%d%s" % (B, L, S, " dropout 0.2" if drop else ""),
          lambda: ops.attention_bwd(qkv, kv, kv[:, :, E:], out, dO, lse, None, dqkv, dkv, dkv[:, :, E:], None,
                                    B=B, Hq=H, Hkv=H, Lq=L, Lk=S, dh=dh, q_strides=st, k_strides=sk, v_strides=sk, o_strides=(L * E, E),
                                    do_strides=(L * E, E), dq_strides=st, dkv_strides=sk, causal=False, tensor_core=True, dropout=drop),
          flops=10.0 * L * S * dh * B * H)
# ---- pscan and the fused selective scan (BASELINE config 5)
for (b_, l_) in [(64, 300), (8, 4096)]:
    A = torch.rand(b_, l_, 256, 16, generator=g).mul_(0.99).to(dev)
    X = torch.randn(b_, l_, 256, 16, generator=g).to(dev)
    timed("pscan fwd (%d,%d,256,16)" % (b_, l_), lambda: ops.pscan_fwd(A, X), bytes_=3.0 * 4 * A.numel())
    Hh = ops.pscan_fwd(A, X)
    timed("pscan bwd (%d,%d,256,16)" % (b_, l_), lambda: ops.pscan_bwd(A, Hh, X), bytes_=5.0 * 4 * A.numel())
    del A, X, Hh
    ED, N, R = 256, 16, 8
    x = torch.randn(b_ * l_, ED, generator=g).to(dev)
    dr = torch.randn(b_ * l_, ED, generator=g).to(dev)
    dbc = torch.randn(b_ * l_, R + 2 * N, generator=g).to(dev)
    z = torch.randn(b_ * l_, ED, generator=g).to(dev)
    A_log = torch.log(torch.arange(1, N + 1).float()).repeat(ED, 1).to(dev)
    D = torch.ones(ED, device=dev)
    dtb = torch.zeros(ED, device=dev)
    timed("selective_scan fused (%d,%d,256,16)" % (b_, l_),
          lambda: ops.selective_scan(x, dr, dtb, A_log, dbc[:, R:R + N], dbc[:, R + N:], D, z, b_, l_),
          bytes_=4.0 * b_ * l_ * (4 * ED + 2 * N))
torch.cuda.synchronize()
print("done")
