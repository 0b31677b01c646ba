"""Timings (CUDA events, L2 flushed between launches) of the training paths of the MoE and Mamba variants at the BASELINE
shapes: MoE layer forward / forward+backward (fp32 exact path), fused selective scan forward / backward, conv backward.
usage: python tools/prof_train_variants.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video2music_b200 import GLUExpert, MoELayer, ops
from video2music_b200.mamba import MambaBlock, MambaConfig

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(name, fn, reps=3, bytes_=None, flops=None):
    fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    extra = ("  %.0f GB/s" % (bytes_ / ms / 1e6) if bytes_ else "") + ("  %.1f TFLOP/s" % (flops / ms / 1e9) if flops else "")
    print("%-58s %9.3f ms%s" % (name, ms, extra), flush=True)


g = torch.Generator(device="cpu").manual_seed(5)
# ---- MoE layer, 64 videos x 300 tokens, 6 experts top-2, d 512, ff 1024 (BASELINE config 4)
T, d, ff = 64 * 300, 512, 1024
moe = MoELayer(GLUExpert(d, ff, 0.0), d, n_experts=6, n_experts_per_token=2, dropout=0.0).to(dev).train()
x = torch.randn(300, 64, d, generator=g).to(dev)
with torch.no_grad():
    timed("MoE fp32 forward (inference path) %d tokens" % T, lambda: moe(x), flops=2.0 * 2 * T * 3 * d * ff)
xg = x.clone().requires_grad_(True)


def moe_step():
    moe.zero_grad(set_to_none=True)
    xg.grad = None
    moe(xg).sum().backward()


timed("MoE fp32 forward + backward %d tokens" % T, moe_step, flops=3 * 2.0 * 2 * T * 3 * d * ff)
moe.compute_dtype = torch.bfloat16
timed("MoE bf16 tensor-core forward + backward %d tokens" % T, moe_step, flops=3 * 2.0 * 2 * T * 3 * d * ff)
from video2music_b200 import SharedMoELayer
smoe = SharedMoELayer(GLUExpert(d, ff, 0.0), d, n_experts=6, n_experts_per_token=2, dropout=0.0).to(dev).train()
smoe.compute_dtype = torch.bfloat16


def smoe_step():
    smoe.zero_grad(set_to_none=True)
    xg.grad = None
    smoe(xg).sum().backward()


timed("SharedMoE bf16 tensor-core forward + backward %d tokens" % T, smoe_step, flops=3 * 2.0 * 3 * T * 3 * d * ff)
moe.compute_dtype = torch.float32
# ---- fused selective scan forward / backward and conv backward (BASELINE config 5 shapes)
ED, N, R = 256, 16, 8
for (B, L) in [(64, 300), (8, 4096)]:
    M = B * L
    xz = torch.randn(M, 2 * ED, generator=g).to(dev)
    xc = torch.randn(M, ED, generator=g).to(dev)
    dr = (torch.randn(M, ED, generator=g) - 1.0).to(dev)
    dbc = torch.randn(M, R + 2 * N, generator=g).to(dev)
    A_log = torch.log(torch.arange(1, N + 1).float()).repeat(ED, 1).to(dev)
    D, dtb, dout = torch.ones(ED, device=dev), torch.zeros(ED, device=dev), torch.randn(M, ED, generator=g).to(dev)
    cw, cb = torch.randn(ED, 4, generator=g).to(dev), torch.zeros(ED, device=dev)
    timed("selective_scan fwd (%d,%d,256,16)" % (B, L),
          lambda: ops.selective_scan(xc, dr, dtb, A_log, dbc[:, R:R + N], dbc[:, R + N:], D, xz[:, ED:], B, L), bytes_=4.0 * M * (4 * ED + 2 * N))
    ddbc, dxz = torch.zeros_like(dbc), torch.empty_like(xz)
    timed("selective_scan bwd (%d,%d,256,16)" % (B, L),
          lambda: ops.selective_scan_bwd(xc, dr, dtb, A_log, dbc[:, R:R + N], dbc[:, R + N:], D, xz[:, ED:], dout, ddbc[:, R:R + N],
                                         ddbc[:, R + N:], dxz[:, ED:], B, L), bytes_=4.0 * M * (8 * ED + 4 * N + 2 * ED * N))
    timed("mamba_conv_silu bwd (%d,%d,256)" % (B, L), lambda: ops.mamba_conv_silu_bwd(xz, ED, cw, cb, dout, dxz, B, L), bytes_=4.0 * M * 3 * ED)
    blk = MambaBlock(MambaConfig(d_model=128, n_layers=1)).to(dev).train()
    xb = torch.randn(B, L, 128, generator=g).to(dev).requires_grad_(True)

    def blk_step():
        blk.zero_grad(set_to_none=True)
        xb.grad = None
        blk(xb).sum().backward()

    timed("MambaBlock d_model 128 forward + backward (%d,%d)" % (B, L), blk_step)
torch.cuda.synchronize()
print("done")
