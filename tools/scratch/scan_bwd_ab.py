"""A/B of the fused selective-scan backward (V2M_SCAN_BWD_OLD=1: first version with the state workspace) at the config-5
shapes: CUDA-event time of the whole selective_scan_bwd call (5 launches) and the result digest for cross-checking."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from video2music_b200 import ops

dev = torch.device("cuda", 0)
tag = "old" if os.environ.get("V2M_SCAN_BWD_OLD") == "1" else "new"
for (B, L) in ((64, 300), (8, 4096)):
    g = torch.Generator().manual_seed(5)
    ED, N, R = 256, 16, 8
    M = B * L
    xz, xc = torch.randn(M, 2 * ED, generator=g).to(dev), torch.randn(M, ED, generator=g).to(dev)
    dr, dbc = (torch.randn(M, ED, generator=g) - 1.0).to(dev), torch.randn(M, R + 2 * N, generator=g).to(dev)
    A_log = torch.log(torch.arange(1, N + 1).float()).repeat(ED, 1).to(dev)
    D, dtb, dout = torch.ones(ED, device=dev), torch.zeros(ED, device=dev), torch.randn(M, ED, generator=g).to(dev)
    for plus in (False, True):
        def run():
            ddbc, dxz = torch.zeros_like(dbc), torch.empty_like(xz)
            r = ops.selective_scan_bwd(xc, dr, dtb, A_log, dbc[:, R:R + N], dbc[:, R + N:], D, xz[:, ED:], dout, ddbc[:, R:R + N],
                                       ddbc[:, R + N:], dxz[:, ED:], B, L, plus=plus)
            return r, ddbc, dxz
        (dxc, ddraw, dA, dD, ddtb), ddbc, dxz = run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        dig = [float(t.double().abs().sum()) for t in (dxc, ddraw, dA, dD, ddtb, ddbc, dxz[:, ED:])]
        print("%s (%d,%d) plus=%d: %.3f ms (incl. two torch fills)  digest %s" % (tag, B, L, plus, ts[2], " ".join("%.6e" % d for d in dig)))
