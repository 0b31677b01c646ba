"""Where does the bf16 Er gradient differ from the reference's fp32 autograd (full-shape training golden)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from conftest import load_golden
import test_gpu_train as T
for name in ("full", "ragged"):
    g, m, y, loss = T._full_case(name, torch.bfloat16)
    params = dict(m.named_parameters())
    for n, gref in g["grads"].items():
        if not n.endswith(".Er"):
            continue
        mine = params[n].grad.detach().double().cpu()
        d = (mine - gref.double())
        rowerr = d.norm(dim=1)
        rowref = gref.double().norm(dim=1)
        top = torch.topk(rowerr, 5).indices.tolist()
        i = int(d.abs().argmax()) // 64
        print(name, n, "max|ref| %.3e at row %d; max|diff| %.3e at row %d (ref row norm %.3e, mine %.3e); worst rows %s" % (
            float(gref.abs().max()), int(gref.abs().argmax()) // 64, float(d.abs().max()), i, float(rowref[i]), float(mine[i].norm()), top))
        print("   row-norm error by distance band: ", ["%.3f" % float(rowerr[a:b].norm() / rowref[a:b].norm().clamp_min(1e-30)) for a, b in ((0, 50), (50, 150), (150, 250), (250, 290), (290, 300))])
