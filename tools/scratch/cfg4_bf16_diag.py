"""Diagnostic: bf16 config-4 training step vs the reference golden -- row-wise error quantiles and gradient-norm errors."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conftest import load_golden
from video2music_b200 import synthetic as syn, MoELayer, MultiheadGQA, SharedMoELayer
from test_oracle import _variant_net
DEV = "cuda:0"
_u = lambda shape, seed, name: syn.unit_uniform(shape, syn._gen(seed, name))
for name in ("post_ln_moe", "post_ln_sharedmoe_b2", "pre_rms_moe"):
    g = load_golden("variant_train.pt")[name]; c = g["spec"]
    for which in ("gqa", "moe", "both"):
        net, sd = _variant_net(c); net.load_state_dict(sd); net = net.to(DEV).train()
        for mod in net.modules():
            if (isinstance(mod, MultiheadGQA) and which in ("gqa", "both")) or (isinstance(mod, (MoELayer, SharedMoELayer)) and which in ("moe", "both")):
                mod.compute_dtype = torch.bfloat16
        src = _u((c["S"], c["B"], 512), c["seed"], "src").to(DEV).requires_grad_(True)
        tgt = _u((c["T"], c["B"], 512), c["seed"], "tgt").to(DEV).requires_grad_(True)
        r = _u((c["T"], c["B"], 512), c["seed"], "r").to(DEV)
        y = net["dec"](tgt, net["enc"](src)); (y * r).sum().backward()
        def rows(a, b):
            a, b = a.detach().float().cpu().reshape(-1, 512), b.float().reshape(-1, 512)
            e = (a - b).norm(dim=1) / b.norm(dim=1).clamp_min(1e-12)
            return "median %.1e p90 %.1e max %.1e frac>5e-2 %.3f" % (e.median(), e.quantile(0.9), e.max(), float((e > 5e-2).float().mean()))
        print(name, which, "| out", rows(y, g["out"]), "| d_src", rows(src.grad, g["d_src"]), "| d_tgt", rows(tgt.grad, g["d_tgt"]))
        gmax = max(g["grad_norms"].values())
        errs = sorted(((abs(float(p.grad.double().norm()) - g["grad_norms"][n]) / g["grad_norms"][n], n, g["grad_norms"][n]) for n, p in net.named_parameters()
                       if p.grad is not None and n in g["grad_norms"] and g["grad_norms"][n] > 1e-6 * gmax), reverse=True)
        print("   worst grad-norm errors:", ["%.2e %s (%.1e)" % e for e in errs[:4]])
