import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from conftest import amt_state_dict
from video2music_b200 import synthetic as syn
DEV = "cuda:0"
for dtype, mode, ce in [(torch.bfloat16, "kernels", False), (torch.bfloat16, "stream", True), (torch.bfloat16, "stream", False)]:
    B, T, P = 6, 60, 3
    m, sd = amt_state_dict(syn.vf_dim(0), 3, chord_embed=ce, wout_gain=4.0)
    m = m.to(DEV).eval().set_compute_dtype(dtype)
    inp = syn.make_inputs(B, 55, 299, 300, 0)
    prim, pr, pa = inp["x"][:, :P], inp["x_root"][:, :P], inp["x_attr"][:, :P]
    u = torch.rand((B, T), generator=syn._gen(9, "u"))
    gen, logits = m.generate(inp["feature_semantic_list"], inp["feature_key"], inp["feature_scene_offset"], inp["feature_motion"],
                     inp["feature_emotion"], primer=prim, primer_root=pr, primer_attr=pa, target_seq_length=T, beam=0,
                     decode_mode=mode, uniforms=u.to(DEV), return_logits=True)
    print(mode, ce, [r[:10] for r in gen.cpu().tolist()], "finite logits per step:", torch.isfinite(logits).all(dim=-1)[0].cpu().tolist())
