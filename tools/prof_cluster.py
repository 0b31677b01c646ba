"""Cluster decode kernel: a few positions at t=150, B=64 (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from video2music_b200 import engine, synthetic as syn
dev = torch.device("cuda", 0)
model, _ = bench.make_model(torch.bfloat16, dev)
inp = syn.make_inputs(64, 1234, 299, 300, 0)
d = {k: v.to(dev) for k, v in inp.items()}
prim, pr, pa = torch.tensor([1]), torch.tensor([1]), torch.tensor([0])
st = engine.build_decode(model._w(), model._cfg(), d["feature_semantic_list"], d["feature_key"].reshape(-1),
                         d["feature_scene_offset"], d["feature_motion"], d["feature_emotion"], prim, pr, pa, 300)
step_t = st.step
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for _ in range(2):
    step_t.fill_(150); st.pos = 150
    engine.run_decode(st, n, mode="cluster")
torch.cuda.synchronize()
print("ok")
