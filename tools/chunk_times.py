"""Per-chunk wait / compute times of warp 0 in the attention phases (library built with EXTRA=-DV2M_CHUNK_STAMPS)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from video2music_b200 import engine, synthetic as syn
pos = 150
dev = torch.device("cuda", 0)
model, _ = bench.make_model(torch.bfloat16, dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
inp = syn.make_inputs(B, 1234, 299, 300, 0)
d = {k: v.to(dev) for k, v in inp.items()}
prim, pr, pa = torch.tensor([1]), torch.tensor([1]), torch.tensor([0])
st = engine.build_decode(model._w(), model._cfg(), d["feature_semantic_list"], d["feature_key"].reshape(-1),
                         d["feature_scene_offset"], d["feature_motion"], d["feature_emotion"], prim, pr, pa, 300, mode="stream")
st.step.fill_(pos); st.pos = pos
engine.run_decode(st, 2, mode="stream")
torch.cuda.synchronize()
ts = torch.zeros(4000, dtype=torch.int64, device=dev)
st.step.fill_(pos); st.pos = pos
engine.run_decode(st, 2, mode="stream", timestamps=ts)
torch.cuda.synchronize()
t = [x for x in ts.cpu().tolist() if x > 0]
dts = [(t[i + 1] - t[i]) / 1000.0 for i in range(len(t) - 1)]
print(len(t), "stamps; deltas (us) of the first 120:")
print(" ".join("%.2f" % x for x in dts[:120]))
