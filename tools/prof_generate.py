"""Short B=64 bf16 generation (for the ncu launch list): prefill + a few decode steps at a late position."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from video2music_b200 import engine, synthetic as syn
seq = int(sys.argv[1]) if len(sys.argv) > 1 else 300
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
start = int(sys.argv[3]) if len(sys.argv) > 3 else 150
dt = torch.bfloat16 if (len(sys.argv) <= 4 or sys.argv[4] == "bf16") else torch.float32
dev = torch.device("cuda", 0)
model, _ = bench.make_model(dt, dev)
inp = syn.make_inputs(64, 1234, 299, 300, 0)
d = {k: v.to(dev) for k, v in inp.items()}
prim, pr, pa = torch.tensor([1]), torch.tensor([1]), torch.tensor([0])
st = engine.build_decode(model._w(), model._cfg(), d["feature_semantic_list"], d["feature_key"].reshape(-1),
                         d["feature_scene_offset"], d["feature_motion"], d["feature_emotion"], prim, pr, pa, seq,
                         mode=(sys.argv[5] if len(sys.argv) > 5 else 'kernels'))
torch.cuda.synchronize()
st.keep[-1]  # noqa
# jump to a mid-sequence position: the caches hold zeros there, which costs the same bytes
import ctypes
step_t = st.step
step_t.fill_(start)
st.pos = start
engine.run_decode(st, steps, use_graph=False, mode=(sys.argv[5] if len(sys.argv) > 5 else 'kernels'))
torch.cuda.synchronize()
print("ok", int(step_t[0].item()))
