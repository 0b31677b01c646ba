"""Sweep (cluster size, clusters) of the persistent decode kernel at B=64, t=150."""
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
os.environ["V2M_VERBOSE"] = "1"
os.environ.pop("V2M_CLUSTER", None)
step_t.fill_(150); st.pos = 150
engine.run_decode(st, 2, mode="cluster")
torch.cuda.synchronize()
os.environ.pop("V2M_VERBOSE")
cfgs = [(16, 7), (16, 4), (15, 8), (14, 8), (12, 8), (12, 11), (10, 8), (10, 13), (9, 16), (8, 16), (8, 8), (8, 18), (6, 22), (4, 32)]
for cs, ncl in cfgs:
    os.environ["V2M_CLUSTER"] = "%d,%d" % (cs, ncl)
    try:
        step_t.fill_(150); st.pos = 150
        engine.run_decode(st, 3, mode="cluster")
        torch.cuda.synchronize()
        step_t.fill_(150); st.pos = 150
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); engine.run_decode(st, 20, mode="cluster"); e1.record(); e1.synchronize()
        print("cs=%2d ncl=%2d rows=%2d : %.1f us/step" % (cs, ncl, -(-64 // ncl), e0.elapsed_time(e1) * 1e3 / 20), flush=True)
    except Exception as ex:
        print("cs=%d ncl=%d failed: %s" % (cs, ncl, str(ex)[:100]), flush=True)
        break
