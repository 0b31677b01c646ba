"""Streamed decode kernel experiments: us per position for B videos over positions [t0, t0 + n) (env switches are read by
decode_run_stream: V2M_STREAM_ROWS, V2M_STREAM_PF, V2M_STREAM_PF_WHAT)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from video2music_b200 import engine, synthetic as syn
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
t0 = int(sys.argv[2]) if len(sys.argv) > 2 else 100
n = int(sys.argv[3]) if len(sys.argv) > 3 else 100
dev = torch.device("cuda", 0)
model, _ = bench.make_model(torch.bfloat16, dev)
inp = syn.make_inputs(B, 1234, 299, 300, 0)
d = {k: v.to(dev) for k, v in inp.items()}
prim, pr, pa = torch.tensor([1]), torch.tensor([1]), torch.tensor([0])
st = engine.build_decode(model._w(), model._cfg(), d["feature_semantic_list"], d["feature_key"].reshape(-1),
                         d["feature_scene_offset"], d["feature_motion"], d["feature_emotion"], prim, pr, pa, 300, mode="stream")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
best = None
for rep in range(3):
    st.step.fill_(t0); st.pos = t0
    flush.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    engine.run_decode(st, n, mode="stream")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    best = ms if best is None else min(best, ms)
env = {k: v for k, v in os.environ.items() if k.startswith("V2M_")}
print("B=%d positions [%d,%d): %.1f us per position  %s" % (B, t0, t0 + n, best * 1000.0 / n, env), flush=True)
