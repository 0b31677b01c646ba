"""Small streamed-decode run for compute-sanitizer: python tools/stream_debug.py [B] [t0] [steps] [rows]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from video2music_b200 import engine, synthetic as syn
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
t0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
if len(sys.argv) > 4:
    os.environ["V2M_STREAM_ROWS"] = sys.argv[4]
os.environ["V2M_VERBOSE"] = "1"
dev = torch.device("cuda", 0)
model, _ = bench.make_model(torch.bfloat16, dev)
inp = syn.make_inputs(B, 1234, 299, 300, 0)
d = {k: v.to(dev) for k, v in inp.items()}
prim, pr, pa = torch.tensor([1]), torch.tensor([1]), torch.tensor([0])
st = engine.build_decode(model._w(), model._cfg(), d["feature_semantic_list"], d["feature_key"].reshape(-1),
                         d["feature_scene_offset"], d["feature_motion"], d["feature_emotion"], prim, pr, pa, 300, mode="stream")
torch.cuda.synchronize()
print("built", flush=True)
st.step.fill_(t0); st.pos = t0
marks = torch.zeros(8192, dtype=torch.int64).pin_memory()
try:
    engine.run_decode(st, 0, mode="stream")
    import ctypes as C
    from video2music_b200 import _lib
    engine.run_decode(st, steps, mode="stream", timestamps=marks)
    torch.cuda.synchronize()
finally:
    m = marks[4096:4096 + 16 * 16].view(16, 16)
    print("markers per CTA (rows) / warp (cols 0..9):")
    for b in range(16):
        print(b, m[b, :10].tolist())
print("ok", st.gen[:, :t0 + steps + 1].tolist()[:2], flush=True)
