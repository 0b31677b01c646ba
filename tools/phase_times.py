"""Per-phase timestamps of the streamed decode kernel (thread 0 of CTA 0): python tools/phase_times.py [position] [rows]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from video2music_b200 import engine, synthetic as syn
pos = int(sys.argv[1]) if len(sys.argv) > 1 else 150
if len(sys.argv) > 2:
    os.environ["V2M_STREAM_ROWS"] = sys.argv[2]
os.environ["V2M_VERBOSE"] = "1"
dev = torch.device("cuda", 0)
model, _ = bench.make_model(torch.bfloat16, dev)
inp = syn.make_inputs(64, 1234, 299, 300, 0)
d = {k: v.to(dev) for k, v in inp.items()}
prim, pr, pa = torch.tensor([1]), torch.tensor([1]), torch.tensor([0])
st = engine.build_decode(model._w(), model._cfg(), d["feature_semantic_list"], d["feature_key"].reshape(-1),
                         d["feature_scene_offset"], d["feature_motion"], d["feature_emotion"], prim, pr, pa, 300, mode="stream")
st.step.fill_(pos); st.pos = pos
engine.run_decode(st, 2, mode="stream")
torch.cuda.synchronize()
per_step = 2 + 8 * 6                       # stamps per position (decode_stream.cu): embed, 8 per layer, logits+argmax
ts = torch.zeros(1 + 4 * per_step, dtype=torch.int64, device=dev)
st.step.fill_(pos); st.pos = pos
engine.run_decode(st, 4, mode="stream", timestamps=ts)
torch.cuda.synchronize()
t = ts.cpu().tolist()
dts = [(t[i + 1] - t[i]) / 1000.0 for i in range(len(t) - 1)]
labels = ["embed"] + sum([["L%d qkv" % l, "L%d self-attn" % l, "L%d so+ln" % l, "L%d cq" % l, "L%d cross-attn" % l,
                           "L%d co+ln" % l, "L%d f1" % l, "L%d f2+ln" % l] for l in range(6)], []) + ["logits+argmax"]
step3 = dts[2 * per_step: 3 * per_step]
print("position %d: step total %.1f us" % (pos + 2, sum(step3)))
agg = {}
for nm, v in zip(labels, step3):
    print("%-16s %6.2f" % (nm, v))
    k = nm.split(" ", 1)[1] if nm.startswith("L") else nm
    agg[k] = agg.get(k, 0.0) + v
print("--- per phase kind, summed over layers")
for k, v in agg.items():
    print("%-16s %7.2f us" % (k, v))
