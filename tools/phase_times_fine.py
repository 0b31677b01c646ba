"""Fine-grained phase timestamps of the streamed decode kernel (library built with EXTRA=-DV2M_FINE_STAMPS)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from video2music_b200 import engine, synthetic as syn
pos = int(sys.argv[1]) if len(sys.argv) > 1 else 150
if len(sys.argv) > 2:
    os.environ["V2M_STREAM_ROWS"] = sys.argv[2]
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
layer = (["qk mma", "v mma", "kv write"] + ["self: Er pass", "self: chunks", "self: combine+gather", "self: -"] +
         ["so mma", "so gather", "so LN"] + ["cq mma", "cq -"] +
         ["cross: chunks", "cross: combine+gather", "cross: -"] +
         ["co mma", "co gather", "co LN"] + ["f1 mma", "f1 gather", "f1 -"] +
         ["f2 mma", "f2 gather", "f2 LN"])
labels = ["embed mma", "embed gather", "embed conv"] + sum([["L%d %s" % (l, x) for x in layer] for l in range(6)], []) + \
         ["logits mma", "logits barrier", "argmax"]
per_step = len(labels)
ts = torch.zeros(1 + 3 * per_step, dtype=torch.int64, device=dev)
st.step.fill_(pos); st.pos = pos
engine.run_decode(st, 3, mode="stream", timestamps=ts)
torch.cuda.synchronize()
t = ts.cpu().tolist()
dts = [(t[i + 1] - t[i]) / 1000.0 for i in range(len(t) - 1)]
step = dts[per_step: 2 * per_step]
print("position %d: step total %.1f us (%d stamps)" % (pos + 1, sum(step), per_step))
agg = {}
for nm, v in zip(labels, step):
    if nm.startswith("L2 ") or not nm.startswith("L"):
        print("%-28s %6.2f" % (nm, v))
    k = nm.split(" ", 1)[1] if nm.startswith("L") else nm
    agg[k] = agg.get(k, 0.0) + v
print("--- summed over layers")
for k, v in agg.items():
    print("%-28s %7.2f us" % (k, v))
