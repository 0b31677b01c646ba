"""Per-phase timestamps of the cluster decode kernel (CTA 0 of cluster 0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from video2music_b200 import engine, synthetic as syn, _lib
dev = torch.device("cuda", 0)
model, _ = bench.make_model(torch.bfloat16, dev)
inp = syn.make_inputs(64, 1234, 299, 300, 0)
d = {k: v.to(dev) for k, v in inp.items()}
prim, pr, pa = torch.tensor([1]), torch.tensor([1]), torch.tensor([0])
st = engine.build_decode(model._w(), model._cfg(), d["feature_semantic_list"], d["feature_key"].reshape(-1),
                         d["feature_scene_offset"], d["feature_motion"], d["feature_emotion"], prim, pr, pa, 300)
step_t = st.step
if len(sys.argv) > 1:
    os.environ["V2M_CLUSTER"] = sys.argv[1]
ts = torch.zeros(4096, dtype=torch.int64, device=dev)
step_t.fill_(150); st.pos = 150
engine.run_decode(st, 2, mode="cluster")
torch.cuda.synchronize()
_lib.check(_lib.load().v2m_debug_set_timestamps(_lib.ptr(ts), 4096))
step_t.fill_(150); st.pos = 150
engine.run_decode(st, 4, mode="cluster")
torch.cuda.synchronize()
_lib.check(_lib.load().v2m_debug_set_timestamps(None, 0))
t = ts.cpu().tolist()
n = 1 + 4 * 52
d = [(t[i + 1] - t[i]) / 1000.0 for i in range(n - 1)]
names = ["embed"] + sum([["L%d qkv" % l, "L%d self" % l, "L%d so" % l, "L%d cq" % l, "L%d cross" % l, "L%d co" % l, "L%d f1" % l, "L%d f2" % l] for l in range(6)], []) + ["logits", "argmax"]
step3 = d[2 * 52: 3 * 52]
print("step total %.1f us" % sum(step3))
for nm, v in zip(names, step3):
    print("%-10s %6.2f" % (nm, v))
