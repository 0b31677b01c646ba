"""Training-step timing (fwd + loss + bwd + Adam) on one GPU: python tools/train_time.py [batch] [dtype] [steps]."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video2music_b200 import VideoMusicTransformer, synthetic as syn, _lib
from video2music_b200.trainer import Trainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dt = torch.bfloat16 if (len(sys.argv) <= 2 or sys.argv[2] == "bf16") else torch.float32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda", 0)
m = VideoMusicTransformer(total_vf_dim=syn.vf_dim(0), rpr=True, dropout=0.0)
shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
m.load_state_dict(syn.fill_like_reference_init(shapes, seed=1), strict=False)
m = m.to(dev).train().set_compute_dtype(dt)
tr = Trainer(m, use_graph=os.environ.get("V2M_TRAIN_GRAPH") == "1")
inp = syn.make_inputs(B, 1234, 299, 300, 0)
b = {k: v.to(dev) for k, v in inp.items()}
for _ in range(3):
    loss = tr.train_step(b)
torch.cuda.synchronize()
_lib.reset_launches()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(steps):
    loss = tr.train_step(b)
e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1) / steps
print("train B=%d %s: %.2f ms/step (wall %.2f), %.1f samples/s, loss %.4f, launches/step %d, %.1f TF/s" % (
    B, dt, ms, (time.perf_counter() - t0) * 1e3 / steps, B / ms * 1e3, float(loss), _lib.launches() // steps, 69.4e9 * B / ms / 1e9))
