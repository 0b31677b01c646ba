"""fp32 exact-path generation, a few positions: for an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from video2music_b200 import synthetic as syn
dev = torch.device("cuda", 0)
model32, _ = bench.make_model(torch.float32, dev)
inp = syn.make_inputs(64, 1234, 299, 300, 0)
d = {k: v.to(dev) for k, v in inp.items()}
prim, pr, pa = torch.tensor([1]), torch.tensor([1]), torch.tensor([0])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
out = model32.generate(d["feature_semantic_list"], d["feature_key"], d["feature_scene_offset"], d["feature_motion"], d["feature_emotion"],
                       primer=prim, primer_root=pr, primer_attr=pa, target_seq_length=n, beam=1, beam_chance=1.0)
torch.cuda.synchronize()
print(out[0].tolist())
