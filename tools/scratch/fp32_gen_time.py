"""fp32 exact-path generation (64 videos x 299 tokens): ms per generation (bench.py's value_fp32 leg, stand-alone)."""
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
gen = lambda: model32.generate(d["feature_semantic_list"], d["feature_key"], d["feature_scene_offset"], d["feature_motion"], d["feature_emotion"],
                               primer=prim, primer_root=pr, primer_attr=pa, target_seq_length=300, beam=1, beam_chance=1.0)
gen(); torch.cuda.synchronize()
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = gen(); e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1)
    print("fp32 generate: %.1f ms, %.0f tok/s" % (ms, 64 * 299 / (ms * 1e-3)), flush=True)
