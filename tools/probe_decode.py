"""Per-kernel decode timings (back-to-back launches in a stream, CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from video2music_b200 import engine, synthetic as syn
dev = torch.device("cuda", 0)
for dt in (torch.bfloat16, torch.float32):
    model, _ = bench.make_model(dt, dev)
    inp = syn.make_inputs(64, 1234, 299, 300, 0)
    d = {k: v.to(dev) for k, v in inp.items()}
    prim, pr, pa = torch.tensor([1]), torch.tensor([1]), torch.tensor([0])
    def build(mode):
        return engine.build_decode(model._w(), model._cfg(), d["feature_semantic_list"], d["feature_key"].reshape(-1),
                                   d["feature_scene_offset"], d["feature_motion"], d["feature_emotion"], prim, pr, pa, 300, mode=mode)
    st = build("kernels")
    step_t = st.step
    step_t.fill_(150)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for kind, name in ((0, "self-attn t=150"), (1, "cross-attn"), (2, "QKV gemm"), (3, "FFN1 gemm")):
        engine.probe_decode_kernel(st, kind, 3)
        torch.cuda.synchronize()
        flush.fill_(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); engine.probe_decode_kernel(st, kind, 20); e1.record(); e1.synchronize()
        print(dt, name, "%.2f us/launch" % (e0.elapsed_time(e1) * 1e3 / 120))
    modes = [("kernels", False, 1), ("kernels", True, 1), ("kernels", True, 2), ("kernels", True, 4), ("kernels", True, 8)]
    modes += ([("stream", False, 1)] if dt == torch.bfloat16 else [])
    for mode, graph, split in modes:
        if mode != st.mode:
            st = build(mode)
            step_t = st.step
        step_t.fill_(140); st.pos = 140
        engine.run_decode(st, 5, use_graph=graph, mode=mode, n_split=split)
        torch.cuda.synchronize()
        step_t.fill_(140); st.pos = 140
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); engine.run_decode(st, 20, use_graph=graph, mode=mode, n_split=split); e1.record(); e1.synchronize()
        print(dt, "decode step mode=%s graph=%s split=%d: %.1f us/step" % (mode, graph, split, e0.elapsed_time(e1) * 1e3 / 20))
