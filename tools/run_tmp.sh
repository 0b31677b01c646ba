#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cached_decode.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2q_tests.log
cat gpurun_out/r2q_tests.log
timeout 600 python - <<'PY' 2>&1 | tail -5 | tee gpurun_out/r2q_gen.log
import torch, time, bench
from video2music_b200 import synthetic as syn, _lib
from video2music_b200.video_music_transformer_v2 import VideoMusicTransformer_GQA
dev = torch.device("cuda", 0)
torch.manual_seed(0)
gm = VideoMusicTransformer_GQA(n_layers=6, total_vf_dim=syn.vf_dim(0)).eval().to(dev)
gi = syn.make_inputs(64, 1234, 299, 300, 0)
feats = [gi[k].to(dev) for k in ("feature_semantic_list", "feature_key", "feature_scene_offset", "feature_motion", "feature_emotion")]
one = torch.tensor([1])
gen_fn = lambda n: gm.generate_cached(*feats, primer=one, primer_root=one, primer_attr=torch.tensor([0]), target_seq_length=n, beam=1, beam_chance=1.0)
for rep in range(3):
    n0 = _lib.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = gen_fn(300); e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1)
    print("call %d: %.1f ms per generation, %.1f tok/s, launches %d" % (rep, ms, 64 * 299 / (ms * 1e-3), _lib.launches() - n0), flush=True)
PY
