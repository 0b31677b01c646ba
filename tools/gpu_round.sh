#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train.py -x -q -m gpu -k "dropout or bwd or backward or train or attention" > gpurun_out/drop_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/drop_tests.log
tail -5 gpurun_out/drop_tests.log | cut -c1-300
timeout 300 python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
from video2music_b200 import VideoMusicTransformer, synthetic as syn
from video2music_b200.trainer import Trainer
for p in (0.0, 0.2):
    m = VideoMusicTransformer(total_vf_dim=syn.vf_dim(0), rpr=True, dropout=p)
    m.load_state_dict(syn.fill_like_reference_init({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=1), strict=False)
    m = m.cuda().train().set_compute_dtype(torch.bfloat16)
    tr = Trainer(m)
    b = {k: v.cuda() for k, v in syn.make_inputs(512, 1234, 299, 300, 0).items()}
    for _ in range(3): tr.train_step(b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): loss = tr.train_step(b)
    e1.record(); e1.synchronize()
    print("train B=512 bf16 dropout=%.1f: %.2f ms/step, loss %.4f" % (p, e0.elapsed_time(e1) / 5, float(loss)))
    del tr, m
PY
