#!/bin/bash
# scratch script for one gpurun call (overwritten per call)
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "moe" > gpurun_out/moe_tests.log 2>&1
echo "moe tests exit $?" >> gpurun_out/moe_tests.log
timeout 120 python - > gpurun_out/moe_time.log 2>&1 <<'PY'
import torch, time
from video2music_b200 import GLUExpert, SharedMoELayer, _lib
torch.manual_seed(0)
m = SharedMoELayer(GLUExpert(512, 1024, 0.0), 512, n_experts=6, n_experts_per_token=2, dropout=0.0).eval().cuda()
x = torch.randn(300, 16, 512, device="cuda")
with torch.no_grad():
    for _ in range(3): y = m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): y = m(x)
    e1.record(); torch.cuda.synchronize()
print("SharedMoE 4800 tokens d512 ff1024 E6 k2: %.3f ms/layer" % (e0.elapsed_time(e1) / 20))
fl = 4800 * (2 + 1) * 3 * 2 * 512 * 1024
print("%.1f TFLOP/s fp32" % (fl / (e0.elapsed_time(e1) / 20 * 1e-3) / 1e12))
PY
tail -5 gpurun_out/moe_tests.log; cat gpurun_out/moe_time.log
