#!/bin/bash
# scratch script for one gpurun call (overwritten per call)
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "moe or mamba or selective or variant or gemm" > gpurun_out/moe_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/moe_tests.log
tail -15 gpurun_out/moe_tests.log
timeout 120 python - > gpurun_out/moe_time.log 2>&1 <<'PY'
import torch
from video2music_b200 import GLUExpert, SharedMoELayer, MoELayer
torch.manual_seed(0)
for cls in (MoELayer, SharedMoELayer):
    m = cls(GLUExpert(512, 1024, 0.0), 512, n_experts=6, n_experts_per_token=2, dropout=0.0).eval().cuda()
    x = torch.randn(300, 64, 512, device="cuda")
    for dt in (torch.float32, torch.bfloat16):
        m.compute_dtype = dt
        with torch.no_grad():
            for _ in range(3): y = m(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): y = m(x)
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        ne = 2 + (cls is SharedMoELayer)
        print("%s 19200 tokens d512 ff1024 E6 k2 %s: %.3f ms/layer, %.1f TFLOP/s" % (cls.__name__, dt, ms, 19200 * ne * 3 * 2 * 512 * 1024 / ms / 1e9))
PY
cat gpurun_out/moe_time.log
timeout 300 python tools/prof_kernels.py 512 2>&1 | grep -i "scan\|attn" 
