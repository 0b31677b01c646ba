set -x
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train.py tests/test_gpu_amt.py -x -q 2>&1 | tail -3
timeout 100 python tools/train_time.py 64 bf16 5 2>&1 | tail -1
timeout 100 python tools/train_time.py 512 bf16 3 2>&1 | tail -1
timeout 100 python tools/probe_decode.py > gpurun_out/s22_probe.log 2>&1; grep "bfloat16 decode step mode=stream\|bfloat16 decode step mode=kernels graph=True split=2" gpurun_out/s22_probe.log
