#!/bin/bash
# scratch script for one gpurun call (overwritten per call)
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "bwd or backward" > gpurun_out/bwd_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/bwd_tests.log
tail -12 gpurun_out/bwd_tests.log
timeout 300 python -m pytest tests/test_gpu_train.py -x -q -m gpu > gpurun_out/train_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/train_tests.log
tail -4 gpurun_out/train_tests.log
timeout 300 python tools/train_time.py 512 bf16 5
