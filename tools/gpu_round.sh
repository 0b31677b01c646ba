#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_amt.py -x -q -m gpu > gpurun_out/train_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/train_tests.log
tail -4 gpurun_out/train_tests.log | cut -c1-300
timeout 300 python tools/train_time.py 512 bf16 5
timeout 300 python tools/train_time.py 64 bf16 10
