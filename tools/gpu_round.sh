#!/bin/bash
# scratch script for one gpurun call (overwritten per call)
cd /root/repo
mkdir -p gpurun_out
export V2M_TRAIN_GRAPH=1
( timeout 200 python tools/train_time.py 64 bf16 20; timeout 200 python tools/train_time.py 64 bf16 20
  timeout 200 python tools/train_time.py 512 bf16 5 ) > gpurun_out/r4_train_time_ln.txt 2>&1
cat gpurun_out/r4_train_time_ln.txt
timeout 1200 python -m pytest tests/test_gpu_train.py tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/r4_ln_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r4_ln_tests.log
tail -4 gpurun_out/r4_ln_tests.log
