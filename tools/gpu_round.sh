#!/bin/bash
# scratch script for one gpurun call (overwritten per call)
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm or linear or moe" > gpurun_out/gemm_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/gemm_tests.log
tail -3 gpurun_out/gemm_tests.log
timeout 300 python tools/prof_kernels.py 512 2>&1 | grep gemm
timeout 300 python tools/train_time.py 512 bf16 5
timeout 600 python -m pytest tests/test_gpu_amt.py tests/test_gpu_train.py -x -q -m gpu > gpurun_out/amt_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/amt_tests.log
tail -3 gpurun_out/amt_tests.log
