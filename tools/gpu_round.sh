#!/bin/bash
# scratch script for one gpurun call (overwritten per call)
cd /root/repo
mkdir -p gpurun_out
( V2M_GEMM_ROWS=0 timeout 200 python tools/scratch/cfg5_train_step.py bf16; timeout 200 python tools/scratch/cfg5_train_step.py bf16
  V2M_GEMM_ROWS=0 timeout 200 python tools/scratch/cfg5_train_step.py; timeout 200 python tools/scratch/cfg5_train_step.py ) > gpurun_out/r4_cfg5_rows_ab3.txt 2>&1
cat gpurun_out/r4_cfg5_rows_ab3.txt
timeout 900 python -m pytest tests/test_gpu_variant_train.py tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/r4_rows_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r4_rows_tests.log
tail -4 gpurun_out/r4_rows_tests.log
