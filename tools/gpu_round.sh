#!/bin/bash
# scratch script for one gpurun call (overwritten per call)
cd /root/repo
mkdir -p gpurun_out
( V2M_GEMM_PAIR=2 timeout 120 python tools/scratch/pair_check.py | grep -v " ok$"; V2M_GEMM_PAIR=0 timeout 120 python tools/scratch/pair_check.py | grep -v " ok$" ) > gpurun_out/r5_pair_check.txt 2>&1
cat gpurun_out/r5_pair_check.txt | tail -6
( timeout 300 python tools/prof_kernels.py 512 2>&1 | grep "attn_bwd\|gemm" ) > gpurun_out/r5_kernels_lean2.txt 2>&1
cat gpurun_out/r5_kernels_lean2.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train.py tests/test_gpu_dropout.py tests/test_gpu_variant_train.py -x -q -m gpu > gpurun_out/r5_tests_b.log 2>&1
echo "tests exit $?" >> gpurun_out/r5_tests_b.log
tail -3 gpurun_out/r5_tests_b.log
( V2M_TRAIN_GRAPH=1 timeout 300 python tools/train_time.py 512 bf16 10; V2M_TRAIN_GRAPH=1 timeout 300 python tools/train_time.py 64 bf16 20 ) 2>&1 | grep "^train" > gpurun_out/r5_train_time5.txt
cat gpurun_out/r5_train_time5.txt
