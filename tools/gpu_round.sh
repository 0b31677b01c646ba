#!/bin/bash
# scratch script for one gpurun call (overwritten per call)
cd /root/repo
mkdir -p gpurun_out
( timeout 200 python tools/prof_attn_fwd.py ) > gpurun_out/r5_attn_fwd_lean.txt 2>&1
cat gpurun_out/r5_attn_fwd_lean.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r5_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r5_tests.log
tail -4 gpurun_out/r5_tests.log
( V2M_TRAIN_GRAPH=1 timeout 300 python tools/train_time.py 512 bf16 10; V2M_GEMM_PAIR=0 V2M_TRAIN_GRAPH=1 timeout 300 python tools/train_time.py 512 bf16 10
  V2M_TRAIN_GRAPH=1 timeout 300 python tools/train_time.py 64 bf16 20 ) 2>&1 | grep "^train" > gpurun_out/r5_train_time3.txt
cat gpurun_out/r5_train_time3.txt
