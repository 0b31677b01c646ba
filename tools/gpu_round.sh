#!/bin/bash
# scratch script for one gpurun call (overwritten per call)
cd /root/repo
mkdir -p gpurun_out
( V2M_ATTN_SHIFT=0 timeout 200 python tools/prof_attn_fwd.py; timeout 200 python tools/prof_attn_fwd.py ) > gpurun_out/r5_attn_fwd_ab.txt 2>&1
cat gpurun_out/r5_attn_fwd_ab.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_amt.py tests/test_gpu_dropout.py tests/test_gpu_train.py -x -q -m gpu > gpurun_out/r5_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r5_tests.log
tail -4 gpurun_out/r5_tests.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bf16_tc_kernel -s 2 -c 1 -o gpurun_out/r5_attn_fwd python tools/prof_attn_fwd.py 512 1 > gpurun_out/r5_ncu.log 2>&1
tail -3 gpurun_out/r5_ncu.log
