#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/prof_kernels.py 512 2>&1 | grep -i "attn"
V2M_ATTN_NW=2 timeout 300 python tools/prof_kernels.py 512 2>&1 | grep -i "attn"
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention or attn or rpr or gqa or dropout" > gpurun_out/attn_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/attn_tests.log
tail -3 gpurun_out/attn_tests.log
