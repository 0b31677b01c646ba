#!/bin/bash
# scratch script for one gpurun call (overwritten per call)
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "bwd or backward" > gpurun_out/bwd_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/bwd_tests.log
tail -3 gpurun_out/bwd_tests.log
timeout 300 python tools/train_time.py 512 bf16 5
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:attn_bwd_rows_kernel --launch-skip 55 --launch-count 1 -o gpurun_out/ncu_bwd_rows1b -f python tools/train_time.py 512 bf16 1 > gpurun_out/ncu_bwd_rows.log 2>&1
echo "ncu exit $?"
