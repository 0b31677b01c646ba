#!/bin/bash
# scratch script for one gpurun call (overwritten per call)
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "mamba or metrics or moe" > gpurun_out/new_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/new_tests.log
tail -6 gpurun_out/new_tests.log
