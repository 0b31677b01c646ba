#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "regression or mamba or moe or variant" > gpurun_out/reg_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/reg_tests.log
tail -8 gpurun_out/reg_tests.log | cut -c1-300
