#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "custom_mha or variant" > gpurun_out/v2_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/v2_tests.log
tail -12 gpurun_out/v2_tests.log | cut -c1-300
