#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_amt.py -x -q -m gpu -k "v2_model" > gpurun_out/v2_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/v2_tests.log
tail -25 gpurun_out/v2_tests.log | cut -c1-300
