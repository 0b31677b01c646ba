#!/bin/bash
# scratch script for one gpurun call (overwritten per call)
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_variant_train.py -x -q -m gpu -k "pscan or selective_scan" > gpurun_out/r4_edge_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r4_edge_tests.log
tail -15 gpurun_out/r4_edge_tests.log
