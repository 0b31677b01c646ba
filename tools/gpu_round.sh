#!/bin/bash
# scratch script for one gpurun call (overwritten per call)
cd /root/repo
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_variant_train.py -q -m gpu > gpurun_out/variant_train_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/variant_train_tests.log
tail -40 gpurun_out/variant_train_tests.log | cut -c1-600
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "mamba or regression or scan" > gpurun_out/regress_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/regress_tests.log
tail -5 gpurun_out/regress_tests.log | cut -c1-300
