#!/bin/bash
# scratch script for one gpurun call (overwritten per call)
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_variant_train.py tests/test_gpu_train.py -q -m gpu > gpurun_out/variant_train_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/variant_train_tests.log
tail -30 gpurun_out/variant_train_tests.log | cut -c1-600
timeout 200 python tools/prof_train_variants.py > gpurun_out/prof_train_variants.log 2>&1
echo "prof exit $?"; tail -14 gpurun_out/prof_train_variants.log | cut -c1-200
