#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -x -q -m gpu -s -k graphed > gpurun_out/train_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/train_tests.log
tail -25 gpurun_out/train_tests.log | cut -c1-400
