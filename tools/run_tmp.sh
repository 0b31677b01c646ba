#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_variant_train.py -m gpu -q -s -k "variant_train_bf16" 2>&1 | grep -v "^    \|^$\|^E  \|^>" | tail -12 | tee gpurun_out/r2y_tests.log
