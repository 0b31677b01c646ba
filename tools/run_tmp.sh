#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_variant_train.py -m gpu -q -s -k "bf16_projections" > gpurun_out/r3l_tests_full.log 2>&1
grep -n "bf16 Mamba projections\|passed\|failed" gpurun_out/r3l_tests_full.log
