#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_variant_train.py -m gpu -q -k "zoo_model" > gpurun_out/r3n_tests_full.log 2>&1
grep -v "^$" gpurun_out/r3n_tests_full.log | tail -40
