#!/bin/bash
cd /root/repo
timeout 1500 python -m pytest tests/test_gpu_train.py tests/test_gpu_dropout.py tests/test_gpu_kernels.py tests/test_gpu_variant_train.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r3g_tests.log
python tools/train_time.py 512 bf16 3 2>&1 | tail -1 | tee gpurun_out/r3g_train.txt
python tools/train_time.py 64 bf16 5 2>&1 | tail -1 | tee -a gpurun_out/r3g_train.txt
