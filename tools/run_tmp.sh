#!/bin/bash
cd /root/repo
timeout 1500 python -m pytest tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r3j_tests.log
python tools/train_time.py 512 bf16 3 2>&1 | tail -1 | tee gpurun_out/r3j_train.txt
python tools/train_time.py 64 bf16 5 2>&1 | tail -1 | tee -a gpurun_out/r3j_train.txt
