#!/bin/bash
cd /root/repo
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_amt.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r3b_tests.log
python tools/scratch/step_kernels_time.py 2>&1 | grep "tiled" | tee gpurun_out/r3b_times.txt
python tools/scratch/cfg5_train_step.py 2>&1 | tail -1 | tee -a gpurun_out/r3b_times.txt
python tools/scratch/cfg4_train_step.py 64 f32 2>&1 | tail -1 | tee -a gpurun_out/r3b_times.txt
