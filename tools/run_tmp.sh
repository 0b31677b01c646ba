#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_cached_decode.py -m gpu -x -q -k "step_ or small_batch or cached" 2>&1 | tail -4 | tee gpurun_out/r2s_tests.log
python tools/scratch/step_kernels_time.py 2>&1 | tail -9 | tee gpurun_out/r2s_step_kernels.txt
