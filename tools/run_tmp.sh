#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "step_ or small_batch" 2>&1 | tail -15 | tee gpurun_out/r2r_tests.log
