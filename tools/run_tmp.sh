#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_amt.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2m_tests.log
cat gpurun_out/r2m_tests.log
timeout 600 python tools/scratch/fp32_gen_time.py 2>&1 | tail -3 | tee gpurun_out/r2m_fp32.log
