#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_amt.py -m gpu -x -q -k "stream" 2>&1 | tail -5 > gpurun_out/r2b_tests.log
cat gpurun_out/r2b_tests.log
timeout 300 python tools/stream_exp.py 64 100 100 2>&1 | tail -1 | tee gpurun_out/r2b_exp.log
timeout 300 python tools/stream_exp.py 5 100 100 2>&1 | tail -1 | tee -a gpurun_out/r2b_exp.log
timeout 300 python tools/stream_exp.py 64 250 49 2>&1 | tail -1 | tee -a gpurun_out/r2b_exp.log
