#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_amt.py -m gpu -x -q -k "stream" 2>&1 | tail -3 | tee gpurun_out/r3o_tests.log
timeout 300 python tools/stream_exp.py 64 100 100 2>&1 | tail -1 | tee gpurun_out/r3o_exp.log
timeout 300 python tools/stream_exp.py 64 250 49 2>&1 | tail -1 | tee -a gpurun_out/r3o_exp.log
