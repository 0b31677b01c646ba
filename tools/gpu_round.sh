#!/bin/bash
# scratch script for one gpurun call (overwritten per call)
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "pscan or scan" > gpurun_out/r4_pscan_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r4_pscan_tests.log
tail -5 gpurun_out/r4_pscan_tests.log
( timeout 120 python tools/scratch/pscan_ab.py
  V2M_PSCAN_CPI=32 timeout 120 python tools/scratch/pscan_ab.py
  V2M_PSCAN_CPI=64 timeout 120 python tools/scratch/pscan_ab.py
  V2M_PSCAN_CPI=32 V2M_PSCAN_LC=64 timeout 120 python tools/scratch/pscan_ab.py ) > gpurun_out/r4_pscan_ab3.txt 2>&1
cat gpurun_out/r4_pscan_ab3.txt
