#!/bin/bash
# scratch script for one gpurun call (overwritten per call)
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_variant_train.py tests/test_gpu_kernels.py -x -q -m gpu -k "mamba or scan or regression or Mamba" > gpurun_out/r4_scan_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r4_scan_tests.log
tail -5 gpurun_out/r4_scan_tests.log
timeout 120 python tools/scratch/scan_bwd_ab.py > gpurun_out/r4_scan_bwd_ab3.txt 2>&1
cat gpurun_out/r4_scan_bwd_ab3.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r4_scan_launches2.csv python tools/scratch/scan_bwd_once.py > gpurun_out/r4_scan_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r4_scan_launches2.csv 2>/dev/null | head -14
