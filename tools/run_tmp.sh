#!/bin/bash
cd /root/repo
python tools/scratch/step_once.py || exit 1
ncu --set full --clock-control none --import-source on -k regex:step_ --launch-skip 4 --launch-count 2 -o gpurun_out/r02_step_kernels python tools/scratch/step_once.py > gpurun_out/r3m_ncu.log 2>&1
python tools/ncu_key_metrics.py gpurun_out/r02_step_kernels.ncu-rep > gpurun_out/r02_step_kernels_ncu_key_metrics.txt 2>&1
cat gpurun_out/r02_step_kernels_ncu_key_metrics.txt | head -60
