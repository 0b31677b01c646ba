#!/bin/bash
# scratch script for one gpurun call (overwritten per call)
cd /root/repo
mkdir -p gpurun_out
( V2M_GEMM_PAIR=2 timeout 120 python tools/scratch/pair_check.py | grep -v " ok$" ) > gpurun_out/r5_pair_check.txt 2>&1
cat gpurun_out/r5_pair_check.txt | tail -4
if [ "$(grep -c 'ALL OK' gpurun_out/r5_pair_check.txt)" = "1" ]; then
  ( echo "== pair kernel, relaxed tmem-empty arrive (default)"; timeout 200 python tools/gemm_bench.py 2>&1 | grep -v Warning | grep 153600
    echo "== V2M_GEMM_PAIR=0"; V2M_GEMM_PAIR=0 timeout 200 python tools/gemm_bench.py 2>&1 | grep -v Warning | grep 153600 ) > gpurun_out/r5_pair_bench4.txt 2>&1
  cat gpurun_out/r5_pair_bench4.txt
  ( V2M_TRAIN_GRAPH=1 timeout 300 python tools/train_time.py 512 bf16 10; V2M_GEMM_PAIR=0 V2M_TRAIN_GRAPH=1 timeout 300 python tools/train_time.py 512 bf16 10 ) 2>&1 | grep "^train" > gpurun_out/r5_train_time6.txt
  cat gpurun_out/r5_train_time6.txt
fi
