#!/bin/bash
# scratch script for one gpurun call (overwritten per call)
cd /root/repo
mkdir -p gpurun_out
( V2M_GEMM_PAIR=2 timeout 120 python tools/scratch/pair_check.py | grep -v " ok$"; echo "pair=2 exit $?"
  V2M_GEMM_PAIR=0 timeout 120 python tools/scratch/pair_check.py | grep -v " ok$"; echo "pair=0 exit $?" ) > gpurun_out/r5_pair_check.txt 2>&1
cat gpurun_out/r5_pair_check.txt | tail -20
if [ "$(grep -c 'ALL OK' gpurun_out/r5_pair_check.txt)" = "2" ]; then
  ( echo "== lean issue loop, pair kernel on (default)"; timeout 200 python tools/gemm_bench.py 2>&1 | grep -v Warning
    echo "== lean issue loop, V2M_GEMM_PAIR=0"; V2M_GEMM_PAIR=0 timeout 200 python tools/gemm_bench.py 2>&1 | grep -v Warning ) > gpurun_out/r5_pair_bench2.txt 2>&1
  cat gpurun_out/r5_pair_bench2.txt
  ( V2M_TRAIN_GRAPH=1 timeout 300 python tools/train_time.py 512 bf16 10; V2M_GEMM_PAIR=0 V2M_TRAIN_GRAPH=1 timeout 300 python tools/train_time.py 512 bf16 10
    V2M_TRAIN_GRAPH=1 timeout 300 python tools/train_time.py 64 bf16 20 ) 2>&1 | grep "^train" > gpurun_out/r5_train_time2.txt
  cat gpurun_out/r5_train_time2.txt
fi
