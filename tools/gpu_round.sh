#!/bin/bash
# scratch script for one gpurun call (overwritten per call)
cd /root/repo
mkdir -p gpurun_out
L=video2music_b200/csrc/libv2m_b200.so
cp $L /tmp/lib_main.so
run() { echo "== $1"; timeout 200 python tools/gemm_bench.py 2>&1 | grep -v Warning; V2M_TRAIN_GRAPH=1 timeout 300 python tools/train_time.py 512 bf16 10 2>&1 | tail -1; }
( run "explicit sts (main)"
  cp tools/scratch/alt/libv2m_b200.so $L; run "generic staging stores in the TMA-store epilogue (alt)"
  cp /tmp/lib_main.so $L; run "explicit sts (main) again"
  cp tools/scratch/alt/libv2m_b200.so $L; run "generic (alt) again"
  cp /tmp/lib_main.so $L ) > gpurun_out/r5_gemm_sts_ab.txt 2>&1
cat gpurun_out/r5_gemm_sts_ab.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_rows_kernel -s 1 -c 1 -o gpurun_out/r5_attn_bwd_rows python tools/prof_attn_bwd_once.py rpr > gpurun_out/r5_ncu_bwd.log 2>&1
tail -2 gpurun_out/r5_ncu_bwd.log
