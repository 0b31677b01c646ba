#!/bin/bash
# scratch script for one gpurun call (overwritten per call)
cd /root/repo
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
cap() {  # name regex skip [batch]
  timeout 400 $NCU -k regex:$2 --launch-skip $3 --launch-count 1 -o gpurun_out/ncu2_$1 -f python tools/prof_kernels.py ${4:-512} > gpurun_out/ncu2_$1.log 2>&1
  echo "ncu $1 exit $?"
}
cap gemm_inproj gemm_bf16_tc_kernel 2
cap attn_self attn_bf16_tc_kernel 2
cap attn_cross attn_bf16_tc_kernel 9
cap selscan "selective_scan_fwd_kernel" 7 64
timeout 600 $NCU -k regex:attn_bwd_rows_kernel --launch-skip 54 --launch-count 2 -o gpurun_out/ncu2_bwd_rows -f python tools/train_time.py 512 bf16 1 > gpurun_out/ncu2_bwd_rows.log 2>&1
echo "ncu rows exit $?"
timeout 300 python tools/prof_kernels.py 512 > gpurun_out/kernels_live3.txt 2>&1
cat gpurun_out/kernels_live3.txt
