#!/bin/bash
# scratch script for one gpurun call (overwritten per call)
cd /root/repo
mkdir -p gpurun_out
timeout 60 python tools/prof_scan_bwd_once.py > gpurun_out/s4_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/s4_plain.log; exit 1; }
timeout 100 ncu --set full --clock-control none --import-source on -k regex:'selective_scan_bwd_kernel|selective_scan_bwd_carry|moe_grouped_dw_kernel|mamba_conv_silu_bwd' -s 5 -c 5 \
  -o gpurun_out/s4_variant_bwd -f python tools/prof_scan_bwd_once.py > gpurun_out/s4_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/s4_ncu.log; ls -la gpurun_out/s4_variant_bwd.ncu-rep
