#!/bin/bash
# scratch script for one gpurun call (overwritten per call)
cd /root/repo
mkdir -p gpurun_out
export V2M_TRAIN_GRAPH=1
( V2M_NO_PDL=1 timeout 300 python tools/train_time.py 64 bf16 20; timeout 300 python tools/train_time.py 64 bf16 20
  V2M_NO_PDL=1 timeout 300 python tools/train_time.py 64 bf16 20; timeout 300 python tools/train_time.py 64 bf16 20
  V2M_NO_PDL=1 timeout 300 python tools/train_time.py 512 bf16 5; timeout 300 python tools/train_time.py 512 bf16 5 ) > gpurun_out/r4_train_time_pdl_graph.txt 2>&1
cat gpurun_out/r4_train_time_pdl_graph.txt
