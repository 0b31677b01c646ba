#!/bin/bash
cd /root/repo
make -C video2music_b200/csrc clean > /dev/null; make -j8 -C video2music_b200/csrc EXTRA=-DV2M_EXP_NOATTN > /dev/null 2>&1 || { echo build failed; exit 1; }
for R in 5 8; do
  V2M_STREAM_ROWS=$R python tools/stream_exp.py 64 100 100 2>&1 | tail -1
done | tee gpurun_out/r3d_noattn.txt
V2M_STREAM_ROWS=8 python tools/stream_exp.py 8 100 100 2>&1 | tail -1 | tee -a gpurun_out/r3d_noattn.txt
V2M_STREAM_ROWS=5 python tools/stream_exp.py 5 100 100 2>&1 | tail -1 | tee -a gpurun_out/r3d_noattn.txt
