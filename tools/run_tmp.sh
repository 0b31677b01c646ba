#!/bin/bash
cd /root/repo
make -C video2music_b200/csrc clean > /dev/null; make -j8 -C video2music_b200/csrc EXTRA=-DV2M_CHUNK_STAMPS > /dev/null 2>&1 || exit 1
python tools/chunk_times.py 5 2>&1 | tail -2 > gpurun_out/r2o_chunks_b5.txt
python tools/chunk_times.py 64 2>&1 | tail -2 > gpurun_out/r2o_chunks_b64.txt
cat gpurun_out/r2o_chunks_b5.txt gpurun_out/r2o_chunks_b64.txt
