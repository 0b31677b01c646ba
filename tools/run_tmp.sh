#!/bin/bash
cd /root/repo
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_cached_decode.py -m gpu -x -q -k "step_ or small_batch or correspondence or cached_generate_gqa" > gpurun_out/r3p_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r3p_memcheck.log
grep -c "Invalid\|out of bounds" gpurun_out/r3p_memcheck.log; tail -6 gpurun_out/r3p_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "step_attention" > gpurun_out/r3p_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/r3p_racecheck.log
tail -4 gpurun_out/r3p_racecheck.log
