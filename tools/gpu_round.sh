#!/bin/bash
# scratch script for one gpurun call (overwritten per call): full validation = GPU tests, smoke, bench
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r4_gpu_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r4_gpu_tests.log
tail -4 gpurun_out/r4_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r4_smoke.log 2>&1
tail -2 gpurun_out/r4_smoke.log
timeout 900 python bench.py > gpurun_out/r4_bench.json 2> gpurun_out/r4_bench.err
echo "bench exit $?"; tail -3 gpurun_out/r4_bench.err
