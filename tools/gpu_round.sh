#!/bin/bash
# One gpurun call: full validation (GPU tests, smoke, bench).  Edited per call during development; this is the end-of-round form.
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/tests_final.log 2>&1
echo "tests exit $?" >> gpurun_out/tests_final.log
tail -3 gpurun_out/tests_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"
