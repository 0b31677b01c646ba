#!/bin/bash
# scratch script for one gpurun call (overwritten per call)
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/gpu_tests.log
tail -4 gpurun_out/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches']); print(d['train']['value'], d['train']['ms_per_step'], d['train']['cuda_graph']); print(json.dumps(d.get('variants'))[:900])"
tail -3 gpurun_out/bench.err
