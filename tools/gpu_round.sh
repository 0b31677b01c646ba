#!/bin/bash
# scratch script for one gpurun call (overwritten per call)
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r5_tests_final.log 2>&1
echo "tests exit $?" >> gpurun_out/r5_tests_final.log
tail -3 gpurun_out/r5_tests_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r5_smoke.log 2>&1
tail -2 gpurun_out/r5_smoke.log
timeout 900 python bench.py > gpurun_out/r5_bench.json 2> gpurun_out/r5_bench.err
echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r5_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['us_per_launch'])
print('train', {k:d['train'].get(k) for k in ('value','ms_per_step','tflops')}, d['train'].get('no_dropout'))
print('fp32', d['value_fp32']['value'])
for r in d.get('roofline_extra',[]): print(r.get('kernel','')[:60], r.get('ms'), r.get('frac'))
PY
