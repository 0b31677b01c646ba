#!/bin/bash
# scratch script for one gpurun call (overwritten per call)
cd /root/repo
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r5_bench_n8.json 2> gpurun_out/r5_bench_n8.err
echo "bench n8 exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r5_bench_n8.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'])
print('train', {k:d['train'].get(k) for k in ('value','ms_per_step','tflops')})
PY
