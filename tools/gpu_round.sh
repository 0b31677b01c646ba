set -x
timeout 500 python bench.py > gpurun_out/s14_bench.json 2> gpurun_out/s14_bench.err; tail -3 gpurun_out/s14_bench.err; python -c "
import json; d=json.load(open('gpurun_out/s14_bench.json')); print({k:d[k] for k in ('value','ms_per_step','share_of_step','gpu_launches')}); print(d['roofline']['frac']); print(d.get('train'))"
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
