set -x
timeout 500 python bench.py --no-cpu-baseline > gpurun_out/s28_bench.json 2> gpurun_out/s28_bench.err; tail -2 gpurun_out/s28_bench.err; python -c "
import json; d=json.load(open('gpurun_out/s28_bench.json')); print({k:d[k] for k in ('value','ms_per_step','share_of_step','gpu_launches')}); print(d.get('train'))"
timeout 200 python -m pytest tests/test_gpu_train.py -x -q 2>&1 | tail -2
