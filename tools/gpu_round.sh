set -x
timeout 400 python bench.py > gpurun_out/s10_bench.json 2> gpurun_out/s10_bench.err; tail -3 gpurun_out/s10_bench.err; python -c "
import json; d=json.load(open('gpurun_out/s10_bench.json')); print({k:d[k] for k in ('value','ms_per_step','share_of_step','gpu_launches')}); print(d['e2e']); print(d['roofline']); print(d.get('train')); print(d.get('cpu_baseline')); print(d['clocks'])"
timeout 300 python tools/prof_generate.py 300 299 0 bf16 stream > gpurun_out/s10_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_stream -c 1 -o gpurun_out/s10_stream python tools/prof_generate.py 300 299 0 bf16 stream > gpurun_out/s10_ncu.log 2>&1; tail -3 gpurun_out/s10_ncu.log
