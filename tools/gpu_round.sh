set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s20_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s20_tests.log; tail -4 gpurun_out/s20_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 500 python bench.py > gpurun_out/s20_bench.json 2> gpurun_out/s20_bench.err; tail -2 gpurun_out/s20_bench.err; python -c "
import json; d=json.load(open('gpurun_out/s20_bench.json')); print({k:d[k] for k in ('value','ms_per_step','share_of_step','gpu_launches')}); print(d['e2e']['value'], d['roofline']['frac'], d['roofline']['us_per_launch']); print(d.get('train')); print(d.get('cpu_baseline'))"
timeout 300 python tools/prof_generate.py 300 299 0 bf16 stream > gpurun_out/s20_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_stream -c 1 -o gpurun_out/s20_stream python tools/prof_generate.py 300 299 0 bf16 stream > gpurun_out/s20_ncu.log 2>&1; tail -2 gpurun_out/s20_ncu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s20_generate_launches.csv python tools/prof_generate.py 300 299 0 bf16 stream > gpurun_out/s20_ncu1.log 2>&1; tail -1 gpurun_out/s20_ncu1.log
