set -x
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/s8_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s8_tests.log; tail -4 gpurun_out/s8_tests.log
timeout 100 python tools/probe_decode.py > gpurun_out/s8_probe.log 2>&1; grep "bfloat16 decode step mode=stream\|split=2" gpurun_out/s8_probe.log
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/s8_bench.json 2> gpurun_out/s8_bench.err; cat gpurun_out/s8_bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'])"; tail -3 gpurun_out/s8_bench.err
