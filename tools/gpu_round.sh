set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s2_tests.log
python bench.py > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err
python tools/probe_decode.py > gpurun_out/s2_probe.log 2>&1
python tools/train_time.py 64 bf16 5 > gpurun_out/s2_train.log 2>&1
python tools/train_time.py 64 fp32 3 >> gpurun_out/s2_train.log 2>&1
python tools/prof_generate.py 300 3 150 > gpurun_out/s2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s2_launches.csv python tools/prof_generate.py 300 3 150 > gpurun_out/s2_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dec_attn -s 6 -c 2 -o gpurun_out/s2_dec_attn python tools/prof_generate.py 300 2 150 > gpurun_out/s2_ncu2.log 2>&1
tail -3 gpurun_out/s2_tests.log; cat gpurun_out/s2_bench.json; cat gpurun_out/s2_probe.log gpurun_out/s2_train.log
