set -x
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train.py -x -q 2>&1 | tail -5
timeout 100 python tools/train_time.py 64 bf16 5 2>&1 | tail -1
timeout 100 python tools/train_time.py 64 fp32 3 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/s13_train_launches.csv python tools/train_time.py 64 bf16 1 > gpurun_out/s13_ncu.log 2>&1
