for b in 64 128 256 512; do python tools/train_time.py $b bf16 5; done > gpurun_out/r2_train_scaling_1gpu.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_train_b64_launches.csv python tools/train_time.py 64 bf16 1 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2_train_b64_launches.csv > gpurun_out/r2_train_b64_summary.txt 2>&1
