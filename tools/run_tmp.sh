ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_train_b64_launches_b.csv python tools/train_time.py 64 bf16 1 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2_train_b64_launches_b.csv > gpurun_out/r2_train_b64_summary_b.txt 2>&1
