for b in 64 512 64 512; do python tools/train_time.py $b bf16 5; done > gpurun_out/r2_train_scaling_1gpu_e.txt 2>&1
