timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/r2_t.log
for b in 64 512; do python tools/train_time.py $b bf16 5; done > gpurun_out/r2_train_scaling_1gpu_d.txt 2>&1
