set -x
timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -k "attention_bwd_tensor_core" -s 2>&1 | tail -25
timeout 200 python -m pytest tests/test_gpu_train.py -x -q -s 2>&1 | tail -15
timeout 100 python tools/train_time.py 64 bf16 5 2>&1 | tail -2
