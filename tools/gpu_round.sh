set -x
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train.py -x -q -k "attention_bwd or train" 2>&1 | tail -2
timeout 100 python tools/train_time.py 64 bf16 5 2>&1 | tail -1
timeout 100 python tools/train_time.py 512 bf16 3 2>&1 | tail -1
