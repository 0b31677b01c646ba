set -x
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "mamba or selective_scan" -s 2>&1 | tail -12
