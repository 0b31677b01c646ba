set -x
timeout 400 python -m pytest tests/test_gpu_amt.py -x -q 2>&1 | tail -5
