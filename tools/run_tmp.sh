python -m pytest tests/test_gpu_variant_train.py -m gpu -q -s -k "bf16 or kgrouped" 2>&1 | tail -25 > gpurun_out/r2_moe16_tests.log
python tools/prof_train_variants.py 2>&1 | head -8 > gpurun_out/r2_moe16_timing.txt
