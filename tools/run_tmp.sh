timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention_bwd_tensor_core or probability_dropout" 2>&1 | tail -3 > gpurun_out/r2_tc5_tests.log
V2M_TC5_NW=4 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention_bwd_tensor_core or probability_dropout" 2>&1 | tail -3 >> gpurun_out/r2_tc5_tests.log
python tools/prof_kernels.py 512 2>&1 | grep -E "attn_bwd cross" > gpurun_out/r2_tc5_kernels.txt
V2M_TC5_NW=4 python tools/prof_kernels.py 512 2>&1 | grep -E "attn_bwd cross" >> gpurun_out/r2_tc5_kernels.txt
