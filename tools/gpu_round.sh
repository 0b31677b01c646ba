set -x
timeout 40 python tools/stream_debug.py 13 140 3 5 > gpurun_out/s15_dbg.log 2>&1; grep -v "^[0-9]* \[" gpurun_out/s15_dbg.log | cut -c1-160 | tail -4
timeout 150 python -m pytest tests/test_gpu_amt.py -x -q -k "stream" 2>&1 | tail -3
timeout 100 python tools/probe_decode.py > gpurun_out/s15_probe.log 2>&1; grep "bfloat16 decode step mode=stream" gpurun_out/s15_probe.log
