set -x
for s in 0 1000 2800 5600 17000; do V2M_STREAM_STAGGER_NS=$s timeout 100 python tools/probe_decode.py 2>&1 | grep "bfloat16 decode step mode=stream" | sed "s/^/stagger $s: /"; done
