"""Live timings (CUDA events, warm caches, 200 launches each inside one CUDA graph) of the generation-step kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from video2music_b200 import ops
dev = torch.device("cuda", 0)
torch.manual_seed(0)
x = torch.randn(64, 512, device=dev); x2 = torch.randn(64, 1024, device=dev)
w = torch.randn(512, 512, device=dev) * 0.05; w3 = torch.randn(1536, 512, device=dev) * 0.05; w2 = torch.randn(512, 1024, device=dev) * 0.05
b = torch.randn(512, device=dev); b3 = torch.randn(1536, device=dev)
K = torch.randn(64, 300, 128, device=dev); V = torch.randn(64, 300, 128, device=dev); q = torch.randn(64, 512, device=dev)
g = torch.ones(512, device=dev); bb = torch.zeros(512, device=dev)
nd = torch.tensor([300], dtype=torch.int32, device=dev)

def timed(name, fn, reps=200):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    gr = torch.cuda.CUDAGraph()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(gr, stream=s):
            for _ in range(reps):
                fn()
    torch.cuda.current_stream().wait_stream(s)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); e1.synchronize()
    print("%-52s %7.2f us per launch" % (name, e0.elapsed_time(e1) * 1000 / reps), flush=True)

timed("step_linear 64 x 512 x 512", lambda: ops.step_linear(x, w, b))
timed("step_linear 64 x 1536 x 512", lambda: ops.step_linear(x, w3, b3))
timed("step_linear 64 x 512 x 1024", lambda: ops.step_linear(x2, w2, b))
timed("linear (tiled fp32 GEMM) 64 x 512 x 512", lambda: ops.linear(x, w, b))
timed("step_attention 64 videos x 8 heads (2 kv), n = 300", lambda: ops.step_attention(q, K, V, Hq=8, Hkv=2, dh=64, n_max=300, kv_strides=(K.stride(0), K.stride(1)), n_dev=nd, q_scale=0.125))
timed("layernorm 64 x 512", lambda: ops.layernorm(x, g, bb))
timed("torch add 64 x 512", lambda: x + x)
