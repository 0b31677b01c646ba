#!/usr/bin/env python
"""Benchmark of the AMT hot path (BASELINE.json): KV-cached greedy chord generation.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype bf16|fp32]

One "step" = one full generation pass over a batch of 64 synthetic videos per GPU (BASELINE config 2:
encoder over 300 video seconds, cross-attention K/V caches, 299 greedy decode steps with a KV cache).
Prints ONE JSON line (rank 0).  `value` is whole-job chord-tokens/s with the inputs resident in HBM,
`e2e` is the same metric through the public API (`VideoMusicTransformer.generate`) from pinned host
buffers (H2D of the features and D2H of the tokens inside the timed region).

`--impl reference` times the reference's own algorithm -- batch 1, no KV cache, the whole model re-run
on the growing prefix every step (video_music_transformer.py:1046-1084) -- on the host cores, using the
CPU restatement in oracle/ (the reference is pure Python and /root/reference does not exist on the
GPU box; the restatement is pinned to it by tests/golden).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 64                 # videos per GPU (BASELINE config 2)
TRAIN_GLOBAL_BATCH = 512   # BASELINE config 3 (split over the ranks: strong scaling)
TRAIN_DROPOUT = 0.2        # train.py's default (utilities/argument_funcs.py:14,52); every dropout site runs fused in the bf16 kernels
WOUT_GAIN = 4.0            # synthetic weights: Wout scaled x4 so that the greedy arg-max is not decided by rounding noise
SEQ = 300                  # target_seq_length -> 299 generated chord tokens per video
METRIC = "generate_chord_tokens_per_s"
UNIT = "tokens/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def tensor_peaks():
    """(burst, sustained) dense bf16 TFLOP/s: a kernel timed alone is held against the burst figure, a kernel inside a
    long step against the sustained one."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["bf16_tflops"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured (MEASURED_PEAKS.json)"
    return 1700.0, 1700.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel_substr`, parsed from the NEWEST ncu raw-page export
    under profiles/ (`ncu -i X.ncu-rep --page raw --csv`, kept as profiles/rNN_*_ncu_raw.csv; one `--set full` capture).
    Returns (bytes or None, file name or None)."""
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_raw.csv")), reverse=True)
    for f in files:
        try:
            rows = list(csv.reader(open(f, newline="")))
        except OSError:
            continue
        hdr = next((r for r in rows if "Kernel Name" in r), None)
        if hdr is None or "dram__bytes_read.sum" not in hdr:
            continue
        units = rows[rows.index(hdr) + 1]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        kn, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        for r in rows[rows.index(hdr) + 2:]:
            if len(r) > max(kn, rd, wr) and kernel_substr in r[kn]:
                try:
                    return int(float(r[rd].replace(",", "")) * scale.get(units[rd], 1.0) +
                               float(r[wr].replace(",", "")) * scale.get(units[wr], 1.0)), os.path.basename(f)
                except ValueError:
                    continue
    return None, None


def bench_config():
    """The workload both arms are timed on (BASELINE config 2)."""
    return {"workload": "AMT greedy chord generation with KV cache: 64 videos/GPU x 300 positions (299 decoded tokens), "
                        "6+6 layers, d_model 512, 8 heads, RPR, vf 776, primer length 1, chord_embed",
            "videos_per_gpu": BATCH, "target_seq_length": SEQ, "sharding": "independent videos per rank, no collective",
            "weights": "seeded random init at the reference constructor's scales, Wout x %g (wout_gain) so that the greedy "
                       "arg-max is not decided by rounding noise" % WOUT_GAIN,
            "l2": "256 MiB written between timed steps (L2 flush); per-step working set 386 MB > L2"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def algorithmic_bytes(B, S, T, E, H, FF, NL, vocab, esz):
    """bytes one whole decode run must touch (each weight once per step, each cached K/V element
    once per step)."""
    w = NL * (3 * E * E + E * E + E * E + E * E + 2 * E * FF + 300 * (E // H)) + E * vocab + (E + 1) * E
    per_step_w = w * esz
    cross = NL * 2 * S * E * B * esz
    steps = T - 1
    self_kv = sum(NL * 2 * (t + 1) * E * B * esz for t in range(steps))
    return steps * (per_step_w + cross) + self_kv, NL * 0 + 2 * S * E * B * esz   # (whole run, one cross-attention launch)


def make_model(dtype, device, seed=1):
    from video2music_b200 import VideoMusicTransformer
    from video2music_b200 import synthetic as syn
    m = VideoMusicTransformer(total_vf_dim=syn.vf_dim(0), rpr=True, chord_embed=True)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(syn.fill_like_reference_init(shapes, seed=seed, wout_gain=WOUT_GAIN), strict=False)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    if device is not None:
        m = m.to(device)
    return m.eval().set_compute_dtype(dtype), sd


def train_leg(args, dev, rank, world, dtype):
    """BASELINE config 3: training step (forward + loss + backward + all-reduce + Adam), global batch 512 split over the
    ranks, bf16 operands / fp32 master weights.  Returned as the extra key "train" of the JSON line."""
    import torch.distributed as dist
    from video2music_b200 import VideoMusicTransformer, _lib
    from video2music_b200 import synthetic as syn
    from video2music_b200.trainer import Trainer, shard_range
    m = VideoMusicTransformer(total_vf_dim=syn.vf_dim(0), rpr=True, dropout=TRAIN_DROPOUT)     # the reference's training default
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(syn.fill_like_reference_init(shapes, seed=1), strict=False)
    m = m.to(dev).train().set_compute_dtype(dtype)
    tr = Trainer(m, use_graph=True)                     # the step is replayed from one CUDA graph after two eager steps
    b0, b1 = shard_range(TRAIN_GLOBAL_BATCH, rank, world)
    inp = syn.make_inputs(b1 - b0, 4321 + rank, SEQ - 1, 300, 0)
    host = {k: v.pin_memory() for k, v in inp.items()}
    for _ in range(4):                                  # two eager steps, the capture, one replay
        tr.train_step(host)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    _lib.reset_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nxt = tr.prefetch(host)                             # H2D of every step's batch is inside the timed region; the copy of
    for i in range(args.train_steps):                   # step i+1 runs on a side stream while step i computes
        cur = nxt
        if i + 1 < args.train_steps:
            nxt = tr.prefetch(host)
        loss = tr.train_step(cur)
    lossv = float(loss)                                 # D2H of the loss
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / args.train_steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    flops = 69.4e9 * TRAIN_GLOBAL_BATCH                 # 3 x 23.14 GF per sample
    tfl = flops / (ms * 1e-3) / 1e12
    _, sustained, psrc = tensor_peaks()
    # the same step without dropout (what the kernels cost without re-hashing the keep masks), batch resident: informational
    no_drop = None
    if world == 1:
        for mod in m.modules():
            if hasattr(mod, "dropout") and isinstance(getattr(mod, "dropout"), float):
                mod.dropout = 0.0
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        m.dropout_p = 0.0 if hasattr(m, "dropout_p") else None
        try:
            tr2 = Trainer(m, use_graph=True)
            dbatch = {k: v.to(dev) for k, v in inp.items()}
            for _ in range(4):
                tr2.train_step(dbatch)
            torch.cuda.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(args.train_steps):
                tr2.train_step(dbatch)
            f1.record()
            f1.synchronize()
            ms2 = f0.elapsed_time(f1) / args.train_steps
            no_drop = {"ms_per_step": ms2, "samples_per_s": TRAIN_GLOBAL_BATCH / (ms2 * 1e-3), "tflops": flops / (ms2 * 1e-3) / 1e12,
                       "note": "dropout 0, batch resident on the device, CUDA graph"}
        except Exception as exc:                        # informational only: never fails the bench line
            no_drop = {"error": str(exc)[:200]}
    return {"no_dropout": no_drop, "roofline": {"bound": "tensor", "achieved": tfl / world, "peak": sustained, "unit": "TFLOP/s", "frac": tfl / world / sustained,
                         "peak_source": psrc + ", sustained cuBLAS bf16 (a kernel timed inside a long step)",
                         "note": "per GPU: 69.4 GFLOP per sample (3 x the dense forward, SURVEY 8d) x global batch / step time / n_gpus"},
            "metric": "train_samples_per_s", "value": TRAIN_GLOBAL_BATCH / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms,
            "global_batch": TRAIN_GLOBAL_BATCH, "per_gpu_batch": b1 - b0, "dropout": TRAIN_DROPOUT, "scaling": "strong", "steps": args.train_steps,
            "loss": lossv, "tflops": flops / (ms * 1e-3) / 1e12, "launches_per_step": tr.launches_per_step or _lib.launches() // args.train_steps, "cuda_graph": tr._graph is not None,
            "collective": ("NCCL all-reduce of the fp32 gradients in %d buckets launched from backward hooks (overlapped with "
                           "backward inside the captured graph), 130 MB per step; launch order last step: %s"
                           % (len(tr.buckets.bounds), tr.buckets.order)) if (world > 1 and tr.buckets is not None)
            else ("NCCL all-reduce of the flat fp32 gradient buffer (130 MB)" if world > 1 else "none (1 rank)"),
            "optimizer": "Adam betas (0.9, 0.98) eps 1e-8, LambdaLR(LrStepTracker) schedule (train.py:237-253)",
            "timed": "e2e: every step copies its batch from pinned host memory (side stream, overlapped with the previous step), loss read back at the end",
            "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in host.values())}


def variants_leg(dev):
    """BASELINE configs 4 and 5 (extra key "variants"; fp32 exact training path, dropout 0, synthetic inputs, CUDA events):
    one forward + backward of (a) a 6-layer encoder of MultiheadGQA(8 q / 2 kv heads) + MoELayer(6 experts, top-2) blocks,
    (b) VideoRegression with the Bi-Mamba+ backbone at 64 videos x 300 s, (c) a Mamba block at the 4096-token stress shape."""
    import torch.nn as nn
    from video2music_b200 import (GLUExpert, MoELayer, MultiheadGQA, TransformerEncoder, TransformerEncoderLayer, VideoRegression)
    from video2music_b200.mamba import MambaBlock, MambaConfig

    def timed(fn, reps=3):
        for _ in range(2):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) / reps

    def step(net, fwd):
        def f():
            net.zero_grad(set_to_none=True)
            fwd().sum().backward()
        return f

    g = torch.Generator().manual_seed(11)
    out = {}
    torch.manual_seed(0)
    layer = TransformerEncoderLayer(MultiheadGQA(512, 8, 2, dropout=0.0), MoELayer(GLUExpert(512, 1024, 0.0), 512, n_experts=6,
                                    n_experts_per_token=2, dropout=0.0), pre_norm=False, norm=nn.LayerNorm(512), dropout=0.0)
    enc = TransformerEncoder(layer, 6, nn.LayerNorm(512)).to(dev).train()
    src = torch.randn(300, 8, 512, generator=g).to(dev)
    ms = timed(step(enc, lambda: enc(src)))
    out["gqa_moe_encoder_train"] = {"ms_per_step": ms, "samples_per_s": 8 / (ms * 1e-3), "shape": "6 layers, 8 videos x 300 tokens, d 512, "
                                    "8 q / 2 kv heads, 6 experts top-2 ff 1024", "dtype": "f32"}
    for mod in enc.modules():
        if isinstance(mod, MoELayer):
            mod.compute_dtype = torch.bfloat16                  # experts on the grouped tcgen05 GEMMs (forward and backward)
    ms = timed(step(enc, lambda: enc(src)))
    out["gqa_moe_encoder_train_moe_bf16"] = {"ms_per_step": ms, "samples_per_s": 8 / (ms * 1e-3), "shape": out["gqa_moe_encoder_train"]["shape"],
                                             "dtype": "bf16 experts (grouped tcgen05 GEMMs, K-grouped ragged dW), fp32 GQA attention / router / norms"}
    for mod in enc.modules():
        if isinstance(mod, MultiheadGQA):
            mod.compute_dtype = torch.bfloat16                  # projections on the tcgen05 GEMM, tcgen05 attention forward, tensor-core backward
    ms = timed(step(enc, lambda: enc(src)))
    bf16_desc = "bf16 GQA attention (tcgen05 projections + forward, tensor-core backward) and bf16 experts; fp32 master weights, router, residual stream"
    out["gqa_moe_encoder_train_bf16"] = {"ms_per_step": ms, "samples_per_s": 8 / (ms * 1e-3), "shape": out["gqa_moe_encoder_train"]["shape"],
                                         "dtype": bf16_desc, "note": "8 videos: bound by the ~900 eager launches of the step"}
    src64 = torch.randn(300, 64, 512, generator=g).to(dev)
    ms = timed(step(enc, lambda: enc(src64)))
    out["gqa_moe_encoder_train_bf16_b64"] = {"ms_per_step": ms, "samples_per_s": 64 / (ms * 1e-3), "shape": "6 layers, 64 videos x 300 tokens, d 512, "
                                             "8 q / 2 kv heads, 6 experts top-2 ff 1024", "dtype": bf16_desc}
    for mod in enc.modules():
        if isinstance(mod, (MoELayer, MultiheadGQA)):
            mod.compute_dtype = torch.float32
    ms = timed(step(enc, lambda: enc(src64)))
    out["gqa_moe_encoder_train_b64"] = {"ms_per_step": ms, "samples_per_s": 64 / (ms * 1e-3), "shape": out["gqa_moe_encoder_train_bf16_b64"]["shape"],
                                        "dtype": "f32"}
    del enc, layer, src64
    reg = VideoRegression(n_layers=6, d_model=128, d_hidden=256, dropout=0.0, total_vf_dim=774, regModel="bimamba+").to(dev).train()
    sem, emo = torch.randn(64, 300, 768, generator=g).to(dev), torch.softmax(torch.randn(64, 300, 6, generator=g), -1).to(dev)
    zz = torch.zeros(64, 300, device=dev)

    def reg_fwd():
        ln, inst = reg(sem, zz, zz, emo)
        return ln.sum() + inst.sum()
    ms = timed(step(reg, reg_fwd))
    out["video_regression_bimamba_plus_train"] = {"ms_per_step": ms, "samples_per_s": 64 / (ms * 1e-3),
                                                  "shape": "6 Bi-Mamba+ layers, 64 videos x 300 s, d_model 128, d_inner 256, d_state 16", "dtype": "f32"}
    from video2music_b200.mamba import MambaBlock, _FFN
    for mod in reg.modules():
        if isinstance(mod, (MambaBlock, _FFN)):
            mod.compute_dtype = torch.bfloat16                  # in_proj / out_proj / feed-forward on the tcgen05 GEMM
    ms = timed(step(reg, reg_fwd))
    out["video_regression_bimamba_plus_train_bf16"] = {"ms_per_step": ms, "samples_per_s": 64 / (ms * 1e-3),
                                                       "shape": out["video_regression_bimamba_plus_train"]["shape"],
                                                       "dtype": "bf16 wide projections and feed-forward (tcgen05 GEMMs); conv, scan, x_proj / dt_proj, master weights fp32"}
    del reg
    # config 4 generation: KV-cached, batched greedy decode of the GQA (8 q / 2 kv heads) + MoE (6 experts, top-2) shell
    from video2music_b200 import _lib
    from video2music_b200 import synthetic as syn
    from video2music_b200.video_music_transformer_v2 import VideoMusicTransformer_GQA
    torch.manual_seed(0)
    gm = VideoMusicTransformer_GQA(total_vf_dim=syn.vf_dim(0)).eval()
    gm.load_state_dict(syn.fill_like_reference_init({k: tuple(v.shape) for k, v in gm.state_dict().items()}, seed=7))
    gm = gm.to(dev)
    gi = syn.make_inputs(BATCH, 1234, SEQ - 1, 300, 0)
    feats = [gi[k].to(dev) for k in ("feature_semantic_list", "feature_key", "feature_scene_offset", "feature_motion", "feature_emotion")]
    one = torch.tensor([1])
    gen_fn = lambda n: gm.generate_cached(*feats, primer=one, primer_root=one, primer_attr=torch.tensor([0]), target_seq_length=n,
                                          beam=1, beam_chance=1.0)
    gen_fn(SEQ)                                          # first generation of a configuration: captures the per-position CUDA graph
    torch.cuda.synchronize(dev)
    n0 = _lib.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gen_fn(SEQ)
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1)
    out["gqa_moe_generate"] = {"tokens_per_s": BATCH * (SEQ - 1) / (ms * 1e-3), "ms_per_generation": ms, "launches": _lib.launches() - n0,
                               "shape": "%d videos x %d positions, 6 layers, 8 q / 2 kv heads, 6 experts top-2, fp32, KV cache "
                                        "(self %d MB for 2 kv heads, cross K|V projected once)" % (BATCH, SEQ, 6 * 2 * BATCH * SEQ * 128 * 4 // (1 << 20)),
                               "dtype": "f32", "cuda_graph": True,
                               "note": "one position = ONE captured CUDA graph replayed 298 times (device-side position; step kernels of csrc/step_f32.cu); "
                                       "the literal loop of the reference re-runs the whole model on the growing prefix for every token, batch 1"}
    del gm, feats
    blk = MambaBlock(MambaConfig(d_model=128, n_layers=1)).to(dev).train()
    xb = torch.randn(8, 4096, 128, generator=g).to(dev)
    ms = timed(step(blk, lambda: blk(xb)))
    out["mamba_block_train_L4096"] = {"ms_per_step": ms, "tokens_per_s": 8 * 4096 / (ms * 1e-3), "shape": "8 x 4096 tokens, d_model 128, "
                                      "d_inner 256, d_state 16", "dtype": "f32"}
    return out


def kernel_rooflines(dev):
    """roofline_extra: the other kernels BASELINE.json names, each timed alone in this run (CUDA events on the launching
    stream, median of 5, L2 flushed between launches) at the BASELINE shapes: fused RPR attention forward and a projection
    GEMM against the measured bf16 burst peak, pscan forward / backward against the measured HBM copy bandwidth."""
    from video2music_b200 import ops
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    hbm, hsrc = peaks()
    burst, _, tsrc = tensor_peaks()

    def timed(fn, reps=5):
        for _ in range(5):                                 # warm-up (the first kernel follows a different workload: clocks, caches)
            fn()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2]

    out = []
    g = torch.Generator(device="cpu").manual_seed(3)
    B, L, S, H, dh, E = TRAIN_GLOBAL_BATCH, SEQ - 1, 300, 8, 64, 512
    qkv = (torch.randn(B, L, 3 * E, generator=g) * 0.3).to(dev).bfloat16()
    Er = (torch.randn(300, dh, generator=g) * 0.3).to(dev).bfloat16()
    o = torch.empty(B, L, E, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B * H, L, device=dev, dtype=torch.float32)
    st = (L * 3 * E, 3 * E)
    ms = timed(lambda: ops.attention(qkv, qkv[:, :, E:], qkv[:, :, 2 * E:], o, B=B, Hq=H, Hkv=H, Lq=L, Lk=L, dh=dh, q_strides=st,
                                     k_strides=st, v_strides=st, o_strides=(L * E, E), causal=True, Er=Er, lse=lse))
    fl = 6.0 * L * L * dh * B * H                          # QK^T + Q Er^T + P V, dense (3 * 2 * T^2 * d per layer and sample)
    out.append({"kernel": "attn_bf16_tc_kernel (fused RPR causal self-attention forward)", "shape": "B=%d H=%d L=%d dh=%d" % (B, H, L, dh),
                "bound": "tensor", "ms": ms, "achieved": fl / ms / 1e9, "peak": burst, "unit": "TFLOP/s", "frac": fl / ms / 1e9 / burst,
                "achieved_causal": 0.5 * fl / ms / 1e9, "peak_source": tsrc + ", burst",
                "note": "achieved = dense-equivalent flops (the causal half the kernel skips is counted); achieved_causal counts only the lower triangle"})
    kv = (torch.randn(B, S, 2 * E, generator=g) * 0.3).to(dev).bfloat16()
    sk = (S * 2 * E, 2 * E)
    ms = timed(lambda: ops.attention(qkv, kv, kv[:, :, E:], o, B=B, Hq=H, Hkv=H, Lq=L, Lk=S, dh=dh, q_strides=st, k_strides=sk,
                                     v_strides=sk, o_strides=(L * E, E), causal=False, lse=lse))
    fl = 4.0 * L * S * dh * B * H
    out.append({"kernel": "attn_bf16_tc_kernel (cross-attention forward)", "shape": "B=%d H=%d L=%d S=%d dh=%d" % (B, H, L, S, dh),
                "bound": "tensor", "ms": ms, "achieved": fl / ms / 1e9, "peak": burst, "unit": "TFLOP/s", "frac": fl / ms / 1e9 / burst,
                "peak_source": tsrc + ", burst"})
    M = B * L
    x = qkv.view(M, 3 * E)[:, :E].contiguous()
    w = (torch.randn(3 * E, E, generator=g) * 0.05).to(dev).bfloat16()
    bias = torch.randn(3 * E, generator=g).to(dev)
    ms = timed(lambda: ops.linear(x, w, bias, out_dtype=torch.bfloat16))
    fl = 2.0 * M * 3 * E * E
    out.append({"kernel": "gemm_bf16_tc_kernel (in_proj)", "shape": "M=%d N=%d K=%d" % (M, 3 * E, E), "bound": "tensor", "ms": ms,
                "achieved": fl / ms / 1e9, "peak": burst, "unit": "TFLOP/s", "frac": fl / ms / 1e9 / burst, "peak_source": tsrc + ", burst"})
    del qkv, kv, o, x
    for (b_, l_) in ((64, 300), (8, 4096)):
        A = torch.rand(b_, l_, 256, 16, generator=g).mul_(0.99).to(dev)
        X = torch.randn(b_, l_, 256, 16, generator=g).to(dev)
        ms = timed(lambda: ops.pscan_fwd(A, X))
        by = 3.0 * 4 * A.numel()
        out.append({"kernel": "pscan_tma_kernel forward", "shape": "(%d,%d,256,16)" % (b_, l_), "bound": "hbm", "ms": ms, "achieved": by / ms / 1e6,
                    "peak": hbm, "unit": "GB/s", "frac": by / ms / 1e6 / hbm, "peak_source": hsrc, "note": "12 B per element: read A, X, write H"})
        Hh = ops.pscan_fwd(A, X)
        ms = timed(lambda: ops.pscan_bwd(A, Hh, X))
        by = 5.0 * 4 * A.numel()
        out.append({"kernel": "pscan_tma_kernel backward", "shape": "(%d,%d,256,16)" % (b_, l_), "bound": "hbm", "ms": ms, "achieved": by / ms / 1e6,
                    "peak": hbm, "unit": "GB/s", "frac": by / ms / 1e6 / hbm, "peak_source": hsrc,
                    "note": "20 B per element: read gradH, A, H, write gradA, gradX"})
        del A, X, Hh
    return out


def fp32_leg(dev, devt, prim, pr, pa, l2_flush, steps=2):
    """Same workload on the fp32 exact path (the one whose greedy tokens are bit-exact against the reference): value_fp32."""
    model32, _ = make_model(torch.float32, dev)

    def gen32():
        return model32.generate(devt["feature_semantic_list"], devt["feature_key"], devt["feature_scene_offset"], devt["feature_motion"],
                                devt["feature_emotion"], primer=prim, primer_root=pr, primer_attr=pa, target_seq_length=SEQ, beam=1,
                                beam_chance=1.0)
    gen32()
    ms = []
    for _ in range(steps):
        l2_flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gen32(); e1.record(); e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = sum(ms) / len(ms)
    run_bytes, _ = algorithmic_bytes(BATCH, 300, SEQ, 512, 8, 1024, 6, 159, 4)
    peak, _ = peaks()
    val = BATCH * (SEQ - 1) / (t / 1e3)
    return {"value": val, "unit": UNIT, "ms_per_step": t, "steps": steps, "dtype": "f32",
            "path": "decode step kernels (skinny_gemm with coalesced loads + transpose-reduce / dec_attn, CUDA graph of 51 launches per position), fp32 prefill",
            "hbm_floor_tokens_per_s": BATCH * (SEQ - 1) / (run_bytes / (peak * 1e9)), "frac_of_floor": val / (BATCH * (SEQ - 1) / (run_bytes / (peak * 1e9)))}


def reference_arm(args):
    """The reference's algorithm on the host cores (oracle port, literal re-forward loop, batch 1)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import amt_oracle as O
    from video2music_b200 import synthetic as syn
    torch.set_num_threads(os.cpu_count() or 1)
    _, sd = make_model(torch.float32, None)
    inp = syn.make_inputs(1, 1234, SEQ - 1, 300, 0)
    prim, pr, pa = torch.tensor([1]), torch.tensor([1]), torch.tensor([0])

    def run(n):
        t0 = time.perf_counter()
        with torch.no_grad():
            O.generate_greedy_literal(sd, inp["feature_semantic_list"], inp["feature_key"][0], inp["feature_scene_offset"],
                                      inp["feature_motion"], inp["feature_emotion"], prim, pr, pa, n, chord_embed=True)
        return time.perf_counter() - t0

    # Same target as our arm: 299 greedy tokens (target_seq_length 300).  The bounded sample is ONE video per step (our arm
    # generates 64 per GPU; the metric is per token, and the reference's loop is batch 1 anyway, generate.py:368-392): one
    # step is ~5 s on 16 cores, so --steps 20 --warmup 5 ends in ~2.5 minutes.
    seq = SEQ
    for _ in range(args.warmup):
        run(seq)
    times = [run(seq) for _ in range(args.steps)]
    tokens = seq - 1
    val = tokens / (sum(times) / len(times))
    out = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": bench_config(),
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                            "sample": "1 video x %d greedy tokens per step (same target_seq_length %d as the GPU arm), the "
                                      "reference's algorithm: batch 1, no KV cache, full re-forward per token "
                                      "(video_music_transformer.py:1069-1084), oracle port on the host cores" % (tokens, seq)},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step leg (BASELINE config 3)")
    ap.add_argument("--train-steps", type=int, default=10)
    ap.add_argument("--no-extra", action="store_true", help="skip the fp32 generation value and the per-kernel roofline_extra list")
    ap.add_argument("--no-variants", action="store_true", help="skip the GQA+MoE / Mamba training timings (BASELINE configs 4, 5)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch.distributed as dist
    from video2music_b200 import _lib, engine
    from video2music_b200 import synthetic as syn

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    model, sd = make_model(dtype, dev)
    # videos are sharded over ranks (no collective on the data path): rank r owns videos [r*64, (r+1)*64)
    inp = syn.make_inputs(BATCH, 1234 + rank, SEQ - 1, 300, 0)
    keys = ("feature_semantic_list", "feature_key", "feature_scene_offset", "feature_motion", "feature_emotion")
    host = {k: inp[k].pin_memory() for k in keys}
    devt = {k: inp[k].to(dev) for k in keys}
    prim, pr, pa = torch.tensor([1]), torch.tensor([1]), torch.tensor([0])
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def gen(src):
        return model.generate(src["feature_semantic_list"], src["feature_key"], src["feature_scene_offset"],
                              src["feature_motion"], src["feature_emotion"], primer=prim, primer_root=pr, primer_attr=pa,
                              target_seq_length=SEQ, beam=1, beam_chance=1.0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps, each bracketed by CUDA events on the launching stream, L2 flushed (untimed) between steps."""
        ms = []
        for _ in range(steps):
            l2_flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        return ms

    def e2e_step():
        d = {k: host[k].to(dev, non_blocking=True) for k in keys}
        toks = gen(d)
        return toks.to("cpu", non_blocking=False)

    for _ in range(args.warmup):
        gen(devt)
        e2e_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.reset_launches()
    barrier()
    ms = timed(lambda: gen(devt), args.steps)
    barrier()
    launches = _lib.launches()
    ms_e2e = timed(e2e_step, args.steps)
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # ---- dominant kernel: the streamed cluster decode kernel (ONE launch = all 299 positions of the rank's 64 videos),
    # timed alone with CUDA events on the stream it is launched on (torch's current stream)
    kern_ms = []
    for _ in range(3):
        st = engine.build_decode(model._w(), model._cfg(), devt["feature_semantic_list"], devt["feature_key"].reshape(-1),
                                 devt["feature_scene_offset"], devt["feature_motion"], devt["feature_emotion"],
                                 prim, pr, pa, SEQ)
        torch.cuda.synchronize()
        l2_flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        engine.run_decode(st, SEQ - 1)
        e1.record()
        e1.synchronize()
        kern_ms.append(e0.elapsed_time(e1))
        decode_mode = st.mode
        del st
    kern_ms = sorted(kern_ms)[1]

    t_step = sum(ms) / len(ms)
    t_e2e = sum(ms_e2e) / len(ms_e2e)
    train = train_leg(args, dev, rank, world, dtype) if not args.no_train else None
    if world > 1:
        t = torch.tensor([t_step, t_e2e, kern_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)                # device-timed, max over ranks
        t_step, t_e2e, kern_ms = float(t[0]), float(t[1]), float(t[2])
    tokens = world * BATCH * (SEQ - 1)
    esz = 2 if dtype == torch.bfloat16 else 4
    run_bytes, _ = algorithmic_bytes(BATCH, 300, SEQ, 512, 8, 1024, 6, 159, esz)
    peak, peak_src = peaks()
    h2d = sum(host[k].numel() * host[k].element_size() for k in keys)
    kname = "decode_stream_kernel (one launch: 299 positions x 64 videos, 6 layers)" if decode_mode == "stream" else \
        "decode step kernels (skinny_gemm / dec_attn, 70 launches per position)"
    traffic, traffic_src = ncu_traffic("decode_stream_kernel") if decode_mode == "stream" else (None, None)
    out = {
        "metric": METRIC, "value": tokens / (t_step / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype if args.dtype == "bf16" else "f32", "data": "synthetic",
        "config": bench_config(),
        "e2e": {"value": tokens / (t_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": BATCH * SEQ * 8, "ms_per_step": t_e2e},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"kernel": kname, "bound": "hbm", "achieved": run_bytes / (kern_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": run_bytes / (kern_ms * 1e-3) / 1e9 / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src, "bytes_per_launch": run_bytes, "us_per_launch": kern_ms * 1e3,
                     "note": "algorithmic bytes = SURVEY 8d decode byte floor (weights once per position + every cached K/V "
                             "element once per position) for the 299 positions one launch processes"},
        "share_of_step": kern_ms / t_step,
    }
    if train is not None:
        out["train"] = train
    if rank == 0 and world == 1 and not args.no_extra:
        try:
            out["value_fp32"] = fp32_leg(dev, devt, prim, pr, pa, l2_flush)
        except Exception as e:
            out["value_fp32"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
        try:
            out["roofline_extra"] = kernel_rooflines(dev)
        except Exception as e:
            out["roofline_extra"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
    if rank == 0 and world == 1 and not args.no_variants:
        try:
            out["variants"] = variants_leg(dev)
        except Exception as e:                                   # the headline line must survive a failure of the extra leg
            out["variants"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import amt_oracle as O
        torch.set_num_threads(os.cpu_count() or 1)
        one = {k: inp[k][:1] for k in keys}
        n_tok = 300
        t0 = time.perf_counter()
        with torch.no_grad():
            O.generate_greedy_literal(sd, one["feature_semantic_list"], one["feature_key"][0], one["feature_scene_offset"],
                                      one["feature_motion"], one["feature_emotion"], prim, pr, pa, n_tok, chord_embed=True)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": (n_tok - 1) / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                               "sample": "1 video x %d greedy tokens of the reference algorithm (batch 1, no KV cache, "
                                         "full re-forward per token) on the host" % (n_tok - 1)}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
