"""Kernel-level execution plan of the AMT forward pass and of KV-cached generation.

Activations are kept batch-first as 2-D row-major matrices [B*L, d] (row = b*L + l), so that every
(video, head) attention problem is a strided slice of one GEMM output and every linear layer is one
GEMM over all videos.  The seq-first (L, B, E) layout of the reference's modules is reproduced only
at the module API surface (rpr.py / video_music_transformer.py).

Reference lines are cited next to each step (paths relative to the reference root).
"""
import ctypes as C
import math
from typing import Dict, Optional

import torch

from . import _lib, ops
from ._lib import Decode, check, load, ptr, stream

CHORD_END, CHORD_PAD, CHORD_SIZE = 157, 158, 159            # utilities/constants.py:50-52
CHORD_ROOT_PAD, CHORD_ATTR_PAD = 14, 15                     # utilities/constants.py:55-62


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


class AMTWeights:
    """Weights of a VideoMusicTransformer resolved by state_dict key in the compute dtype.

    fp32: the nn.Parameters themselves (no copies).  bf16: cached bf16 copies of the matrices with the
    leading dimension padded to a multiple of 8 elements (TMA needs a 16-byte row pitch); vectors
    (biases, LayerNorm affine, embedding tables, positional encodings) always stay fp32.
    The cache is invalidated when any parameter's version counter changes (optimizer step, load_state_dict).
    """

    def __init__(self, module: torch.nn.Module, dtype: torch.dtype):
        self.module = module
        self.dtype = dtype
        self._sd: Dict[str, torch.Tensor] = {}
        self._cache: Dict[str, torch.Tensor] = {}
        self._versions = None
        self.flat16: Optional[Dict[str, torch.Tensor]] = None    # trainer-owned bf16 mirror: name -> view (see trainer.Trainer)
        self.refresh()

    def refresh(self) -> None:
        sd = dict(self.module.named_parameters())
        sd.update(dict(self.module.named_buffers()))
        versions = tuple((k, v._version, v.data_ptr()) for k, v in sd.items())
        if versions != self._versions:
            if self._versions is not None:     # changed through torch (load_state_dict, another optimiser): the trainer's
                self.flat16 = None             # bf16 mirror no longer matches; it is handed back at its next step
            self._sd, self._cache, self._versions = sd, {}, versions

    def invalidate(self) -> None:
        """Drops the compute-dtype copies: for in-place updates that bypass torch's version counters (the fused Adam kernel
        writes the flat parameter buffer through a raw pointer)."""
        self._versions = None

    def has(self, name: str) -> bool:
        return name in self._sd

    def f(self, name: str) -> torch.Tensor:
        """fp32 tensor as stored."""
        return self._sd[name].detach()

    def w(self, name: str, rows: Optional[slice] = None, cols: Optional[int] = None) -> torch.Tensor:
        """Matrix [N, K] in the compute dtype (row slice / leading `cols` columns optional)."""
        if self.dtype == torch.bfloat16 and self.flat16 is not None and cols is None:
            t = self.flat16.get(name)
            if t is not None and t.shape[1] % 8 == 0:              # 16-byte row pitch: usable by TMA as it is
                return t[rows] if rows is not None else t
        key = "%s|%s|%s" % (name, rows, cols)
        t = self._cache.get(key)
        if t is None:
            src = self._sd[name].detach()
            if rows is not None:
                src = src[rows]
            if cols is not None:
                src = src[:, :cols]
            if self.dtype == torch.float32:
                t = src                                    # strided view is fine for the SIMT GEMM
            else:
                t = ops.cast_2d(src, torch.bfloat16, _pad8(src.shape[1]))
            self._cache[key] = t
        return t

    def col(self, name: str, col: int) -> torch.Tensor:
        key = "%s|col%d" % (name, col)
        t = self._cache.get(key)
        if t is None:
            t = self._sd[name].detach()[:, col].contiguous()
            self._cache[key] = t
        return t

    def table(self, name: str) -> torch.Tensor:
        """Vector-like tensor converted to the compute dtype (Er)."""
        if self.dtype == torch.bfloat16 and self.flat16 is not None and name in self.flat16:
            return self.flat16[name]
        key = "%s|tab" % name
        t = self._cache.get(key)
        if t is None:
            src = self._sd[name].detach()
            t = src if self.dtype == torch.float32 else ops.cast_2d(src.contiguous(), torch.bfloat16)
            self._cache[key] = t
        return t


def _mha_self(W: AMTWeights, p: str, x: torch.Tensor, B: int, L: int, E: int, H: int, causal: bool,
              er: Optional[torch.Tensor]) -> torch.Tensor:
    """Self-attention block up to and including out_proj + residual: returns x + MHA(x) (pre-LayerNorm).
    rpr.py:253 (in-proj), :328 (q scaling), :387-414 (attention core), :417 (out-proj), :58/:104 (residual)."""
    dh = E // H
    qkv = ops.linear(x, W.w(p + "in_proj_weight"), W.f(p + "in_proj_bias"), k=E, alpha=float(dh) ** -0.5, alpha_cols=E)
    ctx = torch.empty((B * L, E), device=x.device, dtype=x.dtype)
    ld = qkv.stride(0)
    ops.attention(qkv, qkv[:, E:], qkv[:, 2 * E:], ctx, B=B, Hq=H, Hkv=H, Lq=L, Lk=L, dh=dh,
                  q_strides=(L * ld, ld), k_strides=(L * ld, ld), v_strides=(L * ld, ld), o_strides=(L * E, E),
                  causal=causal, Er=er)
    return ops.linear(ctx, W.w(p + "out_proj.weight"), W.f(p + "out_proj.bias"), k=E, residual=x)


def _ffn(W: AMTWeights, p: str, x: torch.Tensor, E: int) -> torch.Tensor:
    """x + linear2(relu(linear1(x)))  (rpr.py:67-68)."""
    hdn = ops.linear(x, W.w(p + "linear1.weight"), W.f(p + "linear1.bias"), k=E, relu=True)
    return ops.linear(hdn, W.w(p + "linear2.weight"), W.f(p + "linear2.bias"), k=hdn.shape[1], residual=x)


def _ln(W: AMTWeights, p: str, x: torch.Tensor) -> torch.Tensor:
    return ops.layernorm(x, W.f(p + ".weight"), W.f(p + ".bias"))


def encode_memory(W: AMTWeights, cfg, sem, scene, motion, emotion) -> torch.Tensor:
    """Video stream: concat features, Linear_vis, + PE, 6 post-norm encoder layers, final norm
    (video_music_transformer.py:1003-1033; stock nn.TransformerEncoder built at :967-971). -> [B*S, E]"""
    B, S = sem.shape[0], sem.shape[1]
    E, H = cfg["d_model"], cfg["nhead"]
    vf_dim = W.f("Linear_vis.weight").shape[1]
    ld = vf_dim if W.dtype == torch.float32 else _pad8(vf_dim)
    vin = ops.concat_features(sem, scene, motion, emotion, W.dtype, ld)
    x = ops.linear(vin, W.w("Linear_vis.weight"), W.f("Linear_vis.bias"), k=vf_dim,
                   residual=W.f("positional_encoding_video.pe").view(-1, E), res_mod=S)        # :1022,1030
    for l in range(cfg["n_layers"]):
        p = "transformer.encoder.layers.%d." % l
        x = _ln(W, p + "norm1", _mha_self(W, p + "self_attn.", x, B, S, E, H, causal=False, er=None))
        x = _ln(W, p + "norm2", _ffn(W, p, x, E))
    return _ln(W, "transformer.encoder.norm", x)


def chord_stream(W: AMTWeights, cfg, x, x_root, x_attr, feature_key) -> torch.Tensor:
    """Chord stream input: (embedding_root + embedding_attr | chord embedding) ++ key -> Linear_chord -> + PE
    (video_music_transformer.py:984-1001,1029). -> [B*T, E]"""
    B, T = x.shape
    E = cfg["d_model"]
    if cfg["chord_embed"]:
        e = ops.embed_sum(x, W.f("chord_embedding_model.weight"), None, None, W.dtype)
    else:
        e = ops.embed_sum(x_root, W.f("embedding_root.weight"), x_attr, W.f("embedding_attr.weight"), W.dtype)
    key_rows = feature_key.reshape(B, 1).float().expand(B, T).contiguous().view(-1)
    return ops.linear(e, W.w("Linear_chord.weight", cols=E), W.f("Linear_chord.bias"), k=E,
                      row_scale=key_rows, col_vec=W.col("Linear_chord.weight", E),
                      residual=W.f("positional_encoding.pe").view(-1, E), res_mod=T)


def amt_forward(W: AMTWeights, cfg, x, x_root, x_attr, sem, key, scene, motion, emotion, mask: bool = True) -> torch.Tensor:
    """VideoMusicTransformer.forward (video_music_transformer.py:978-1044), eval semantics. -> (B, T, CHORD_SIZE) fp32."""
    W.refresh()
    B, T = x.shape
    S = sem.shape[1]
    E, H = cfg["d_model"], cfg["nhead"]
    dh = E // H
    mem = encode_memory(W, cfg, sem, scene, motion, emotion)
    xf = chord_stream(W, cfg, x, x_root, x_attr, key)
    for l in range(cfg["n_layers"]):
        p = "transformer.decoder.layers.%d." % l
        er = W.table(p + "self_attn.Er") if W.has(p + "self_attn.Er") else None
        xf = _ln(W, p + "norm1", _mha_self(W, p + "self_attn.", xf, B, T, E, H, causal=bool(mask), er=er))   # rpr.py:56-59
        # cross attention over the video memory (rpr.py:62-65): q from the chords, k|v from memory
        q = ops.linear(xf, W.w(p + "multihead_attn.in_proj_weight", rows=slice(0, E)),
                       W.f(p + "multihead_attn.in_proj_bias")[:E], k=E, alpha=float(dh) ** -0.5, alpha_cols=E)
        kv = ops.linear(mem, W.w(p + "multihead_attn.in_proj_weight", rows=slice(E, 3 * E)),
                        W.f(p + "multihead_attn.in_proj_bias")[E:], k=E)
        ctx = torch.empty((B * T, E), device=xf.device, dtype=xf.dtype)
        ops.attention(q, kv, kv[:, E:], ctx, B=B, Hq=H, Hkv=H, Lq=T, Lk=S, dh=dh,
                      q_strides=(T * E, E), k_strides=(S * 2 * E, 2 * E), v_strides=(S * 2 * E, 2 * E),
                      o_strides=(T * E, E), causal=False)
        r = ops.linear(ctx, W.w(p + "multihead_attn.out_proj.weight"), W.f(p + "multihead_attn.out_proj.bias"), k=E,
                       residual=xf)
        xf = _ln(W, p + "norm2", r)
        xf = _ln(W, p + "norm3", _ffn(W, p, xf, E))                                                          # rpr.py:67-69
    xf = _ln(W, "transformer.decoder.norm", xf)                                                               # rpr.py:32-33
    y = ops.linear(xf, W.w("Wout.weight"), W.f("Wout.bias"), k=E, out_dtype=torch.float32)                    # :1042
    return y.view(B, T, -1)


def pack_fragments(w: torch.Tensor) -> torch.Tensor:
    """[N, K] bf16 matrix -> tiles of 16 rows x 16 k in mma.m16n8k16 A-fragment order, [ceil(N/16)][K/16][32 lanes][8]
    (rows zero-padded to a multiple of 16).  One 16-row x 512-k tile is then 16 KB of contiguous memory that the streamed
    decode kernel (csrc/decode_stream.cu) fetches with a single bulk copy and reads with conflict-free 16-byte loads.
    Lane l = 4*g + q holds rows (g, g+8) x columns (2q, 2q+1, 2q+8, 2q+9) of a tile: [k-half][row-half][pair]."""
    assert w.dim() == 2 and w.dtype == torch.bfloat16 and w.shape[1] % 16 == 0
    N, K = w.shape
    Np = (N + 15) // 16 * 16
    if Np != N:
        w = torch.cat([w, w.new_zeros((Np - N, K))], 0)
    v = w.contiguous().view(Np // 16, 2, 8, K // 16, 2, 4, 2)          # n16, row-half, g, k16, k-half, q, pair
    return v.permute(0, 3, 2, 5, 4, 1, 6).contiguous().view(Np // 16, K // 16, 32, 8)


class DecodeState:
    """Device buffers of one batched KV-cached generation run (kept alive while kernels are in flight)."""

    def __init__(self):
        self.keep = []
        self.params: Optional[Decode] = None
        self.packed: Optional[Decode] = None      # same buffers, fragment-packed matrices (streamed cluster kernel)
        self.mode = "kernels"
        self.gen = None
        self.logits_all = None
        self.launches_per_step = 0
        self.pos = 0            # host copy of the device position counter
        self.step = None


def swizzled_er_copies(er: torch.Tensor) -> torch.Tensor:
    """[er_len, 64] bf16 -> [8, er_len, 64]: copy s stores row r with its eight 16-byte chunks at position c ^ ((r - s) & 7),
    so that a bulk copy of rows [start, start + n) taken from copy (start & 7) lands in shared memory with the XOR swizzle
    (slice row & 7) that ldmatrix reads without bank conflicts (csrc/decode_stream.cu)."""
    n = er.shape[0]
    rows = torch.arange(n, device=er.device)
    chunks = torch.arange(8, device=er.device)
    v = er.contiguous().view(n, 8, 8)
    out = torch.empty((8, n, 8, 8), device=er.device, dtype=er.dtype)
    for s_ in range(8):
        src_chunk = chunks[None, :] ^ ((rows[:, None] - s_) & 7)          # position p holds chunk p ^ key (XOR is an involution)
        out[s_] = torch.gather(v, 1, src_chunk[:, :, None].expand(n, 8, 8))
    return out.view(8, n, 64)


def stream_config_ok(dt, E, H, FF, B) -> bool:
    """Configurations the streamed cluster kernel covers (all clusters co-resident: <= 8 videos per 8-SM cluster)."""
    return dt == torch.bfloat16 and E == 512 and H == 8 and FF == 1024 and B <= 8 * 13


def build_decode(W: AMTWeights, cfg, sem, key, scene, motion, emotion, primer, primer_root, primer_attr,
                 target_seq_length: int, want_logits: bool = False, mode: str = "auto", sample: bool = False,
                 max_conseq_N: int = 0, max_conseq_chord: int = 2, uniforms: Optional[torch.Tensor] = None) -> DecodeState:
    """Encoder pass + cross-attention K/V caches + decode buffers.
    Replaces the per-step re-forward of generate() (video_music_transformer.py:1069-1071).
    mode "stream" / "kernels" / "auto" picks the decode path (see run_decode); the caches are built for that path only."""
    W.refresh()
    dev = sem.device
    B, S = sem.shape[0], sem.shape[1]
    E, H, FF, NL = cfg["d_model"], cfg["nhead"], cfg["d_ff"], cfg["n_layers"]
    dh = E // H
    cap = target_seq_length
    if not W.has("transformer.decoder.layers.0.self_attn.Er"):
        raise NotImplementedError("KV-cached generate() is built for the RPR decoder (rpr=True, the shipped AMT, "
                                  "video2music.py:640); this model has no Er tables")
    if E != 512 or dh != 64:
        raise NotImplementedError("the decode kernels are built for d_model 512 with head_dim 64 (got d_model %d, %d heads)" % (E, H))
    dt = W.dtype
    st = DecodeState()
    if mode == "auto":
        mode = "stream" if (stream_config_ok(dt, E, H, FF, B) and (not sample or 1 <= max_conseq_chord <= 8)) else "kernels"
    if mode == "stream" and not stream_config_ok(dt, E, H, FF, B):
        raise RuntimeError("streamed decode needs bf16, d_model 512, 8 heads, dim_feedforward 1024, batch <= 104")
    st.mode = mode
    mem = encode_memory(W, cfg, sem, scene, motion, emotion)

    d = Decode()
    d.dtype = _lib.dtype_code(dt)
    d.B, d.H, d.E, d.FF, d.S, d.cap, d.n_layers = B, H, E, FF, S, cap, NL
    d.vocab, d.vocab_limit = CHORD_SIZE, CHORD_END
    d.chord_embed = int(cfg["chord_embed"])
    d.sample, d.max_conseq_N, d.max_conseq_chord = int(sample), int(max_conseq_N), int(max_conseq_chord)
    if sample:
        # one uniform per (video, position): Categorical(probs).sample() of the reference (:1103-1104) as an inverse-CDF draw
        if uniforms is None:
            uniforms = torch.rand((B, cap), device=dev, dtype=torch.float32)
        assert uniforms.shape == (B, cap) and uniforms.dtype == torch.float32
        uniforms = uniforms.to(dev).contiguous()
        st.keep.append(uniforms)
        d.uniforms = ptr(uniforms)
    if primer.dim() == 1:
        primer, primer_root, primer_attr = (t.view(1, -1).expand(B, -1) for t in (primer, primer_root, primer_attr))
    P = primer.shape[1]
    assert 1 <= P <= cap
    d.primer_len = P
    gen = torch.full((B, cap), CHORD_PAD, dtype=torch.int64, device=dev)                  # :1059-1061
    gen_root = torch.full((B, cap), CHORD_ROOT_PAD, dtype=torch.int64, device=dev)
    gen_attr = torch.full((B, cap), CHORD_ATTR_PAD, dtype=torch.int64, device=dev)
    gen[:, :P] = primer.to(dev)                                                             # :1063-1066
    gen_root[:, :P] = primer_root.to(dev)
    gen_attr[:, :P] = primer_attr.to(dev)
    if mode == "stream":
        # 256-byte rows [K | V] per (video, head, position), chunks XOR-swizzled by position (decode_stream.cu)
        self_kv = torch.zeros((NL, B, H, cap, 2 * dh), device=dev, dtype=dt)
        cross_sw = torch.empty((NL, B, H, S, 2 * dh), device=dev, dtype=dt)
        cross_kv = torch.empty((1, 2, B, H, S, dh), device=dev, dtype=dt)      # staging of one layer (GEMM output layout)
    else:
        self_kv = torch.zeros((NL, 2, B, H, cap, dh), device=dev, dtype=dt)
        cross_kv = torch.empty((NL, 2, B, H, S, dh), device=dev, dtype=dt)
        cross_sw = None
    keep = st.keep
    keep += [mem, gen, gen_root, gen_attr, self_kv, cross_kv, cross_sw]

    def hold(t):
        keep.append(t)
        return ptr(t)

    for l in range(NL):
        p = "transformer.decoder.layers.%d." % l
        L = d.layer[l]
        # K|V of the video memory once per sequence (the reference recomputes them every layer AND every step, rpr.py:62)
        ckv = cross_kv[0 if mode == "stream" else l]
        ops.linear(mem, W.w(p + "multihead_attn.in_proj_weight", rows=slice(E, 3 * E)),
                   W.f(p + "multihead_attn.in_proj_bias")[E:], k=E, out=ckv,
                   head_scatter=dict(S=S, H=H, dh=dh, cap=S, pos0=0, part_stride=B * H * S * dh))
        if mode == "stream":
            check(load().v2m_kv_interleave(ptr(ckv[0]), ptr(ckv[1]), ptr(cross_sw[l]), B * H * S, S, stream()))
            _lib.count_launches(1)
        L.w_qkv = hold(W.w(p + "self_attn.in_proj_weight"))
        L.b_qkv = hold(W.f(p + "self_attn.in_proj_bias"))
        L.w_so, L.b_so = hold(W.w(p + "self_attn.out_proj.weight")), hold(W.f(p + "self_attn.out_proj.bias"))
        L.w_cq = hold(W.w(p + "multihead_attn.in_proj_weight", rows=slice(0, E)))
        L.b_cq = hold(W.f(p + "multihead_attn.in_proj_bias")[:E])
        L.w_co = hold(W.w(p + "multihead_attn.out_proj.weight"))
        L.b_co = hold(W.f(p + "multihead_attn.out_proj.bias"))
        L.w_f1, L.b_f1 = hold(W.w(p + "linear1.weight")), hold(W.f(p + "linear1.bias"))
        L.w_f2, L.b_f2 = hold(W.w(p + "linear2.weight")), hold(W.f(p + "linear2.bias"))
        for i in (1, 2, 3):
            setattr(L, "ln%d_g" % i, hold(W.f(p + "norm%d.weight" % i)))
            setattr(L, "ln%d_b" % i, hold(W.f(p + "norm%d.bias" % i)))
        er = W.table(p + "self_attn.Er")
        d.er_len = er.shape[0]
        L.er = hold(er)
        if mode == "stream":
            L.self_k, L.cross_k = ptr(self_kv[l]), ptr(cross_sw[l])
            key_sw = p + "self_attn.Er|sw8"
            if key_sw not in W._cache:
                W._cache[key_sw] = swizzled_er_copies(er)
            L.er_sw = hold(W._cache[key_sw])
        else:
            L.self_k, L.self_v = ptr(self_kv[l, 0]), ptr(self_kv[l, 1])
            L.cross_k, L.cross_v = ptr(cross_kv[l, 0]), ptr(cross_kv[l, 1])
    # the skinny decode GEMMs need dense [N, K] weights (K == leading dimension)
    for l in range(NL):
        p = "transformer.decoder.layers.%d." % l
        for name, rows in ((p + "self_attn.in_proj_weight", None), (p + "multihead_attn.in_proj_weight", slice(0, E))):
            t = W.w(name, rows=rows)
            assert t.is_contiguous(), name
    d.lnf_g, d.lnf_b = hold(W.f("transformer.decoder.norm.weight")), hold(W.f("transformer.decoder.norm.bias"))
    d.w_out, d.b_out = hold(W.w("Wout.weight")), hold(W.f("Wout.bias"))
    d.emb_root, d.emb_attr = hold(W.f("embedding_root.weight")), hold(W.f("embedding_attr.weight"))
    if cfg["chord_embed"]:
        d.emb_chord = hold(W.f("chord_embedding_model.weight"))
    wc = W.w("Linear_chord.weight", cols=E)
    if not wc.is_contiguous():
        wc = wc.contiguous()
    d.w_chord, d.wc_key, d.b_chord = hold(wc), hold(W.col("Linear_chord.weight", E)), hold(W.f("Linear_chord.bias"))
    pe = W.f("positional_encoding.pe").view(-1, E)
    assert pe.shape[0] >= cap, "target_seq_length exceeds max_sequence_chord"
    d.pe = hold(pe)
    d.key = hold(key.reshape(B).float().contiguous().to(dev))
    d.gen, d.gen_root, d.gen_attr = ptr(gen), ptr(gen_root), ptr(gen_attr)
    step = torch.zeros((8,), dtype=torch.int32, device=dev)     # one position counter per sub-batch chain
    d.step = hold(step)
    st.step = step
    d.h = hold(torch.empty((B, E), device=dev, dtype=torch.float32))
    d.r = hold(torch.empty((B, E), device=dev, dtype=dt))
    d.qbuf = hold(torch.empty((B, 3 * E), device=dev, dtype=torch.float32))
    d.ctx = hold(torch.empty((B, E), device=dev, dtype=dt))
    d.xn = hold(torch.empty((B, E), device=dev, dtype=dt))
    d.ff = hold(torch.empty((B, FF), device=dev, dtype=dt))
    d.logits = hold(torch.empty((B, CHORD_SIZE), device=dev, dtype=torch.float32))
    if want_logits:
        st.logits_all = torch.zeros((B, cap, CHORD_SIZE), device=dev, dtype=torch.float32)
        d.logits_all = ptr(st.logits_all)
    st.params, st.gen = d, gen
    st.launches_per_step = int(load().v2m_decode_launches_per_step(C.byref(d)))
    if mode == "stream":
        # fragment-packed copies of the matrices for the streamed cluster kernel (cached with the bf16 weights)
        def packed(name, rows=None, cols=None):
            key = "%s|%s|%s|frag" % (name, rows, cols)
            t = W._cache.get(key)
            if t is None:
                src = W.w(name, rows=rows, cols=cols)
                t = pack_fragments(src[:, :src.shape[1] // 16 * 16] if src.shape[1] % 16 else src)
                W._cache[key] = t
            return hold(t)
        dp = Decode()
        C.memmove(C.byref(dp), C.byref(d), C.sizeof(Decode))
        for l in range(NL):
            p = "transformer.decoder.layers.%d." % l
            L = dp.layer[l]
            L.w_qkv = packed(p + "self_attn.in_proj_weight")
            L.w_so = packed(p + "self_attn.out_proj.weight")
            L.w_cq = packed(p + "multihead_attn.in_proj_weight", rows=slice(0, E))
            L.w_co = packed(p + "multihead_attn.out_proj.weight")
            L.w_f1 = packed(p + "linear1.weight")
            L.w_f2 = packed(p + "linear2.weight")
        dp.w_out = packed("Wout.weight")
        b_out = W.f("Wout.bias")
        dp.b_out = hold(torch.cat([b_out, b_out.new_zeros((-b_out.numel()) % 16)]))     # tiles of 16 rows read whole
        dp.w_chord = packed("Linear_chord.weight", cols=E)
        st.packed = dp
    return st


def run_decode(st: DecodeState, n_steps: int, use_graph: bool = True, mode: str = "auto", n_split: int = 0,
               timestamps: Optional[torch.Tensor] = None) -> None:
    """Advance the generation by n_steps positions.

    mode "stream": the whole loop as ONE persistent thread-block-cluster kernel (bf16, d_model 512, 8 heads,
                   csrc/decode_stream.cu);
    mode "kernels": 70 kernels per position, replayed from a CUDA graph when use_graph (csrc/decode.cu, also the fp32 path);
    mode "auto": whatever build_decode chose.  The caches are laid out for one path, so mode must match build_decode's.
    n_split: independent sub-batch chains inside the graph (0 = pick from the batch size)."""
    d = st.params
    if mode == "auto":
        mode = st.mode
    if mode != st.mode:
        raise RuntimeError("decode state was built for mode %r, not %r" % (st.mode, mode))
    if n_split <= 0:
        n_split = 2 if d.B >= 32 else 1
    if mode == "stream":
        if st.packed is None:
            raise RuntimeError("streamed decode needs bf16, d_model 512, 8 heads, dim_feedforward 1024")
        tsp = ptr(timestamps) if timestamps is not None else None
        check(load().v2m_decode_run_stream(C.byref(st.packed), st.pos, n_steps, tsp,
                                           0 if timestamps is None else timestamps.numel(), stream()))
        _lib.count_launches(1)
    else:
        split = n_split if use_graph else 0
        check(load().v2m_decode_run(C.byref(d), n_steps, split if use_graph else 0, stream()))
        _lib.count_launches(n_steps * st.launches_per_step * max(1, split))
    st.pos += n_steps
    st.step.fill_(st.pos)


def probe_decode_kernel(st: DecodeState, kind: int, reps: int) -> None:
    """Measurement aid for bench.py: `reps` rounds of one decode kernel kind over all layers (see v2m_decode_probe)."""
    check(load().v2m_decode_probe(C.byref(st.params), kind, reps, stream()))
    _lib.count_launches(reps * st.params.n_layers)
