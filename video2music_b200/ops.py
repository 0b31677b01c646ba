"""Thin functional wrappers: torch CUDA tensors in, kernels of libv2m_b200.so out.

torch is used for device memory and the current stream only; every arithmetic
op below is one of our kernels.
"""
import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import BF16, F32, Attn, Epilogue, check, dtype_code, load, ptr, require_device, stream


def _rows(t: torch.Tensor) -> int:
    return t.numel() // t.shape[-1]


# Optional device counter (uint32 viewed as int32 tensor of one element) that every dropout kernel adds to its seed: set by the
# trainer when the training step is replayed from a CUDA graph, where the host-side seeds are frozen into the graph.
DROP_SEED_DEV: Optional[torch.Tensor] = None


def drop_args(p: float, seed: int):
    """(scale, threshold, seed) of the stateless dropout mask shared by the forward and backward kernels: one hash byte per
    element, kept iff byte >= round(p * 256); the drop probability is therefore quantised to 1/256 (0.1 -> 26/256) and the
    kept values are scaled by 256 / (256 - threshold)."""
    assert 0.0 < p < 1.0
    t = min(max(int(round(p * 256.0)), 1), 255)
    return 256.0 / (256.0 - t), t, int(seed) & 0xFFFFFFFF


_drop_state = {"calls": 0, "fixed": None}


def next_dropout_seed() -> int:
    """Seed of the next dropout site of a module forward: a new mask per site and per call, reproducible under
    torch.manual_seed (the draws are a stateless hash, not torch's Philox stream), different on every rank."""
    import torch.distributed as dist
    st = _drop_state
    st["calls"] += 1
    base = st["fixed"] if st["fixed"] is not None else torch.initial_seed()
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    return (base * 1000003 + st["calls"] * 7919 + rank * 104729) & 0xFFFFFFFF


class fixed_dropout_seed:
    """with fixed_dropout_seed(s): ...  -- the dropout sites inside draw seeds from (s, site index), the site index
    restarting at every entry: two forwards under the same `s` see identical masks (tests, finite differences)."""

    def __init__(self, seed: int):
        self.seed = int(seed)

    def __enter__(self):
        self.prev = dict(_drop_state)
        _drop_state["fixed"], _drop_state["calls"] = self.seed, 0
        return self

    def __exit__(self, *exc):
        _drop_state.update(self.prev)
        return False


def dropout(x: torch.Tensor, p: float, seed: int) -> torch.Tensor:
    """Inverted dropout of a 2-D fp32 / bf16 tensor: element (r, c) is kept iff drop_keep(seed, r, c) (the mask of the GEMM
    epilogues and of dy_prep, see drop_args) and scaled by 256 / (256 - round(256 p)).  The same call on the incoming gradient
    is the backward.  nn.Dropout in training mode (moe.py:48,197; bimamba.py:50-98; video_regression.py:203)."""
    if p <= 0.0:
        return x
    assert x.dim() == 2
    x = x if x.stride(1) == 1 else x.contiguous()
    dz, _ = dy_prep(x, None, False, 1.0, 0, x.dtype, want_dz=True, dropout=(p, seed))
    return dz


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, k: Optional[int] = None,
           relu: bool = False, alpha: float = 1.0, alpha_cols: int = 0,
           residual: Optional[torch.Tensor] = None, res_mod: int = 0,
           row_scale: Optional[torch.Tensor] = None, col_vec: Optional[torch.Tensor] = None,
           out_dtype: Optional[torch.dtype] = None, out: Optional[torch.Tensor] = None,
           head_scatter: Optional[dict] = None, dropout: Optional[tuple] = None) -> torch.Tensor:
    """y = epilogue(x @ w[:, :k].T).  x (M, >=k) and w (N, >=k) are row-major 2-D tensors of the same
    dtype (fp32 -> SIMT exact GEMM, bf16 -> tcgen05 GEMM); leading dims are taken from the strides.
    dropout = (p, seed, after_residual): inverted dropout fused into the epilogue (both paths), see drop_args()."""
    require_device(x)
    assert x.dim() == 2 and w.dim() == 2 and x.stride(1) == 1 and w.stride(1) == 1
    assert x.dtype == w.dtype, (x.dtype, w.dtype)
    M, N = x.shape[0], w.shape[0]
    K = k if k is not None else min(x.shape[1], w.shape[1])
    lda, ldw = x.stride(0), w.stride(0)
    if out_dtype is None:
        out_dtype = x.dtype
    ep = Epilogue()
    ep.bias = ptr(bias)
    ep.alpha, ep.alpha_cols, ep.relu = alpha, alpha_cols, int(relu)
    if residual is not None:
        assert residual.stride(-1) == 1
        ep.residual, ep.ldr, ep.res_mod = ptr(residual), residual.stride(-2), res_mod
        ep.residual_bf16 = int(residual.dtype == torch.bfloat16)
    if row_scale is not None:
        ep.row_scale, ep.col_vec = ptr(row_scale), ptr(col_vec)
    if dropout is not None and dropout[0] > 0.0:
        ep.drop_scale, ep.drop_thresh, ep.drop_seed = drop_args(dropout[0], dropout[1])
        ep.drop_after_res = int(bool(dropout[2]))
        ep.drop_seed_dev = ptr(DROP_SEED_DEV)
    if head_scatter is not None:
        assert out is not None
        ep.head_scatter = 1
        for f in ("S", "H", "dh", "cap", "pos0", "part_stride"):
            setattr(ep, f, head_scatter[f])
        ldc = 0
    else:
        if out is None:
            out = torch.empty((M, N), device=x.device, dtype=out_dtype)
        ldc = out.stride(0)
    lib = load()
    if x.dtype == torch.float32:
        assert out.dtype == torch.float32
        check(lib.v2m_gemm_f32(ptr(x), lda, ptr(w), ldw, ptr(out), ldc, M, N, K, C.byref(ep), stream()))
    else:
        check(lib.v2m_gemm_bf16(ptr(x), lda, ptr(w), ldw, ptr(out), ldc, dtype_code(out.dtype), M, N, K,
                                C.byref(ep), stream()))
    _lib.count_launches(1)
    return out


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, *, B: int, Hq: int, Hkv: int,
              Lq: int, Lk: int, dh: int, q_strides, k_strides, v_strides, o_strides, causal: bool,
              Er: Optional[torch.Tensor] = None, q_scale: float = 1.0, lse: Optional[torch.Tensor] = None,
              p_out: Optional[torch.Tensor] = None, dropout: Optional[tuple] = None,
              lk_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """softmax(q k^T + skew(q Er^T) + causal) v.  *_strides = (batch stride, row stride) in elements;
    q/k/v/out may be column slices of wider matrices (head h starts at column h*dh of the given pointer).
    lk_dev (int32 device scalar, fp32 path): only the first min(Lk, lk_dev) keys exist (graph-replayed cached decode)."""
    require_device(q)
    a = Attn()
    a.q, a.k, a.v, a.o = ptr(q), ptr(k), ptr(v), ptr(out)
    a.q_sb, a.q_sl = q_strides
    a.k_sb, a.k_sl = k_strides
    a.v_sb, a.v_sl = v_strides
    a.o_sb, a.o_sl = o_strides
    a.B, a.Hq, a.Hkv, a.Lq, a.Lk, a.dh = B, Hq, Hkv, Lq, Lk, dh
    a.causal = int(causal)
    if Er is not None:
        assert Er.dtype == q.dtype and Er.is_contiguous()
        a.Er, a.er_len = ptr(Er), Er.shape[0]
    a.q_scale = q_scale
    a.lse, a.p_out = ptr(lse), ptr(p_out)
    if dropout is not None and dropout[0] > 0.0:                 # (p, seed): dropout of the probabilities, bf16 path
        a.drop_scale, a.drop_thresh, a.drop_seed = drop_args(dropout[0], dropout[1])
        a.drop_seed_dev = ptr(DROP_SEED_DEV)
    if lk_dev is not None:
        assert lk_dev.dtype == torch.int32 and lk_dev.is_cuda and q.dtype == torch.float32
        a.lk_dev = ptr(lk_dev)
    check(load().v2m_attn_fwd(C.byref(a), dtype_code(q.dtype), stream()))
    _lib.count_launches(1)
    return out


def step_linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, k: Optional[int] = None, relu: bool = False,
                row_scale: Optional[torch.Tensor] = None, col_vec: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 y = act(x @ w[:, :k].T + bias + row_scale[:, None] * col_vec[None, :]) for the few rows of one generation step
    (csrc/step_f32.cu: a weight stream over the whole chip instead of one 128-row GEMM tile)."""
    require_device(x)
    assert x.dim() == 2 and w.dim() == 2 and x.stride(1) == 1 and w.stride(1) == 1 and x.dtype == w.dtype == torch.float32
    M, N = x.shape[0], w.shape[0]
    K = k if k is not None else min(x.shape[1], w.shape[1])
    y = torch.empty((M, N), device=x.device, dtype=torch.float32)
    check(load().v2m_step_linear_f32(ptr(x), x.stride(0), ptr(w), w.stride(0), ptr(bias), ptr(row_scale), ptr(col_vec), ptr(y), N,
                                     M, N, K, int(relu), stream()))
    _lib.count_launches(1)
    return y


def step_attention(q: torch.Tensor, K: torch.Tensor, V: torch.Tensor, *, Hq: int, Hkv: int, dh: int, n_max: int, kv_strides,
                   n_dev: Optional[torch.Tensor] = None, q_scale: float = 1.0) -> torch.Tensor:
    """One query row per (video, query head) over the first n cached rows: q (B, Hq*dh) fp32, K / V caches addressed as
    [b*kv_strides[0] + j*kv_strides[1] + h*dh + d]; n = min(n_max, n_dev) with n_dev an int32 device scalar (optional)."""
    require_device(q)
    assert q.dtype == K.dtype == V.dtype == torch.float32 and q.dim() == 2 and q.stride(1) == 1
    B = q.shape[0]
    out = torch.empty((B, Hq * dh), device=q.device, dtype=torch.float32)
    if n_dev is not None:
        assert n_dev.dtype == torch.int32 and n_dev.is_cuda
    check(load().v2m_step_attn_f32(ptr(q), q.stride(0), ptr(K), ptr(V), kv_strides[0], kv_strides[1], ptr(out), Hq * dh, B, Hq, Hkv, dh,
                                   n_max, ptr(n_dev), q_scale, stream()))
    _lib.count_launches(1)
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, *, res: Optional[torch.Tensor] = None,
              out_dtype: Optional[torch.dtype] = None, eps: float = 1e-5) -> torch.Tensor:
    require_device(x)
    assert x.is_contiguous()
    D = x.shape[-1]
    y = torch.empty(x.shape, device=x.device, dtype=out_dtype or x.dtype)
    check(load().v2m_layernorm(ptr(x), dtype_code(x.dtype), ptr(res), dtype_code(res.dtype) if res is not None else 0,
                               ptr(gamma), ptr(beta), ptr(y), dtype_code(y.dtype), None, 0, _rows(x), D, eps, stream()))
    _lib.count_launches(1)
    return y


def embed_sum(idx_a: torch.Tensor, table_a: torch.Tensor, idx_b: Optional[torch.Tensor],
              table_b: Optional[torch.Tensor], out_dtype: torch.dtype) -> torch.Tensor:
    require_device(table_a)
    rows, D = idx_a.numel(), table_a.shape[1]
    idx_a = idx_a.contiguous()
    if idx_b is not None:
        idx_b = idx_b.contiguous()
    out = torch.empty((rows, D), device=table_a.device, dtype=out_dtype)
    check(load().v2m_embed_sum(ptr(idx_a), ptr(table_a), ptr(idx_b), ptr(table_b), ptr(out), dtype_code(out_dtype), D,
                               rows, D, stream()))
    _lib.count_launches(1)
    return out


def concat_features(sem: torch.Tensor, scene: torch.Tensor, motion: torch.Tensor, emotion: torch.Tensor,
                    out_dtype: torch.dtype, ld_out: int) -> torch.Tensor:
    """[semantic | scene offset | motion | emotion | zero pad] per (video, second), video_music_transformer.py:1003-1018."""
    require_device(sem)
    rows = sem.shape[0] * sem.shape[1]
    sem, scene, motion, emotion = (t.float().contiguous() for t in (sem, scene, motion, emotion))
    motion_dim = 1 if motion.dim() == 2 else motion.shape[-1]
    out = torch.empty((rows, ld_out), device=sem.device, dtype=out_dtype)
    check(load().v2m_concat_features(ptr(sem), sem.shape[-1], ptr(scene), ptr(motion), motion_dim, ptr(emotion),
                                     emotion.shape[-1], ptr(out), dtype_code(out_dtype), ld_out, rows, stream()))
    _lib.count_launches(1)
    return out


def cast_2d(src: torch.Tensor, dst_dtype: torch.dtype, ld_dst: Optional[int] = None) -> torch.Tensor:
    """Row-major 2-D copy with dtype conversion and optional zero padding of the leading dimension."""
    require_device(src)
    assert src.dim() == 2 and src.stride(1) == 1
    rows, cols = src.shape
    ld = ld_dst or cols
    dst = torch.empty((rows, ld), device=src.device, dtype=dst_dtype)
    check(load().v2m_cast_2d(ptr(src), dtype_code(src.dtype), src.stride(0), ptr(dst), dtype_code(dst_dtype), ld, rows,
                             cols, int(ld != cols), stream()))
    _lib.count_launches(1)
    return dst


def rope_quirk(x: torch.Tensor, cache: torch.Tensor, B: int, H: int) -> torch.Tensor:
    """RoPE of the reference's V2 attention on a (len*B, E) fp32 projection (rows ordered (l, b)); cache = the first len
    positions of RotaryPositionalEmbeddings.cache, contiguous [len, E/2, 2]."""
    require_device(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and cache.dtype == torch.float32 and cache.is_contiguous()
    E = x.shape[1]
    length = x.shape[0] // B
    # a wider cache (RotaryPositionalEmbeddings(2 * d_model) feeding a d_model-wide projection in the V3 encoder) is viewed the
    # same way by the reference ([-1, len, 1, dh/2, 2], first H slabs): only its first H * len * dh/2 entries are used
    assert cache.shape[0] == length and cache.shape[1] * 2 >= E
    y = torch.empty_like(x)
    check(load().v2m_rope_quirk(ptr(x), ptr(cache), ptr(y), length, B, H, E // H, stream()))
    _lib.count_launches(1)
    return y


def cast_into(src: torch.Tensor, dst: torch.Tensor) -> None:
    """dst[r, c] = src[r, c] converted to dst's dtype (same 2-D shape, unit inner strides)."""
    require_device(src)
    assert src.dim() == 2 and dst.shape == src.shape and src.stride(1) == 1 and dst.stride(1) == 1
    check(load().v2m_cast_2d(ptr(src), dtype_code(src.dtype), src.stride(0), ptr(dst), dtype_code(dst.dtype), dst.stride(0), src.shape[0],
                             src.shape[1], 0, stream()))
    _lib.count_launches(1)


def _binary(a: torch.Tensor, b: torch.Tensor, mode: int, alpha: float) -> torch.Tensor:
    require_device(a)
    a, b = a.contiguous(), b.contiguous()
    assert a.shape == b.shape and a.dtype == b.dtype == torch.float32
    out = torch.empty_like(a)
    check(load().v2m_binary_f32(ptr(a), ptr(b), ptr(out), a.numel(), mode, alpha, stream()))
    _lib.count_launches(1)
    return out


def swiglu(a: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """a * silu(g)  (moe.py:47)."""
    return _binary(a, g, 0, 0.0)


def sigmoid(a: torch.Tensor) -> torch.Tensor:
    """1 / (1 + exp(-a))  (video_regression.py:197-200)."""
    return _binary(a, a, 2, 0.0)


def sigmoid_bwd(dy: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    """dy * s * (1 - s): gradient of sigmoid given its output."""
    return _binary(dy, s, 3, 0.0)


def axpy(a: torch.Tensor, b: torch.Tensor, alpha: float) -> torch.Tensor:
    """a + alpha * b  (moe.py:301)."""
    return _binary(a, b, 1, alpha)


def pscan_fwd(A: torch.Tensor, X: torch.Tensor) -> torch.Tensor:
    require_device(A)
    A, X = A.contiguous(), X.contiguous()
    B, L, D, N = A.shape
    H = torch.empty_like(X)
    check(load().v2m_pscan_fwd(ptr(A), ptr(X), ptr(H), B, L, D, N, stream()))
    _lib.count_launches(1)
    return H


def pscan_bwd(A: torch.Tensor, H: torch.Tensor, gH: torch.Tensor):
    require_device(A)
    A, H, gH = A.contiguous(), H.contiguous(), gH.contiguous()
    B, L, D, N = A.shape
    gA, gX = torch.empty_like(A), torch.empty_like(A)
    check(load().v2m_pscan_bwd(ptr(A), ptr(H), ptr(gH), ptr(gA), ptr(gX), B, L, D, N, stream()))
    _lib.count_launches(1)
    return gA, gX


def moe_route(x: torch.Tensor, wg: torch.Tensor, bg: torch.Tensor, k: int, *, sel_bias: Optional[torch.Tensor] = None,
              inv_t_pre: float = 1.0, inv_t_post: float = 1.0, want_logits: bool = False):
    """x (..., d) fp32 -> (selected_experts int64 (..., k), weights fp32 (..., k), histogram int32 (E,), logits?)."""
    require_device(x)
    x = x.contiguous()
    d = x.shape[-1]
    tokens = x.numel() // d
    E = wg.shape[0]
    idx = torch.empty(x.shape[:-1] + (k,), device=x.device, dtype=torch.int64)
    w = torch.empty(x.shape[:-1] + (k,), device=x.device, dtype=torch.float32)
    hist = torch.zeros((E,), device=x.device, dtype=torch.int32)
    logits = torch.empty(x.shape[:-1] + (E,), device=x.device, dtype=torch.float32) if want_logits else None
    check(load().v2m_moe_route(ptr(x), ptr(wg.contiguous()), ptr(bg.contiguous()), ptr(sel_bias), inv_t_pre, inv_t_post,
                               tokens, d, E, k, ptr(idx), ptr(w), ptr(logits), ptr(hist), stream()))
    _lib.count_launches(1)
    return idx, w, hist, logits


def moe_experts(x: torch.Tensor, idx: torch.Tensor, w: torch.Tensor, hist: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor,
                wg: torch.Tensor, bg: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor) -> torch.Tensor:
    """out[t] = sum_r w[t,r] * GLUExpert_{idx[t,r]}(x[t])  (model/moe.py:191-199) with stacked expert weights
    w1/wg [E, ff, d], w2 [E, d_out, ff]: permute -> grouped (x W1^T + b1) * silu(x Wg^T + bg) -> grouped linear2 -> combine.
    Five launches, group sizes stay on the device (no host sync)."""
    require_device(x)
    T, k = idx.shape
    E, ff, d = w1.shape
    d_out = w2.shape[1]
    dev = x.device
    meta = torch.empty((2 * E + 1 + T * k,), device=dev, dtype=torch.int32)
    off, cursor, perm = meta[:E + 1], meta[E + 1:2 * E + 1], meta[2 * E + 1:]
    xp = torch.empty((T * k, d), device=dev, dtype=torch.float32)
    h = torch.empty((T * k, ff), device=dev, dtype=torch.float32)
    yp = torch.empty((T * k, d_out), device=dev, dtype=torch.float32)
    out = torch.empty((T, d_out), device=dev, dtype=torch.float32)
    lib, st = load(), stream()
    check(lib.v2m_moe_permute(ptr(x), ptr(idx), ptr(hist), T, k, d, E, 1, ptr(off), ptr(cursor), ptr(xp), dtype_code(xp.dtype), ptr(perm),
                              None, 0, st))
    check(lib.v2m_moe_grouped_gemm(ptr(xp), d, ptr(w1), ptr(b1), ptr(wg), ptr(bg), ff * d, ff, ptr(off), E, T * k, ptr(h), ff, ff, d, st))
    check(lib.v2m_moe_grouped_gemm(ptr(h), ff, ptr(w2), ptr(b2), None, None, d_out * ff, d_out, ptr(off), E, T * k, ptr(yp), d_out,
                                   d_out, ff, st))
    check(lib.v2m_moe_combine(ptr(yp), ptr(perm), ptr(w), ptr(out), T, k, d_out, st))
    _lib.count_launches(5)
    return out


def moe_experts_bf16(x: torch.Tensor, idx: torch.Tensor, w: torch.Tensor, hist: torch.Tensor, w1g: torch.Tensor, b1g: torch.Tensor,
                     w2: torch.Tensor, b2: torch.Tensor) -> torch.Tensor:
    """Tensor-core expert path: x fp32 (T, d) is permuted into bf16 expert-contiguous rows with 128-row aligned groups, the
    experts run as two grouped tcgen05 GEMMs (stacked bf16 weights w1g [E, 2 ff, d] = linear1 | gate, w2 [E, d_out, ff];
    fp32 biases) with the SwiGLU between them, and the fp32 results are combined with the fp32 router weights."""
    require_device(x)
    T, k = idx.shape
    E, ff2, d = w1g.shape
    ff, d_out = ff2 // 2, w2.shape[1]
    dev = x.device
    m_cap = (T * k + E * 127 + 127) // 128 * 128
    n_tiles = m_cap // 128
    meta = torch.empty((2 * E + 1 + T * k + n_tiles,), device=dev, dtype=torch.int32)
    off, cursor, perm, tile_group = meta[:E + 1], meta[E + 1:2 * E + 1], meta[2 * E + 1:2 * E + 1 + T * k], meta[2 * E + 1 + T * k:]
    xp = torch.zeros((m_cap, d), device=dev, dtype=torch.bfloat16)
    a = torch.empty((m_cap, ff2), device=dev, dtype=torch.bfloat16)
    h = torch.empty((m_cap, ff), device=dev, dtype=torch.bfloat16)
    yp = torch.empty((m_cap, d_out), device=dev, dtype=torch.float32)
    out = torch.empty((T, d_out), device=dev, dtype=torch.float32)
    lib, st = load(), stream()
    check(lib.v2m_moe_permute(ptr(x), ptr(idx), ptr(hist), T, k, d, E, 128, ptr(off), ptr(cursor), ptr(xp), dtype_code(xp.dtype), ptr(perm),
                              ptr(tile_group), n_tiles, st))
    check(lib.v2m_gemm_bf16_grouped(ptr(xp), d, ptr(w1g), d, ptr(a), ff2, dtype_code(a.dtype), m_cap, ff2, d, E, ptr(tile_group), ptr(b1g), 0, st))
    check(lib.v2m_swiglu_pair_bf16(ptr(a), ptr(h), m_cap, ff, st))
    check(lib.v2m_gemm_bf16_grouped(ptr(h), ff, ptr(w2), ff, ptr(yp), d_out, dtype_code(yp.dtype), m_cap, d_out, ff, E, ptr(tile_group), ptr(b2), 0, st))
    check(lib.v2m_moe_combine(ptr(yp), ptr(perm), ptr(w), ptr(out), T, k, d_out, st))
    _lib.count_launches(7)
    return out


def _moe_bf16_forward(x, idx, w, hist, w1g, b1g, w2, b2):
    T, k = idx.shape
    E, ff2, d = w1g.shape
    ff, d_out = ff2 // 2, w2.shape[1]
    dev = x.device
    m_cap = (T * k + E * 127 + 127) // 128 * 128
    n_tiles = m_cap // 128
    meta = torch.empty((2 * E + 1 + T * k + n_tiles,), device=dev, dtype=torch.int32)
    off, cursor, perm, tile_group = meta[:E + 1], meta[E + 1:2 * E + 1], meta[2 * E + 1:2 * E + 1 + T * k], meta[2 * E + 1 + T * k:]
    xp = torch.zeros((m_cap, d), device=dev, dtype=torch.bfloat16)
    a = torch.empty((m_cap, ff2), device=dev, dtype=torch.bfloat16)
    h = torch.empty((m_cap, ff), device=dev, dtype=torch.bfloat16)
    yp = torch.empty((m_cap, d_out), device=dev, dtype=torch.float32)
    out = torch.empty((T, d_out), device=dev, dtype=torch.float32)
    lib, st = load(), stream()
    check(lib.v2m_moe_permute(ptr(x), ptr(idx), ptr(hist), T, k, d, E, 128, ptr(off), ptr(cursor), ptr(xp), dtype_code(xp.dtype), ptr(perm),
                              ptr(tile_group), n_tiles, st))
    check(lib.v2m_gemm_bf16_grouped(ptr(xp), d, ptr(w1g), d, ptr(a), ff2, dtype_code(a.dtype), m_cap, ff2, d, E, ptr(tile_group), ptr(b1g), 0, st))
    check(lib.v2m_swiglu_pair_bf16(ptr(a), ptr(h), m_cap, ff, st))
    check(lib.v2m_gemm_bf16_grouped(ptr(h), ff, ptr(w2), ff, ptr(yp), d_out, dtype_code(yp.dtype), m_cap, d_out, ff, E, ptr(tile_group), ptr(b2), 0, st))
    check(lib.v2m_moe_combine(ptr(yp), ptr(perm), ptr(w), ptr(out), T, k, d_out, st))
    _lib.count_launches(7)
    return out, (xp, a, h, yp, perm, off, tile_group)


def moe_experts_bf16_fwd_saved(x: torch.Tensor, idx: torch.Tensor, w: torch.Tensor, hist: torch.Tensor, w1g: torch.Tensor,
                               b1g: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor):
    """Training-mode forward of `moe_experts_bf16` (same launches): returns (out, saved) with saved = (xp bf16 (m_cap, d), a bf16
    (m_cap, 2 ff) = [linear1 | gate] pre-activations, h bf16, yp fp32, perm, off (128-row aligned group starts), tile_group)."""
    require_device(x)
    return _moe_bf16_forward(x, idx, w, hist, w1g, b1g, w2, b2)


def moe_experts_bf16_bwd(dout: torch.Tensor, saved, idx: torch.Tensor, w: torch.Tensor, scale: float, w1g_t: torch.Tensor,
                         w2_t: torch.Tensor, n_experts: int):
    """Backward of the bf16 expert path on the tensor cores.  w1g_t bf16 [E, d, 2 ff] and w2_t bf16 [E, ff, d_out] are the
    transposed weight stacks (the dX products are grouped GEMMs against them); the ragged weight gradients dW_e = dY_e^T X_e
    are K-grouped tcgen05 GEMMs with the group bounds read on the device.  Returns (dx_experts fp32 (T, d), dlogits (T, E),
    dW1g fp32 (E, 2 ff, d), db1g (E, 2 ff), dW2 (E, d_out, ff), db2 (E, d_out))."""
    require_device(dout)
    xp, a, h, yp, perm, off, tile_group = saved
    T, k = idx.shape
    E = n_experts
    m_cap, d = xp.shape
    ff, d_out = h.shape[1], yp.shape[1]
    dev = dout.device
    f32 = dict(device=dev, dtype=torch.float32)
    bf = dict(device=dev, dtype=torch.bfloat16)
    dout = dout.float().contiguous()
    dyp32, dlogits = torch.zeros((m_cap, d_out), **f32), torch.empty((T, E), **f32)      # padding rows of every group stay zero
    lib, st = load(), stream()
    check(lib.v2m_moe_combine_bwd(ptr(dout), ptr(yp), ptr(perm), ptr(w), ptr(idx), scale, T, k, d_out, E, ptr(dyp32), ptr(dlogits), st))
    _lib.count_launches(1)
    dyp = cast_2d(dyp32, torch.bfloat16)
    dW2, db2 = torch.empty((E, d_out, ff), **f32), torch.empty((E, d_out), **f32)
    dW1g, db1g = torch.empty((E, 2 * ff, d), **f32), torch.empty((E, 2 * ff), **f32)
    dh, dag = torch.empty((m_cap, ff), **bf), torch.empty((m_cap, 2 * ff), **bf)
    dxp, dx = torch.empty((m_cap, d), **f32), torch.empty((T, d), **f32)
    ones = torch.ones((T, k), **f32)
    check(lib.v2m_gemm_bf16_kgrouped(ptr(dyp), dyp.stride(0), ptr(h), ff, ptr(dW2), ff, d_out * ff, d_out, ff, m_cap, E, ptr(off), st))
    check(lib.v2m_moe_group_colsum_bf16(ptr(dyp), dyp.stride(0), ptr(off), E, ptr(db2), d_out, m_cap, st))
    check(lib.v2m_gemm_bf16_grouped(ptr(dyp), dyp.stride(0), ptr(w2_t), d_out, ptr(dh), ff, dtype_code(dh.dtype), m_cap, ff, d_out, E,
                                    ptr(tile_group), None, 0, st))
    check(lib.v2m_swiglu_pair_bwd_bf16(ptr(a), ptr(dh), ptr(dag), m_cap, ff, st))
    check(lib.v2m_gemm_bf16_kgrouped(ptr(dag), 2 * ff, ptr(xp), d, ptr(dW1g), d, 2 * ff * d, 2 * ff, d, m_cap, E, ptr(off), st))
    check(lib.v2m_moe_group_colsum_bf16(ptr(dag), 2 * ff, ptr(off), E, ptr(db1g), 2 * ff, m_cap, st))
    check(lib.v2m_gemm_bf16_grouped(ptr(dag), 2 * ff, ptr(w1g_t), 2 * ff, ptr(dxp), d, dtype_code(dxp.dtype), m_cap, d, 2 * ff, E,
                                    ptr(tile_group), None, 0, st))
    check(lib.v2m_moe_combine(ptr(dxp), ptr(perm), ptr(ones), ptr(dx), T, k, d, st))
    _lib.count_launches(12)
    return dx, dlogits, dW1g, db1g, dW2, db2


def swiglu_pair_bwd(a: torch.Tensor, dh: torch.Tensor) -> torch.Tensor:
    """Gradient of h = a[:, :ff] * silu(a[:, ff:]) w.r.t. the bf16 pair matrix a (M, 2 ff): (M, 2 ff) bf16."""
    require_device(a)
    assert a.dtype == dh.dtype == torch.bfloat16 and a.is_contiguous() and dh.is_contiguous()
    M, ff = dh.shape
    dag = torch.empty_like(a)
    check(load().v2m_swiglu_pair_bwd_bf16(ptr(a), ptr(dh), ptr(dag), M, ff, stream()))
    _lib.count_launches(1)
    return dag


def moe_experts_fwd_saved(x: torch.Tensor, idx: torch.Tensor, w: torch.Tensor, hist: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor,
                          wg: torch.Tensor, bg: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor, drops=None):
    """Training-mode forward of `moe_experts`: same result, but the two halves of the SwiGLU pair are kept (a = x W1^T + b1,
    g = x Wg^T + bg) together with the permuted input, the hidden rows and the per-row expert outputs, which the backward
    needs.  Returns (out, saved) with saved = (xp, a, g, h, yp, perm, off).
    drops = (p_hidden, seed_hidden, p_out, seed_out): GLUExpert's dropout of the hidden rows (moe.py:48) and the layer's dropout
    of each expert output row (moe.py:197) in training mode, masks over (permuted row, column); h and yp are saved dropped."""
    require_device(x)
    T, k = idx.shape
    E, ff, d = w1.shape
    d_out = w2.shape[1]
    dev = x.device
    meta = torch.empty((2 * E + 1 + T * k,), device=dev, dtype=torch.int32)
    off, cursor, perm = meta[:E + 1], meta[E + 1:2 * E + 1], meta[2 * E + 1:]
    f32 = dict(device=dev, dtype=torch.float32)
    xp, a, g = torch.empty((T * k, d), **f32), torch.empty((T * k, ff), **f32), torch.empty((T * k, ff), **f32)
    yp, out = torch.empty((T * k, d_out), **f32), torch.empty((T, d_out), **f32)
    lib, st = load(), stream()
    check(lib.v2m_moe_permute(ptr(x), ptr(idx), ptr(hist), T, k, d, E, 1, ptr(off), ptr(cursor), ptr(xp), dtype_code(xp.dtype), ptr(perm),
                              None, 0, st))
    check(lib.v2m_moe_grouped_gemm(ptr(xp), d, ptr(w1), ptr(b1), None, None, ff * d, ff, ptr(off), E, T * k, ptr(a), ff, ff, d, st))
    check(lib.v2m_moe_grouped_gemm(ptr(xp), d, ptr(wg), ptr(bg), None, None, ff * d, ff, ptr(off), E, T * k, ptr(g), ff, ff, d, st))
    _lib.count_launches(4)
    h = swiglu(a, g)
    if drops is not None and drops[0] > 0.0:
        h = dropout(h, drops[0], drops[1])
    check(lib.v2m_moe_grouped_gemm(ptr(h), ff, ptr(w2), ptr(b2), None, None, d_out * ff, d_out, ptr(off), E, T * k, ptr(yp), d_out,
                                   d_out, ff, st))
    if drops is not None and drops[2] > 0.0:
        yp = dropout(yp, drops[2], drops[3])
    check(lib.v2m_moe_combine(ptr(yp), ptr(perm), ptr(w), ptr(out), T, k, d_out, st))
    _lib.count_launches(2)
    return out, (xp, a, g, h, yp, perm, off)


def moe_experts_bwd(dout: torch.Tensor, saved, idx: torch.Tensor, w: torch.Tensor, scale: float, w1g_t: torch.Tensor,
                    w2_t: torch.Tensor, n_experts: int, drops=None):
    """Backward of `moe_experts_fwd_saved`.  w1g_t [E, d, 2 ff] = (linear1 | gate) weights transposed, w2_t [E, ff, d_out].
    Returns (dx_experts (T, d), dlogits (T, E), dW1g (E, 2 ff, d), db1g (E, 2 ff), dW2 (E, d_out, ff), db2 (E, d_out))."""
    require_device(dout)
    xp, a, g, h, yp, perm, off = saved
    T, k = idx.shape
    E = n_experts
    M, d = xp.shape
    ff, d_out = a.shape[1], yp.shape[1]
    f32 = dict(device=dout.device, dtype=torch.float32)
    dout = dout.contiguous()
    dyp, dlogits = torch.empty((M, d_out), **f32), torch.empty((T, E), **f32)
    dh, dag, dxp, dx = torch.empty((M, ff), **f32), torch.empty((M, 2 * ff), **f32), torch.empty((M, d), **f32), torch.empty((T, d), **f32)
    dW2, db2 = torch.empty((E, d_out, ff), **f32), torch.empty((E, d_out), **f32)
    dW1g, db1g = torch.empty((E, 2 * ff, d), **f32), torch.empty((E, 2 * ff), **f32)
    ones = torch.ones((T, k), **f32)
    lib, st = load(), stream()
    check(lib.v2m_moe_combine_bwd(ptr(dout), ptr(yp), ptr(perm), ptr(w), ptr(idx), scale, T, k, d_out, E, ptr(dyp), ptr(dlogits), st))
    if drops is not None and drops[2] > 0.0:                      # gradient of the dropped expert output -> of the GEMM output
        dyp = dropout(dyp, drops[2], drops[3])
    check(lib.v2m_moe_grouped_dw(ptr(dyp), d_out, ptr(h), ff, ptr(off), E, ptr(dW2), ptr(db2), d_out, ff, st))
    check(lib.v2m_moe_grouped_gemm(ptr(dyp), d_out, ptr(w2_t), None, None, None, ff * d_out, 0, ptr(off), E, M, ptr(dh), ff, ff, d_out, st))
    if drops is not None and drops[0] > 0.0:                      # ... of the dropped hidden rows -> of the SwiGLU output
        dh = dropout(dh, drops[0], drops[1])
    check(lib.v2m_swiglu_bwd(ptr(a), ptr(g), ptr(dh), ptr(dag), M, ff, st))
    check(lib.v2m_moe_grouped_dw(ptr(dag), 2 * ff, ptr(xp), d, ptr(off), E, ptr(dW1g), ptr(db1g), 2 * ff, d, st))
    check(lib.v2m_moe_grouped_gemm(ptr(dag), 2 * ff, ptr(w1g_t), None, None, None, d * 2 * ff, 0, ptr(off), E, M, ptr(dxp), d, d, 2 * ff, st))
    check(lib.v2m_moe_combine(ptr(dxp), ptr(perm), ptr(ones), ptr(dx), T, k, d, st))
    _lib.count_launches(7)
    return dx, dlogits, dW1g, db1g, dW2, db2


def dw_f32(dz: torch.Tensor, x: torch.Tensor, K: int) -> torch.Tensor:
    """fp32 weight gradient dW[N, K] = dz[M, N]^T x[M, :K] (row-major operands with unit inner stride), rows split over the GPU."""
    require_device(dz)
    assert dz.dtype == x.dtype == torch.float32 and dz.stride(1) == 1 and x.stride(1) == 1 and x.shape[1] >= K
    M, N = dz.shape
    dw = torch.empty((N, K), device=dz.device, dtype=torch.float32)
    check(load().v2m_dw_f32(ptr(dz), dz.stride(0), ptr(x), x.stride(0), M, ptr(dw), None, N, K, stream()))
    _lib.count_launches(1)
    return dw


def swiglu_bwd(a: torch.Tensor, g: torch.Tensor, dh: torch.Tensor) -> torch.Tensor:
    """Gradient of h = a * silu(g) as one [M, 2 ff] matrix (da | dg)."""
    require_device(a)
    a, g, dh = a.contiguous(), g.contiguous(), dh.contiguous()
    M, ff = a.shape
    dag = torch.empty((M, 2 * ff), device=a.device, dtype=torch.float32)
    check(load().v2m_swiglu_bwd(ptr(a), ptr(g), ptr(dh), ptr(dag), M, ff, stream()))
    _lib.count_launches(1)
    return dag


def swiglu_pair(a: torch.Tensor) -> torch.Tensor:
    """(M, 2 ff) bf16 -> (M, ff) bf16: a[:, :ff] * silu(a[:, ff:])."""
    require_device(a)
    assert a.dtype == torch.bfloat16 and a.is_contiguous()
    M, ff = a.shape[0], a.shape[1] // 2
    h = torch.empty((M, ff), device=a.device, dtype=torch.bfloat16)
    check(load().v2m_swiglu_pair_bf16(ptr(a), ptr(h), M, ff, stream()))
    _lib.count_launches(1)
    return h


# ----------------------------------------------------------------------------- backward-pass kernels
def gemm_strided(a: torch.Tensor, a_rs: int, a_cs: int, w: torch.Tensor, w_rs: int, w_cs: int, M: int, N: int, K: int,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 C[M,N] = sum_k A(m,k) W(n,k) with A(m,k) = a[m*a_rs + k*a_cs], W(n,k) = w[n*w_rs + k*w_cs]."""
    require_device(a)
    assert a.dtype == w.dtype == torch.float32
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=torch.float32)
    ep = Epilogue()
    check(load().v2m_gemm_f32_strided(ptr(a), a_rs, a_cs, ptr(w), w_rs, w_cs, ptr(out), out.stride(0), M, N, K, C.byref(ep),
                                      stream()))
    _lib.count_launches(1)
    return out


def dy_prep(dy: torch.Tensor, y: Optional[torch.Tensor], relu: bool, alpha: float, alpha_cols: int, out_dtype: torch.dtype,
            want_dz: bool = True, dropout: Optional[tuple] = None, db_out: Optional[torch.Tensor] = None):
    """dz = dy * relu'(y) * alpha_n * dropmask ; db = column sums of dz.  Returns (dz or None, db fp32 [N]).
    dropout = (p, seed) of the forward epilogue whose mask is recomputed here."""
    require_device(dy)
    assert dy.dim() == 2 and dy.stride(1) == 1
    M, N = dy.shape
    ld = N if out_dtype == torch.float32 else (N + 7) // 8 * 8        # TMA operands need a 16-byte row pitch
    dz = None
    if want_dz:
        dz = torch.zeros((M, ld), device=dy.device, dtype=out_dtype) if ld != N else torch.empty((M, N), device=dy.device, dtype=out_dtype)
    # db_out: the column sums are ADDED to this caller-owned fp32 buffer (a bias gradient inside the optimiser's buffer)
    db = torch.zeros((N,), device=dy.device, dtype=torch.float32) if db_out is None else db_out
    assert db.shape == (N,) and db.dtype == torch.float32 and db.is_contiguous()
    check(load().v2m_dy_prep(ptr(dy), dtype_code(dy.dtype), dy.stride(0), ptr(y), dtype_code(y.dtype) if y is not None else 0,
                             y.stride(0) if y is not None else 0, int(relu), alpha, alpha_cols, ptr(dz),
                             dtype_code(out_dtype), ld, ptr(db), M, N, *(drop_args(*dropout[:2]) if dropout and dropout[0] > 0 else (0.0, 0, 0)),
                             ptr(DROP_SEED_DEV) if dropout and dropout[0] > 0 else None, stream()))
    _lib.count_launches(1)
    return (dz[:, :N] if (dz is not None and ld != N) else dz), db


def mamba_conv_silu(xz: torch.Tensor, ED: int, w: torch.Tensor, bias: Optional[torch.Tensor], B: int, L: int) -> torch.Tensor:
    """silu(depthwise causal conv1d) of the x half of xz (B*L, >= ED) -> (B*L, ED)   (mamba.py:270-276)."""
    require_device(xz)
    assert xz.dtype == torch.float32 and xz.stride(1) == 1 and w.is_contiguous()
    y = torch.empty((B * L, ED), device=xz.device, dtype=torch.float32)
    check(load().v2m_mamba_conv_silu(ptr(xz), xz.stride(0), ptr(w), ptr(bias), ptr(y), ED, B, L, ED, w.shape[1], stream()))
    _lib.count_launches(1)
    return y


def mamba_step(xz: torch.Tensor, ED: int, conv_w: torch.Tensor, conv_b: Optional[torch.Tensor], x_proj_w: torch.Tensor,
               dt_w: torch.Tensor, dt_b: torch.Tensor, A_log: torch.Tensor, D: torch.Tensor, h: Optional[torch.Tensor],
               inputs: torch.Tensor):
    """Core of MambaBlock.step (mamba.py:407-470) between in_proj and out_proj: xz (B, 2 ED) = [x | z], cache (h (B, ED, N) or
    None, inputs (B, ED, d_conv - 1)) -> (gated output (B, ED), new h, new inputs).  The cache tensors are not modified."""
    require_device(xz)
    B = xz.shape[0]
    N, R = A_log.shape[1], dt_w.shape[1]
    KW = conv_w.shape[1]
    assert xz.dtype == torch.float32 and xz.stride(1) == 1 and inputs.shape == (B, ED, KW - 1)
    inputs = inputs.float().contiguous()
    f32 = dict(device=xz.device, dtype=torch.float32)
    xs, in_new = torch.empty((B, ED), **f32), torch.empty((B, ED, KW - 1), **f32)
    check(load().v2m_mamba_step_conv(ptr(xz), xz.stride(0), ptr(inputs), ptr(conv_w.contiguous()), ptr(conv_b), ptr(xs), ptr(in_new),
                                     B, ED, KW, stream()))
    _lib.count_launches(1)
    dbc = linear(xs, x_proj_w)                                                  # (B, R + 2N), mamba.py:445
    h_new, out = torch.empty((B, ED, N), **f32), torch.empty((B, ED), **f32)
    if h is not None:
        h = h.float().contiguous()
        assert h.shape == (B, ED, N)
    z = xz[:, ED:]
    check(load().v2m_mamba_step_ssm(ptr(xs), ptr(dbc), dbc.stride(0), ptr(dt_w.contiguous()), ptr(dt_b), ptr(A_log.contiguous()), ptr(D),
                                    ptr(z), xz.stride(0), ptr(h), ptr(h_new), ptr(out), B, ED, N, R, stream()))
    _lib.count_launches(1)
    return out, h_new, in_new


def selective_scan(x: torch.Tensor, delta_raw: torch.Tensor, dt_bias: Optional[torch.Tensor], A_log: torch.Tensor,
                   Bm: torch.Tensor, Cm: torch.Tensor, D: torch.Tensor, z: Optional[torch.Tensor], B: int, L: int,
                   plus: bool = False) -> torch.Tensor:
    """Fused selective scan (mamba.py:293-351 + the gating of :281-287): 2-D row-major views (B*L, .) with unit inner
    stride; Bm / Cm share their leading dimension (slices of the x_proj output)."""
    require_device(x)
    ED, N = A_log.shape
    for t in (x, delta_raw, Bm, Cm) + ((z,) if z is not None else ()):
        assert t.dtype == torch.float32 and t.stride(1) == 1
    assert Bm.stride(0) == Cm.stride(0)
    out = torch.empty((B * L, ED), device=x.device, dtype=torch.float32)
    lib = load()
    ws_bytes = int(lib.v2m_selective_scan_workspace(B, L, ED, N))
    ws = torch.empty((ws_bytes // 4,), device=x.device, dtype=torch.float32) if ws_bytes else None
    check(lib.v2m_selective_scan_fwd(ptr(x), x.stride(0), ptr(delta_raw), delta_raw.stride(0), ptr(dt_bias), ptr(A_log.contiguous()),
                                     ptr(Bm), ptr(Cm), Bm.stride(0), ptr(D), ptr(z), z.stride(0) if z is not None else 0,
                                     ptr(out), ED, B, L, ED, N, int(plus), ptr(ws), ws_bytes, stream()))
    _lib.count_launches(3 if ws_bytes else 1)
    return out


def mamba_conv_silu_bwd(xz: torch.Tensor, ED: int, w: torch.Tensor, bias: Optional[torch.Tensor], dy: torch.Tensor, dxz: torch.Tensor,
                        B: int, L: int):
    """Backward of mamba_conv_silu: writes the input gradient into the first ED columns of dxz (B*L, >= ED); returns
    (dw (ED, KW), dbias (ED) or None)."""
    require_device(xz)
    assert xz.stride(1) == 1 and dy.stride(1) == 1 and dxz.stride(1) == 1 and w.is_contiguous()
    dw = torch.zeros_like(w)
    dbias = torch.zeros((ED,), device=xz.device, dtype=torch.float32) if bias is not None else None
    check(load().v2m_mamba_conv_silu_bwd(ptr(xz), xz.stride(0), ptr(w), ptr(bias), ptr(dy), dy.stride(0), ptr(dxz), dxz.stride(0), ptr(dw),
                                         ptr(dbias), B, L, ED, w.shape[1], stream()))
    _lib.count_launches(1)
    return dw, dbias


def selective_scan_bwd(x: torch.Tensor, delta_raw: torch.Tensor, dt_bias: Optional[torch.Tensor], A_log: torch.Tensor, Bm: torch.Tensor,
                       Cm: torch.Tensor, D: torch.Tensor, z: Optional[torch.Tensor], dout: torch.Tensor, dBm: torch.Tensor,
                       dCm: torch.Tensor, dz: Optional[torch.Tensor], B: int, L: int, plus: bool = False):
    """Backward of selective_scan.  dBm / dCm: zeroed (B*L, N) views (slices of the x_proj output gradient) that receive the
    channel sums; dz: (B*L, ED) view that receives the gate gradient.  Returns (dx, ddelta_raw, dA_log, dD, ddt_bias)."""
    require_device(x)
    ED, N = A_log.shape
    for t in (x, delta_raw, Bm, Cm, dout, dBm, dCm) + ((z, dz) if z is not None else ()):
        assert t.dtype == torch.float32 and t.stride(1) == 1
    assert Bm.stride(0) == Cm.stride(0) and dBm.stride(0) == dCm.stride(0)
    f32 = dict(device=x.device, dtype=torch.float32)
    dx, ddraw = torch.empty((B * L, ED), **f32), torch.empty((B * L, ED), **f32)
    dA_log, dD = torch.zeros((ED, N), **f32), torch.zeros((ED,), **f32)
    ddtb = torch.zeros((ED,), **f32) if dt_bias is not None else None
    lib = load()
    hs_bytes = int(lib.v2m_selective_scan_bwd_workspace(B, L, ED, N))
    hs = torch.empty((hs_bytes // 4,), **f32)
    check(lib.v2m_selective_scan_bwd(ptr(x), x.stride(0), ptr(delta_raw), delta_raw.stride(0), ptr(dt_bias), ptr(A_log.contiguous()),
                                     ptr(Bm), ptr(Cm), Bm.stride(0), ptr(D), ptr(z), z.stride(0) if z is not None else 0,
                                     ptr(dout), dout.stride(0), ptr(hs), hs_bytes, ptr(dx), ED, ptr(ddraw), ED, ptr(dBm), ptr(dCm),
                                     dBm.stride(0), ptr(dz), dz.stride(0) if dz is not None else 0, ptr(dA_log), ptr(dD), ptr(ddtb),
                                     B, L, ED, N, int(plus), stream()))
    _lib.count_launches(1)
    return dx, ddraw, dA_log, dD, ddtb


def rmsnorm(x: torch.Tensor, w: Optional[torch.Tensor], eps: float = 1e-5) -> torch.Tensor:
    require_device(x)
    x = x.contiguous()
    y = torch.empty_like(x)
    check(load().v2m_rmsnorm(ptr(x), ptr(w), ptr(y), _rows(x), x.shape[-1], eps, stream()))
    _lib.count_launches(1)
    return y


def rmsnorm_bwd(x: torch.Tensor, w: Optional[torch.Tensor], dy: torch.Tensor, eps: float = 1e-5):
    """(dx, dw or None) of rmsnorm."""
    require_device(x)
    x, dy = x.contiguous(), dy.contiguous()
    D = x.shape[-1]
    dx = torch.empty_like(x)
    dw = torch.zeros((D,), device=x.device, dtype=torch.float32) if w is not None else None
    check(load().v2m_rmsnorm_bwd(ptr(x), ptr(w), ptr(dy), ptr(dx), ptr(dw), _rows(x), D, eps, stream()))
    _lib.count_launches(1)
    return dx, dw


def layernorm_bwd(x: torch.Tensor, gamma: torch.Tensor, dy: torch.Tensor, eps: float = 1e-5, dg_out: Optional[torch.Tensor] = None,
                  db_out: Optional[torch.Tensor] = None):
    """dg_out / db_out: caller-owned fp32 buffers the affine gradients are ADDED to (instead of fresh zeroed ones)."""
    require_device(x)
    x, dy = x.contiguous(), dy.contiguous()
    D = x.shape[-1]
    M = x.numel() // D
    dx = torch.empty_like(x)
    dg = torch.zeros((D,), device=x.device, dtype=torch.float32) if dg_out is None else dg_out
    db = torch.zeros((D,), device=x.device, dtype=torch.float32) if db_out is None else db_out
    check(load().v2m_layernorm_bwd(ptr(x), dtype_code(x.dtype), ptr(gamma), ptr(dy), dtype_code(dy.dtype), ptr(dx),
                                   dtype_code(dx.dtype), ptr(dg), ptr(db), M, D, eps, stream()))
    _lib.count_launches(1)
    return dx, dg, db


def embed_bwd(idx: torch.Tensor, d: torch.Tensor, n_rows_table: int, D: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out: caller-owned fp32 (rows, D) buffer the table gradient is ADDED to (instead of a fresh zeroed one)."""
    require_device(d)
    idx = idx.contiguous().view(-1)
    dt = torch.zeros((n_rows_table, D), device=d.device, dtype=torch.float32) if out is None else out
    assert dt.shape == (n_rows_table, D) and dt.is_contiguous()
    check(load().v2m_embed_bwd(ptr(idx), ptr(d), dtype_code(d.dtype), d.stride(0), ptr(dt), idx.numel(), D, stream()))
    _lib.count_launches(1)
    return dt


def amt_loss(logits: torch.Tensor, tgt: torch.Tensor, tgt_emotion: torch.Tensor, ignore: int = 158, smooth: float = 0.1,
             w_ce: float = 0.4, w_bce: float = 0.6, want_grad: bool = True, norm: Optional[torch.Tensor] = None):
    """Returns (scratch [ce_sum, bce_sum, n_valid] on device, dlogits or None).  norm (device fp32 [n_valid, rows])
    replaces the local normalisers (data-parallel ranks pass global / world, see count_valid)."""
    require_device(logits)
    if tgt.dtype != torch.int64:
        raise TypeError("amt_loss: targets must be int64 (got %s)" % tgt.dtype)
    logits = logits.contiguous()
    Cn = logits.shape[-1]
    R = logits.numel() // Cn
    tgt = tgt.contiguous().view(-1)
    tgt_emotion = tgt_emotion.contiguous().float()
    scratch = torch.empty((3,), device=logits.device, dtype=torch.float32)
    dl = torch.empty_like(logits) if want_grad else None
    check(load().v2m_amt_loss(ptr(logits), ptr(tgt), ptr(tgt_emotion), R, Cn, ignore, smooth, w_ce, w_bce, ptr(scratch), ptr(dl),
                              ptr(norm), stream()))
    _lib.count_launches(2)
    return scratch, dl


def count_valid(tgt: torch.Tensor, ignore: int = 158) -> torch.Tensor:
    """Device fp32 scalar: number of targets != ignore (the normaliser of nn.CrossEntropyLoss(ignore_index), train.py:222)."""
    require_device(tgt)
    if tgt.dtype != torch.int64:
        raise TypeError("count_valid: targets must be int64 (got %s)" % tgt.dtype)
    tgt = tgt.contiguous().view(-1)
    out = torch.empty((1,), device=tgt.device, dtype=torch.float32)
    check(load().v2m_count_valid(ptr(tgt), tgt.numel(), ignore, ptr(out), stream()))
    _lib.count_launches(1)
    return out


def amt_metrics(logits: torch.Tensor, tgt: torch.Tensor, pad: int = 158, ks=(1, 3, 5)) -> torch.Tensor:
    """int32 device counters [valid, argmax hits, hits@ks[0], hits@ks[1], hits@ks[2]] over all (video, position) rows
    (compute_vevo_accuracy / compute_hits_k, dataset/vevo_dataset.py:653-701); accuracy = c[1] / c[0], hits@k = c[2+i] / c[0]."""
    require_device(logits)
    logits = logits.float().contiguous()
    Cn = logits.shape[-1]
    R = logits.numel() // Cn
    tgt = tgt.contiguous().view(-1)
    assert tgt.numel() == R and tgt.dtype == torch.int64 and len(ks) == 3
    counters = torch.empty((5,), device=logits.device, dtype=torch.int32)
    check(load().v2m_amt_metrics(ptr(logits), ptr(tgt), R, Cn, pad, int(ks[0]), int(ks[1]), int(ks[2]), ptr(counters), stream()))
    _lib.count_launches(1)
    return counters


def amt_correspondence(logits: torch.Tensor, tgt_emotion: torch.Tensor, tgt_emotion_prob: torch.Tensor, threshold: float = 0.8,
                       chord_end: int = 157) -> torch.Tensor:
    """int32 device counters [pt, right] of compute_vevo_correspondence (dataset/vevo_dataset.py:747-810) over all rows:
    logits (..., Cn), tgt_emotion (..., Ce) with the 14 chord-quality columns first and the padding flag last, tgt_emotion_prob (...).
    correspondence = right / pt (the reference returns -1 when pt == 0); `compute_vevo_correspondence` below wraps it."""
    require_device(logits)
    logits = logits.float().contiguous()
    Cn = logits.shape[-1]
    R = logits.numel() // Cn
    emo = tgt_emotion.float().contiguous()
    Ce = emo.shape[-1]
    prob = tgt_emotion_prob.float().contiguous().view(-1)
    assert emo.numel() == R * Ce and prob.numel() == R, (emo.shape, prob.shape, R)
    counters = torch.empty((2,), device=logits.device, dtype=torch.int32)
    check(load().v2m_amt_correspondence(ptr(logits), ptr(emo), ptr(prob), R, Cn, Ce, float(threshold), chord_end, ptr(counters), stream()))
    _lib.count_launches(1)
    return counters


def compute_vevo_correspondence(out: torch.Tensor, tgt: torch.Tensor, tgt_emotion: torch.Tensor, tgt_emotion_prob: torch.Tensor,
                                emotion_threshold: float) -> float:
    """Same signature and return convention as the reference's function (dataset/vevo_dataset.py:747): 1.0 for an empty batch,
    -1 when no position qualifies, else right / pt.  `tgt` is unused there as well."""
    if tgt_emotion.numel() == 0:
        return 1.0
    pt, right = amt_correspondence(out, tgt_emotion, tgt_emotion_prob, emotion_threshold).tolist()
    return -1 if pt == 0 else right / pt


def adam_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, lr: float, b1: float, b2: float, eps: float,
              step: int, grad_scale: float = 1.0, *, dyn: Optional[torch.Tensor] = None, p16: Optional[torch.Tensor] = None,
              zero_grad: bool = False, counter: Optional[torch.Tensor] = None, weight_decay: float = 0.0) -> None:
    """torch.optim.Adam semantics on flat buffers (weight_decay > 0: torch.optim.AdamW's decoupled decay).  dyn (device fp32 [lr, 1-b1^t, 1-b2^t]) replaces the scalar lr / step at run
    time; p16: bf16 mirror of the updated parameters; zero_grad clears g; counter (device, 4 bytes) is incremented."""
    require_device(p)
    assert p.is_contiguous() and g.is_contiguous() and p.dtype == torch.float32
    check(load().v2m_adam_step(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), lr, b1, b2, eps, weight_decay, step, grad_scale, ptr(dyn), ptr(p16),
                               int(zero_grad), ptr(counter), stream()))
    _lib.count_launches(1)


def attention_bwd(q, k, v, o, dO, lse, Er, dq, dk, dv, dEr, *, B, Hq, Hkv, Lq, Lk, dh, q_strides, k_strides, v_strides,
                  o_strides, do_strides, dq_strides, dkv_strides, causal, q_scale=1.0, tensor_core=False, dropout=None, dq_scale=1.0):
    """tensor_core=True (bf16, head_dim 64): dk / dv are bf16 outputs written once by the mma.sync kernels of
    csrc/attn_bwd_tc.cu; otherwise dk / dv are zeroed fp32 buffers the exact SIMT kernel accumulates into."""
    require_device(q)
    a = _lib.AttnBwd()
    a.q, a.k, a.v, a.o, a.dO, a.lse, a.Er = ptr(q), ptr(k), ptr(v), ptr(o), ptr(dO), ptr(lse), ptr(Er)
    a.dq, a.dk, a.dv, a.dEr = ptr(dq), ptr(dk), ptr(dv), ptr(dEr)
    (a.q_sb, a.q_sl), (a.k_sb, a.k_sl), (a.v_sb, a.v_sl) = q_strides, k_strides, v_strides
    (a.o_sb, a.o_sl), (a.do_sb, a.do_sl), (a.dq_sb, a.dq_sl), (a.dkv_sb, a.dkv_sl) = o_strides, do_strides, dq_strides, dkv_strides
    a.B, a.Hq, a.Hkv, a.Lq, a.Lk, a.dh, a.causal = B, Hq, Hkv, Lq, Lk, dh, int(causal)
    a.er_len = Er.shape[0] if Er is not None else 0
    a.dtype = dtype_code(q.dtype)
    a.q_scale = q_scale
    a.dq_scale = dq_scale                                        # dq stored times dq_scale (backward of a query scaling applied upstream)
    if dropout is not None and dropout[0] > 0.0:                 # the forward's (p, seed)
        a.drop_scale, a.drop_thresh, a.drop_seed = drop_args(dropout[0], dropout[1])
        a.drop_seed_dev = ptr(DROP_SEED_DEV)
    if tensor_core:
        n = int(load().v2m_attn_bwd_tc_workspace(B, Hq, Lq, Lk, int(Er is not None)))
        ws = torch.empty(n, dtype=torch.uint8, device=q.device)
        check(load().v2m_attn_bwd_tc(C.byref(a), ptr(ws), n, stream()))
        _lib.count_launches(3 if Er is not None else 2)
        return
    check(load().v2m_attn_bwd(C.byref(a), stream()))
    _lib.count_launches(1)


def linear_general(a: torch.Tensor, b: torch.Tensor, *, a_mn: bool, b_mn: bool, M: int, N: int, K: int,
                   out_dtype: torch.dtype = torch.bfloat16, out: Optional[torch.Tensor] = None, accumulate: bool = False,
                   gate: Optional[torch.Tensor] = None, gate_scale: float = 1.0, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bf16 tcgen05 GEMM C[M,N] = A B^T-style with optionally transposed storage: a is [M,K] (or [K,M] when a_mn),
    b is [N,K] (or [K,N] when b_mn), both row-major 2-D bf16 tensors (leading dims from the strides).
    accumulate (fp32 `out` given by the caller): out += A B^T through coalesced vector reductions."""
    require_device(a)
    assert a.dtype == b.dtype == torch.bfloat16 and a.stride(1) == 1 and b.stride(1) == 1
    if out is None:
        assert not accumulate
        out = torch.empty((M, N), device=a.device, dtype=out_dtype)
    else:
        assert out.shape == (M, N) and out.stride(1) == 1 and out.dtype == out_dtype
    ep = Epilogue()
    ep.accumulate = int(accumulate)
    if residual is not None:                     # C = acc + residual (bf16): the residual-branch gradient joins the dX GEMM
        assert gate is None and residual.dtype == torch.bfloat16 and residual.shape == (M, N) and residual.stride(1) == 1 and not accumulate
        ep.residual, ep.ldr, ep.residual_bf16 = ptr(residual), residual.stride(0), 1
    if gate is not None:                         # C = gate > 0 ? acc * gate_scale : 0 (ReLU / dropout backward fused into the dX GEMM)
        assert gate.dtype == torch.bfloat16 and gate.shape == (M, N) and gate.stride(1) == 1 and not accumulate
        ep.residual, ep.ldr, ep.residual_bf16, ep.residual_gate, ep.gate_scale = ptr(gate), gate.stride(0), 1, 1, float(gate_scale)
    check(load().v2m_gemm_bf16_general(ptr(a), a.stride(0), int(a_mn), ptr(b), b.stride(0), int(b_mn), ptr(out), out.stride(0),
                                       dtype_code(out_dtype), M, N, K, C.byref(ep), stream()))
    _lib.count_launches(1)
    return out
