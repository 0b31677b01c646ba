"""Training path: the AMT forward as a chain of torch.autograd.Functions whose forward AND backward are our
kernels (torch.autograd only does the graph bookkeeping; the reference gets its gradients from autograd over ATen ops,
utilities/run_model_vevo.py:84-121).

Functions:  LinearFn (GEMM + fused bias / q-scale / ReLU / residual, backward = dy_prep + 2 GEMMs),
AttnSelfFn / AttnCrossFn (fused attention forward with lse, backward = attn_bwd: dQ, dK, dV, dEr), LayerNormFn,
EmbedKeyFn (embedding sums + key column), AmtLossFn (0.4 CE + 0.6 BCE with the gradient produced in the same pass).
"""
from typing import Optional

import torch

from . import engine, ops

F32, BF16 = torch.float32, torch.bfloat16


def _pad8(n):
    return (n + 7) // 8 * 8


# ----------------------------------------------------------------------------------------------- GEMM helpers
def _gemm_dx(dz: torch.Tensor, w: torch.Tensor, K: int) -> torch.Tensor:
    """dX[M,K] = dz[M,N] @ W[N,K]   (W read as the transposed operand in place)."""
    M, N = dz.shape
    if dz.dtype == F32:
        return ops.gemm_strided(dz, dz.stride(0), 1, w, 1, w.stride(0), M, K, N)
    return ops.linear_general(dz, w, a_mn=False, b_mn=True, M=M, N=K, K=N, out_dtype=BF16)


def _gemm_dw(dz: torch.Tensor, x: torch.Tensor, K: int) -> torch.Tensor:
    """dW[N,K] = dz[M,N]^T @ x[M,K]   (both operands read transposed in place), fp32 result."""
    M, N = dz.shape
    if dz.dtype == F32:
        return ops.dw_f32(dz, x, K)
    return ops.linear_general(dz, x, a_mn=True, b_mn=True, M=N, N=K, K=M, out_dtype=F32)


# ---- direct gradient accumulation -------------------------------------------------------------------------------------
# A Trainer re-homes every parameter's gradient into one flat fp32 buffer that its optimiser kernel clears after every step
# (trainer.FlatParams) and marks the parameter with `_v2m_direct_grad`.  For such parameters the backward kernels ADD their
# result straight into that buffer (split-K reductions of the dW GEMM, the column-sum / LayerNorm / embedding / Er atomics
# all accumulate anyway) and the Function returns None for the parameter: no temporary, no zero-fill, no AccumulateGrad add --
# two to three tiny launches less per parameter and step (~400 of 830 at the AMT).  `_v2m_grad_ready` (set by
# trainer.GradBuckets) is told that the gradient is complete, as the post-accumulate hook would have been.
def direct_grad(p) -> Optional[torch.Tensor]:
    if p is None or not getattr(p, "_v2m_direct_grad", False) or p.grad is None:
        return None
    return p.grad


def grad_ready(p) -> None:
    cb = getattr(p, "_v2m_grad_ready", None)
    if cb is not None:
        cb()


class LinearFn(torch.autograd.Function):
    """y = drop?(relu?((x @ w[:, :K].T + b) * alpha_n)) + residual.   x (M, >=K) compute dtype; w, b fp32 masters.
    dropout = None or (p, seed, after_residual): mask fused into the GEMM epilogue, recomputed by dy_prep in backward."""

    @staticmethod
    def forward(ctx, x, w, b, wc, K, relu, alpha, alpha_cols, residual, res_mod, out_dtype, dropout=None, owner=None,
                in_gate_scale=None, pre_gated=False, pre_scaled=False, in_link=None, res_link=None):
        """owner = (weight parameter, bias parameter or None, row slice or None): the parameters `w` / `b` are (views of), for
        direct gradient accumulation (see direct_grad).
        in_gate_scale (bf16 path): x is the output of a relu (+ fused dropout of that scale) LinearFn built with pre_gated=True;
        this layer's dX GEMM applies that activation's backward in its epilogue (dx = x > 0 ? dx * scale : 0), so the layer
        below needs no pass over dy / y / dz (its backward then only sums the bias gradient)."""
        if dropout is not None and dropout[0] <= 0.0:
            dropout = None
        ctx.owner = owner
        ctx.in_gate_scale = in_gate_scale if x.dtype == BF16 else None
        ctx.pre_gated = bool(pre_gated) and x.dtype == BF16
        assert not ctx.pre_gated or (relu and alpha_cols == 0)
        # pre_scaled: the consumer (attention backward, dq_scale) already multiplied the gradient of the alpha-scaled columns by alpha
        ctx.pre_scaled = bool(pre_scaled)
        assert not ctx.pre_scaled or (not relu and dropout is None)
        # Residual link of a sub-layer (bf16 path): its input x feeds the first linear (in_link) and the residual add of the last one
        # (res_link, the same dict).  The last linear's backward -- which always runs first -- leaves the residual-branch gradient
        # in the dict instead of returning it, and the first linear's dX GEMM adds it in its epilogue: autograd no longer sums
        # the two branches with a separate pass over three activation-sized tensors per sub-layer.
        ctx.in_link = in_link if x.dtype == BF16 else None
        ctx.res_link = res_link if (x.dtype == BF16 and residual is not None and res_mod == 0 and residual.dtype == BF16) else None
        # drop(acc + residual) is only used for the constant positional-encoding rows (no gradient flows to the residual)
        assert not (dropout is not None and dropout[2] and residual is not None and residual.requires_grad)
        y = ops.linear(x, wc, b, k=K, relu=relu, alpha=alpha, alpha_cols=alpha_cols, residual=residual, res_mod=res_mod,
                       out_dtype=out_dtype, dropout=dropout)
        ctx.save_for_backward(x, wc, y if relu else None)
        ctx.meta = (K, relu, alpha, alpha_cols, residual is not None and res_mod == 0, b is not None, tuple(w.shape), x.dtype)
        ctx.dropout = dropout
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wc, y = ctx.saved_tensors
        K, relu, alpha, alpha_cols, res_grad, has_b, wshape, cdt = ctx.meta
        dy = dy.contiguous()
        plain = ((not relu) and (alpha_cols == 0 or ctx.pre_scaled) and dy.dtype == cdt and ctx.dropout is None) or (ctx.pre_gated and dy.dtype == cdt)
        wp, bp, rows = ctx.owner if ctx.owner is not None else (None, None, None)
        gw, gb = direct_grad(wp), direct_grad(bp) if has_b else None
        if gw is not None and rows is not None:
            gw = gw[rows]
        if gb is not None and rows is not None:
            gb = gb[rows]
        direct_w = gw is not None and cdt == BF16 and tuple(gw.shape) == (wshape[0], K) and gw.stride(1) == 1 and (gw.stride(0) * 4) % 16 == 0 \
            and gw.data_ptr() % 16 == 0
        direct_b = gb is not None and gb.is_contiguous()
        if (ctx.pre_gated or ctx.pre_scaled) and plain:                      # relu' / dropout mask / column scaling were applied by the consumer
            dz, db = ops.dy_prep(dy, None, False, 1.0, 0, cdt, want_dz=False, dropout=None, db_out=gb if direct_b else None)
        else:
            dz, db = ops.dy_prep(dy, y, relu, alpha, alpha_cols, cdt, want_dz=not plain, dropout=ctx.dropout, db_out=gb if direct_b else None)
        if plain:
            dz = dy
        extra = ctx.in_link.pop("dres", None) if ctx.in_link is not None else None    # residual-branch gradient of this sub-layer
        if not ctx.needs_input_grad[0]:
            assert extra is None, "a linked residual gradient needs the gradient of the sub-layer input"
            dx = None
        elif extra is not None and cdt == BF16 and extra.shape == (dz.shape[0], K) and x.shape[1] == K:
            M_, N_ = dz.shape
            dx = ops.linear_general(dz, wc, a_mn=False, b_mn=True, M=M_, N=K, K=N_, out_dtype=BF16, residual=extra)
            extra = None
        elif ctx.in_gate_scale is not None and cdt == BF16:
            M_, N_ = dz.shape
            dx = ops.linear_general(dz, wc, a_mn=False, b_mn=True, M=M_, N=K, K=N_, out_dtype=BF16, gate=x[:, :K], gate_scale=ctx.in_gate_scale)
        else:
            dx = _gemm_dx(dz, wc, K)
        if extra is not None:                                                # not fusable: plain sum
            dx = dx + extra.to(dx.dtype)
        if dx is not None and dx.shape[1] != x.shape[1]:                     # x carries zero padding columns
            full = torch.zeros_like(x)
            full[:, :K] = dx
            dx = full
        if direct_w:                                                         # dW added straight into the flat gradient buffer
            M_, N_ = dz.shape
            ops.linear_general(dz, x, a_mn=True, b_mn=True, M=N_, N=K, K=M_, out_dtype=F32, out=gw, accumulate=True)
            dw = None
        else:
            dw = _gemm_dw(dz, x, K)
            if dw.shape[1] != wshape[1]:
                dw = dw[:, :wshape[1]].contiguous()
        if direct_b:
            db = None
        if direct_w:
            grad_ready(wp)
        if direct_b and bp is not wp:
            grad_ready(bp)
        dres = dy if res_grad else None
        if dres is not None and ctx.res_link is not None and ctx.res_link.get("armed"):
            ctx.res_link["dres"] = dres                                      # picked up by the sub-layer's first linear (see forward)
            dres = None
        return dx, dw, (db if has_b else None), None, None, None, None, None, dres, None, None, None, None, None, None, None, None, None


class AttnSelfFn(torch.autograd.Function):
    """ctx[M,E] = attention over the fused qkv [M, 3E] (q pre-scaled by the projection epilogue), optional Er, causal."""

    @staticmethod
    def forward(ctx, qkv, er, erc, B, L, H, causal, dropout=None, er_owner=None, dq_scale=1.0):
        """dq_scale: the q gradient is stored times dq_scale (the in-projection that produced qkv was built with pre_scaled=True: the
        backward of its 1 / sqrt(d) column scaling is applied here, for free, instead of by a pass over the whole gradient)."""
        if dropout is not None and dropout[0] <= 0.0:
            dropout = None
        ctx.er_owner = er_owner           # the Er parameter, for direct gradient accumulation
        ctx.dq_scale = dq_scale
        E = qkv.shape[1] // 3
        dh = E // H
        out = torch.empty((B * L, E), device=qkv.device, dtype=qkv.dtype)
        lse = torch.empty((B * H, L), device=qkv.device, dtype=F32)
        ld = qkv.stride(0)
        ops.attention(qkv, qkv[:, E:], qkv[:, 2 * E:], out, B=B, Hq=H, Hkv=H, Lq=L, Lk=L, dh=dh,
                      q_strides=(L * ld, ld), k_strides=(L * ld, ld), v_strides=(L * ld, ld), o_strides=(L * E, E),
                      causal=causal, Er=erc, lse=lse, dropout=dropout)
        ctx.save_for_backward(qkv, out, lse, erc)
        ctx.meta = (B, L, H, causal, er is not None)
        ctx.dropout = dropout
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse, erc = ctx.saved_tensors
        B, L, H, causal, has_er = ctx.meta
        E = qkv.shape[1] // 3
        dh = E // H
        dout = dout.contiguous()
        ld = qkv.stride(0)
        dqkv = torch.empty_like(qkv)
        ger = direct_grad(ctx.er_owner) if has_er else None
        direct_er = ger is not None and ger.is_contiguous() and tuple(ger.shape) == tuple(erc.shape)
        der = (ger if direct_er else torch.zeros(erc.shape, device=qkv.device, dtype=F32)) if has_er else None
        common = dict(B=B, Hq=H, Hkv=H, Lq=L, Lk=L, dh=dh, q_strides=(L * ld, ld), k_strides=(L * ld, ld), v_strides=(L * ld, ld),
                      o_strides=(L * E, E), do_strides=(L * E, E), dq_strides=(L * ld, ld), causal=causal, dropout=ctx.dropout,
                      dq_scale=ctx.dq_scale)
        if qkv.dtype == BF16 and dh == 64:
            # tensor-core kernels write dK / dV straight into the fused gradient (bf16), no fp32 staging
            ops.attention_bwd(qkv, qkv[:, E:], qkv[:, 2 * E:], out, dout, lse, erc if has_er else None, dqkv, dqkv[:, E:],
                              dqkv[:, 2 * E:], der, dkv_strides=(L * ld, ld), tensor_core=True, **common)
        else:
            dkv32 = torch.zeros((B * L, 2 * E), device=qkv.device, dtype=F32)
            ops.attention_bwd(qkv, qkv[:, E:], qkv[:, 2 * E:], out, dout, lse, erc if has_er else None, dqkv, dkv32, dkv32[:, E:],
                              der, dkv_strides=(L * 2 * E, 2 * E), **common)
            dqkv[:, E:] = dkv32                                           # fp32 accumulators -> gradient dtype
        if has_er and direct_er:
            grad_ready(ctx.er_owner)
            der = None
        return dqkv, der, None, None, None, None, None, None, None, None


class AttnCrossFn(torch.autograd.Function):
    """ctx[Mq,E] = attention of q [B*T, E] (pre-scaled) over kv [B*S, 2E] (non-causal, no Er)."""

    @staticmethod
    def forward(ctx, q, kv, B, T, S, H, dropout=None, dq_scale=1.0):
        if dropout is not None and dropout[0] <= 0.0:
            dropout = None
        ctx.dq_scale = dq_scale
        E = q.shape[1]
        dh = E // H
        out = torch.empty((B * T, E), device=q.device, dtype=q.dtype)
        lse = torch.empty((B * H, T), device=q.device, dtype=F32)
        ops.attention(q, kv, kv[:, E:], out, B=B, Hq=H, Hkv=H, Lq=T, Lk=S, dh=dh, q_strides=(T * E, E),
                      k_strides=(S * 2 * E, 2 * E), v_strides=(S * 2 * E, 2 * E), o_strides=(T * E, E), causal=False, lse=lse,
                      dropout=dropout)
        ctx.save_for_backward(q, kv, out, lse)
        ctx.meta = (B, T, S, H)
        ctx.dropout = dropout
        return out

    @staticmethod
    def backward(ctx, dout):
        q, kv, out, lse = ctx.saved_tensors
        B, T, S, H = ctx.meta
        E = q.shape[1]
        dh = E // H
        dout = dout.contiguous()
        dq = torch.empty_like(q)
        common = dict(B=B, Hq=H, Hkv=H, Lq=T, Lk=S, dh=dh, q_strides=(T * E, E), k_strides=(S * 2 * E, 2 * E),
                      v_strides=(S * 2 * E, 2 * E), o_strides=(T * E, E), do_strides=(T * E, E), dq_strides=(T * E, E),
                      dkv_strides=(S * 2 * E, 2 * E), causal=False, dropout=ctx.dropout, dq_scale=ctx.dq_scale)
        if q.dtype == BF16 and dh == 64:
            dkv = torch.empty_like(kv)
            ops.attention_bwd(q, kv, kv[:, E:], out, dout, lse, None, dq, dkv, dkv[:, E:], None, tensor_core=True, **common)
            return dq, dkv, None, None, None, None, None, None
        dkv32 = torch.zeros((B * S, 2 * E), device=q.device, dtype=F32)
        ops.attention_bwd(q, kv, kv[:, E:], out, dout, lse, None, dq, dkv32, dkv32[:, E:], None, **common)
        return dq, dkv32.to(kv.dtype) if kv.dtype != F32 else dkv32, None, None, None, None, None, None


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps, owner=None):
        y = ops.layernorm(x, gamma, beta, eps=eps)
        ctx.save_for_backward(x, gamma)
        ctx.eps = eps
        ctx.owner = owner                 # (gamma parameter, beta parameter) for direct gradient accumulation
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma = ctx.saved_tensors
        gp, bp = ctx.owner if ctx.owner is not None else (None, None)
        gg, gb = direct_grad(gp), direct_grad(bp)
        if gg is not None and gb is not None and gg.is_contiguous() and gb.is_contiguous():
            dx, _, _ = ops.layernorm_bwd(x, gamma, dy, ctx.eps, dg_out=gg, db_out=gb)
            grad_ready(gp)
            grad_ready(bp)
            return dx, None, None, None, None
        dx, dg, db = ops.layernorm_bwd(x, gamma, dy, ctx.eps)
        return dx, dg, db, None, None


class EmbedKeyFn(torch.autograd.Function):
    """[emb_a[idx_a] (+ emb_b[idx_b]) | key | 0-pad] as a (rows, pad8(D+1)) matrix in the compute dtype
    (video_music_transformer.py:984-999)."""

    @staticmethod
    def forward(ctx, idx_a, table_a, idx_b, table_b, key_rows, dtype, owner=None):
        ctx.owner = owner                 # (table_a parameter, table_b parameter) for direct gradient accumulation
        rows, D = idx_a.numel(), table_a.shape[1]
        ld = _pad8(D + 1)
        out = torch.zeros((rows, ld), device=table_a.device, dtype=dtype)
        e = ops.embed_sum(idx_a, table_a, idx_b, table_b, dtype)
        out[:, :D] = e
        out[:, D] = key_rows.to(dtype)
        ctx.save_for_backward(idx_a, idx_b)
        ctx.meta = (table_a.shape[0], table_b.shape[0] if table_b is not None else 0, D,
                    table_a.requires_grad, table_b is not None and table_b.requires_grad)
        return out

    @staticmethod
    def backward(ctx, dout):
        idx_a, idx_b = ctx.saved_tensors
        na, nb, D, ga, gb = ctx.meta
        dout = dout.contiguous()
        pa, pb = ctx.owner if ctx.owner is not None else (None, None)
        da = db = None
        if ga:
            g = direct_grad(pa)
            if g is not None and g.is_contiguous():
                ops.embed_bwd(idx_a, dout, na, D, out=g)
                grad_ready(pa)
            else:
                da = ops.embed_bwd(idx_a, dout, na, D)
        if gb:
            g = direct_grad(pb)
            if g is not None and g.is_contiguous():
                ops.embed_bwd(idx_b, dout, nb, D, out=g)
                grad_ready(pb)
            else:
                db = ops.embed_bwd(idx_b, dout, nb, D)
        return None, da, None, db, None, None, None


class AmtLossFn(torch.autograd.Function):
    """total = 0.4 * CE(label_smoothing 0.1, ignore 158) + 0.6 * BCEWithLogits  (run_model_vevo.py:101-119)."""

    @staticmethod
    def forward(ctx, logits, tgt, tgt_emotion, smooth, w_ce, w_bce, norm=None):
        """norm (device fp32 [n_valid, rows]): normalisers replacing this batch's own (a data-parallel rank passes
        global / world: the mean over ranks of the returned loss and of its gradient are then the global-batch ones)."""
        scratch, dl = ops.amt_loss(logits, tgt, tgt_emotion, 158, smooth, w_ce, w_bce, norm=norm)
        R = logits.numel() // logits.shape[-1]
        ce = scratch[0] / (norm[0] if norm is not None else scratch[2]).clamp_min(1.0)
        bce = scratch[1] / ((norm[1] if norm is not None else float(R)) * logits.shape[-1])
        ctx.save_for_backward(dl)
        ctx.parts = (ce.detach(), bce.detach())
        return w_ce * ce + w_bce * bce

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return dl * g, None, None, None, None, None, None


# ----------------------------------------------------------------------------------------------- model forward
def _lin(W, x, wname, bname, *, K=None, relu=False, alpha=1.0, alpha_cols=0, residual=None, res_mod=0, rows=None,
         out_dtype=None, dropout=None, in_gate_scale=None, pre_gated=False, pre_scaled=False, in_link=None, res_link=None):
    w = W._sd[wname]
    b = W._sd[bname] if bname is not None else None
    if rows is not None:
        w_v, b_v = w[rows], (b[rows] if b is not None else None)
    else:
        w_v, b_v = w, b
    wc = W.w(wname, rows=rows)
    K = K if K is not None else w.shape[1]
    return LinearFn.apply(x, w_v, b_v, wc, K, relu, alpha, alpha_cols, residual, res_mod, out_dtype or x.dtype, dropout, (w, b, rows),
                          in_gate_scale, pre_gated, pre_scaled, in_link, res_link)


def _ln(W, name, x):
    g, b = W._sd[name + ".weight"], W._sd[name + ".bias"]
    return LayerNormFn.apply(x, g, b, 1e-5, (g, b))


def amt_forward_autograd(model, x, x_root, x_attr, sem, key, scene, motion, emotion, mask: bool = True, dropout_p: float = 0.0,
                         seed: int = 0) -> torch.Tensor:
    """Differentiable VideoMusicTransformer.forward (video_music_transformer.py:978-1044).
    dropout_p > 0 (training, bf16): every nn.Dropout of the reference graph -- positional encodings
    (positional_encoding.py:23), attention probabilities (rpr.py:407 and the stock MultiheadAttention), sub-layer outputs
    before the residual adds and the FFN hidden layer (rpr.py:58-69, nn.TransformerEncoderLayer) -- is fused into the kernel
    that produces the tensor; each site gets its own seed derived from `seed`.  The masks come from a stateless integer
    hash, not from torch's Philox stream, so individual draws differ from the reference's for the same torch seed."""
    site = [0]

    def dr(after_res=False):                       # (p, seed, after_residual) of the next dropout site, None when off
        if dropout_p <= 0.0:
            return None
        site[0] += 1
        return (dropout_p, (seed * 0x9E3779B1 + site[0] * 0x85EBCA6B) & 0xFFFFFFFF, after_res)

    def dra():                                     # attention-probability site: (p, seed)
        d = dr()
        return None if d is None else d[:2]
    W = model._w()
    W.refresh()
    cfg = model._cfg()
    dt = W.dtype
    B, T = x.shape
    S = sem.shape[1]
    E, H, NL = cfg["d_model"], cfg["nhead"], cfg["n_layers"]
    dh = E // H
    scal = float(dh) ** -0.5
    sd = W._sd
    fuse_gate = dt == BF16                           # ReLU / dropout backward of the FFN fused into the next layer's dX GEMM
    fuse_q = True                                    # backward of the 1 / sqrt(d) of the query columns applied by the attention backward
    link = (lambda: {"armed": True}) if dt == BF16 else (lambda: None)   # residual links of the sub-layers (see LinearFn.forward)
    # ---- video stream
    vf_dim = sd["Linear_vis.weight"].shape[1]
    vin = ops.concat_features(sem, scene, motion, emotion, dt, vf_dim if dt == F32 else _pad8(vf_dim))
    pe_v = sd["positional_encoding_video.pe"].view(-1, E)
    xv = _lin(W, vin, "Linear_vis.weight", "Linear_vis.bias", K=vf_dim, residual=pe_v, res_mod=S, dropout=dr(True))
    for l in range(NL):
        p = "transformer.encoder.layers.%d." % l
        lk = link()
        qkv = _lin(W, xv, p + "self_attn.in_proj_weight", p + "self_attn.in_proj_bias", alpha=scal, alpha_cols=E, pre_scaled=fuse_q, in_link=lk)
        a = AttnSelfFn.apply(qkv, None, None, B, S, H, False, dra(), None, scal if fuse_q else 1.0)
        r = _lin(W, a, p + "self_attn.out_proj.weight", p + "self_attn.out_proj.bias", residual=xv, dropout=dr(), res_link=lk)
        xv = _ln(W, p + "norm1", r)
        d1 = dr()                                    # relu' and this dropout's mask are applied by linear2's dX GEMM (bf16 path)
        lk = link()
        hdn = _lin(W, xv, p + "linear1.weight", p + "linear1.bias", relu=True, dropout=d1, pre_gated=fuse_gate, in_link=lk)
        r = _lin(W, hdn, p + "linear2.weight", p + "linear2.bias", residual=xv, dropout=dr(), res_link=lk,
                 in_gate_scale=(ops.drop_args(*d1[:2])[0] if d1 is not None else 1.0) if fuse_gate else None)
        xv = _ln(W, p + "norm2", r)
    mem = _ln(W, "transformer.encoder.norm", xv)
    # ---- chord stream
    key_rows = key.reshape(B, 1).float().expand(B, T).reshape(-1)
    if cfg["chord_embed"]:
        xin = EmbedKeyFn.apply(x.reshape(-1), sd["chord_embedding_model.weight"], None, None, key_rows, dt,
                               (sd["chord_embedding_model.weight"], None))
    else:
        xin = EmbedKeyFn.apply(x_root.reshape(-1), sd["embedding_root.weight"], x_attr.reshape(-1), sd["embedding_attr.weight"],
                               key_rows, dt, (sd["embedding_root.weight"], sd["embedding_attr.weight"]))
    pe_c = sd["positional_encoding.pe"].view(-1, E)
    xf = _lin(W, xin, "Linear_chord.weight", "Linear_chord.bias", K=E + 1, residual=pe_c, res_mod=T, dropout=dr(True))
    for l in range(NL):
        p = "transformer.decoder.layers.%d." % l
        er = sd.get(p + "self_attn.Er")
        erc = W.table(p + "self_attn.Er") if er is not None else None
        lk = link()
        qkv = _lin(W, xf, p + "self_attn.in_proj_weight", p + "self_attn.in_proj_bias", alpha=scal, alpha_cols=E, pre_scaled=fuse_q, in_link=lk)
        a = AttnSelfFn.apply(qkv, er, erc, B, T, H, bool(mask), dra(), er, scal if fuse_q else 1.0)
        r = _lin(W, a, p + "self_attn.out_proj.weight", p + "self_attn.out_proj.bias", residual=xf, dropout=dr(), res_link=lk)
        xf = _ln(W, p + "norm1", r)
        lk = link()
        q = _lin(W, xf, p + "multihead_attn.in_proj_weight", p + "multihead_attn.in_proj_bias", rows=slice(0, E), alpha=scal,
                 alpha_cols=E, pre_scaled=fuse_q, in_link=lk)
        kv = _lin(W, mem, p + "multihead_attn.in_proj_weight", p + "multihead_attn.in_proj_bias", rows=slice(E, 3 * E))
        a = AttnCrossFn.apply(q, kv, B, T, S, H, dra(), scal if fuse_q else 1.0)
        r = _lin(W, a, p + "multihead_attn.out_proj.weight", p + "multihead_attn.out_proj.bias", residual=xf, dropout=dr(), res_link=lk)
        xf = _ln(W, p + "norm2", r)
        d1 = dr()                                    # relu' and this dropout's mask are applied by linear2's dX GEMM (bf16 path)
        lk = link()
        hdn = _lin(W, xf, p + "linear1.weight", p + "linear1.bias", relu=True, dropout=d1, pre_gated=fuse_gate, in_link=lk)
        r = _lin(W, hdn, p + "linear2.weight", p + "linear2.bias", residual=xf, dropout=dr(), res_link=lk,
                 in_gate_scale=(ops.drop_args(*d1[:2])[0] if d1 is not None else 1.0) if fuse_gate else None)
        xf = _ln(W, p + "norm3", r)
    xf = _ln(W, "transformer.decoder.norm", xf)
    y = _lin(W, xf, "Wout.weight", "Wout.bias", out_dtype=F32)
    return y.view(B, T, -1)


# ----------------------------------------------------------------------------------------------- variant blocks (fp32)
# Autograd through the MoE / GQA / generic-wrapper blocks (BASELINE config 4): forward AND backward are our kernels; the
# reference gets these gradients from torch autograd over model/moe.py, model/grouped_query_attention.py and
# model/custom_transformer.py:1220-1292.

def tracking(*things) -> bool:
    """True when a backward pass may be asked for: grad mode on and any tensor / module parameter requires grad."""
    if not torch.is_grad_enabled():
        return False
    for t in things:
        if isinstance(t, torch.Tensor):
            if t.requires_grad:
                return True
        elif isinstance(t, torch.nn.Module):
            if any(p.requires_grad for p in t.parameters()):
                return True
    return False


def rows_f32(x: torch.Tensor) -> torch.Tensor:
    """(..., d) -> contiguous fp32 (rows, d), staying in the autograd graph (view bookkeeping only)."""
    return x.reshape(-1, x.shape[-1]).float().contiguous()


def linear_fn(x2: torch.Tensor, lin: torch.nn.Linear, relu: bool = False) -> torch.Tensor:
    """nn.Linear on (rows, d) fp32 through LinearFn."""
    return LinearFn.apply(x2, lin.weight, lin.bias, lin.weight, x2.shape[1], relu, 1.0, 0, None, 0, F32, None)


def weight_bf16(lin: torch.nn.Linear) -> torch.Tensor:
    """bf16 copy of an nn.Linear weight for the tensor-core GEMMs, cached on the module and refreshed when the fp32 master changes."""
    w = lin.weight
    key = (w.data_ptr(), w._version)
    hit = lin.__dict__.get("_v2m_w16")
    if hit is None or hit[0] != key:
        hit = (key, w.detach().to(BF16).contiguous())
        lin.__dict__["_v2m_w16"] = hit
    return hit[1]


def linear_bf16_fn(x16: torch.Tensor, lin: torch.nn.Linear, *, alpha: float = 1.0, out_dtype=BF16, relu: bool = False) -> torch.Tensor:
    """nn.Linear on (rows, d) bf16 rows on the tcgen05 GEMM (fp32 master weights and gradients); alpha scales every output column
    (the 1 / sqrt(head_dim) of the query projection)."""
    N = lin.weight.shape[0]
    return LinearFn.apply(x16, lin.weight, lin.bias, weight_bf16(lin), x16.shape[1], relu, alpha, N if alpha != 1.0 else 0, None, 0,
                          out_dtype, None)


def bf16_ok(*lins) -> bool:
    """Shapes the tcgen05 GEMM takes for forward, dX and dW of these nn.Linear layers (16-byte rows: multiples of 8 both ways)."""
    return all(l.weight.shape[0] % 8 == 0 and l.weight.shape[1] % 8 == 0 for l in lins)


class AddFn(torch.autograd.Function):
    """a + alpha * b."""

    @staticmethod
    def forward(ctx, a, b, alpha):
        ctx.alpha = alpha
        return ops.axpy(a, b, alpha)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        return g, (g if ctx.alpha == 1.0 else ops.axpy(g, g, ctx.alpha - 1.0)), None


class SwigluFn(torch.autograd.Function):
    """a * silu(g)  (GLUExpert, moe.py:46-47)."""

    @staticmethod
    def forward(ctx, a, g):
        ctx.save_for_backward(a, g)
        return ops.swiglu(a, g)

    @staticmethod
    def backward(ctx, dh):
        a, g = ctx.saved_tensors
        dag = ops.swiglu_bwd(a, g, dh)
        ff = a.shape[1]
        return dag[:, :ff], dag[:, ff:]


class SigmoidFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a):
        s = ops.sigmoid(a)
        ctx.save_for_backward(s)
        return s

    @staticmethod
    def backward(ctx, dy):
        (s,) = ctx.saved_tensors
        return ops.sigmoid_bwd(dy.contiguous(), s)


class RMSNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, eps):
        ctx.save_for_backward(x, w)
        ctx.eps = eps
        return ops.rmsnorm(x, w, eps)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx, dw = ops.rmsnorm_bwd(x, w, dy, ctx.eps)
        return dx, dw, None


class RopeQuirkFn(torch.autograd.Function):
    """RoPE of the V2 / V3 attention with the reference's literal reinterpretation (custom_transformer.py:1044-1053; ops.rope_quirk): a
    rotation of every (even, odd) pair by a cache entry, so the backward is the same kernel with the sines negated."""

    @staticmethod
    def forward(ctx, x, cache, B, H):
        ctx.save_for_backward(cache)
        ctx.meta = (B, H)
        return ops.rope_quirk(x.contiguous(), cache, B, H)

    @staticmethod
    def backward(ctx, dy):
        (cache,) = ctx.saved_tensors
        inv = cache.clone()
        inv[..., 1].neg_()
        return ops.rope_quirk(dy.contiguous(), inv, *ctx.meta), None, None, None


class DropoutFn(torch.autograd.Function):
    """nn.Dropout in training mode on a (rows, cols) tensor: y = x o mask(seed) * 1/(1-p); the backward applies the same
    stateless mask to the incoming gradient (nothing is stored)."""

    @staticmethod
    def forward(ctx, x, p, seed):
        ctx.ps = (p, seed)
        return ops.dropout(x, p, seed)

    @staticmethod
    def backward(ctx, dy):
        return ops.dropout(dy.contiguous(), *ctx.ps), None, None


def drop_rows(x2: torch.Tensor, drop: torch.nn.Dropout, training: bool) -> torch.Tensor:
    """`drop(x)` of a reference forward on a 2-D tensor: identity in eval mode or with p = 0."""
    if not training or drop.p <= 0.0:
        return x2
    return DropoutFn.apply(x2.contiguous(), float(drop.p), ops.next_dropout_seed())


def drop_any(x: torch.Tensor, drop: torch.nn.Dropout, training: bool) -> torch.Tensor:
    if not training or drop.p <= 0.0:
        return x
    return drop_rows(x.reshape(-1, x.shape[-1]), drop, training).view(x.shape)


def has_dropout(module: torch.nn.Module) -> bool:
    """True when `module` is in training mode and any nn.Dropout inside it has p > 0."""
    return module.training and any(isinstance(m, torch.nn.Dropout) and m.p > 0 for m in module.modules())


def glu_expert_fn(e, x2: torch.Tensor) -> torch.Tensor:
    """GLUExpert.forward (moe.py:44-49) on (rows, d) with gradients: three LinearFn GEMMs around SwigluFn, the expert's
    dropout on the hidden rows in training mode (moe.py:48)."""
    h = SwigluFn.apply(linear_fn(x2, e.linear1), linear_fn(x2, e.gate))
    return linear_fn(drop_rows(h, e.dropout, e.training), e.linear2)


class MoEExpertsFn(torch.autograd.Function):
    """out[t] = sum_r softmax(top-k gate logits)[t,r] * expert_{idx[t,r]}(x[t])  (moe.py:180-199 / 244-299).
    The routing (idx, w, hist) is computed by the fused router outside; its gradient (softmax over the selected logits ->
    gate weight / bias / x) is produced here.  params = the experts' (linear1.weight, linear1.bias, gate.weight, gate.bias,
    linear2.weight, linear2.bias) in expert order; `stacks` = their [E, ...] copies."""

    @staticmethod
    def forward(ctx, x2, gate_w, gate_b, idx, w, hist, scale, stacks, drops, *params):
        w1, b1, wg, bg, w2, b2 = stacks
        out, saved = ops.moe_experts_fwd_saved(x2, idx, w, hist, w1, b1, wg, bg, w2, b2, drops=drops)
        ctx.save_for_backward(x2, gate_w, idx, w, w1, wg, w2, *saved)
        ctx.scale = scale
        ctx.drops = drops
        return out

    @staticmethod
    def backward(ctx, dout):
        x2, gate_w, idx, w, w1, wg, w2, *saved = ctx.saved_tensors
        E, ff, d = w1.shape
        w1g_t = torch.cat([w1, wg], 1).transpose(1, 2).contiguous()          # [E, d, 2 ff]
        w2_t = w2.transpose(1, 2).contiguous()                               # [E, ff, d_out]
        dx_e, dlogits, dW1g, db1g, dW2, db2 = ops.moe_experts_bwd(dout, tuple(saved), idx, w, ctx.scale, w1g_t, w2_t, E, drops=ctx.drops)
        dgate_w = _gemm_dw(dlogits, x2, d)
        _, dgate_b = ops.dy_prep(dlogits, None, False, 1.0, 0, F32, want_dz=False)
        dx = ops.axpy(dx_e, _gemm_dx(dlogits, gate_w.detach().contiguous(), d), 1.0)
        grads = []
        for e in range(E):
            grads += [dW1g[e, :ff], db1g[e, :ff], dW1g[e, ff:], db1g[e, ff:], dW2[e], db2[e]]
        return (dx, dgate_w, dgate_b, None, None, None, None, None, None, *grads)


class MoEExpertsBf16Fn(torch.autograd.Function):
    """MoEExpertsFn on the bf16 tensor-core path: forward = the grouped tcgen05 GEMMs of the inference path with their
    intermediates kept, backward = grouped GEMMs against the transposed weight stacks (dX) and K-grouped GEMMs over the ragged
    expert groups (dW), all bf16 operands with fp32 accumulation; routing weights, the combine and every gradient that
    reaches a parameter stay fp32.  stacks16 = (w1g bf16 [E, 2 ff, d], b1g fp32, w2 bf16 [E, d_out, ff], b2 fp32)."""

    @staticmethod
    def forward(ctx, x2, gate_w, gate_b, idx, w, hist, scale, stacks16, *params):
        w1g, b1g, w2, b2 = stacks16
        out, saved = ops.moe_experts_bf16_fwd_saved(x2, idx, w, hist, w1g, b1g, w2, b2)
        ctx.save_for_backward(x2, gate_w, idx, w, w1g, w2, *saved)
        ctx.scale = scale
        return out

    @staticmethod
    def backward(ctx, dout):
        x2, gate_w, idx, w, w1g, w2, *saved = ctx.saved_tensors
        E, ff2, d = w1g.shape
        ff = ff2 // 2
        w1g_t = w1g.transpose(1, 2).contiguous()                             # bf16 [E, d, 2 ff]
        w2_t = w2.transpose(1, 2).contiguous()                               # bf16 [E, ff, d_out]
        dx_e, dlogits, dW1g, db1g, dW2, db2 = ops.moe_experts_bf16_bwd(dout, tuple(saved), idx, w, ctx.scale, w1g_t, w2_t, E)
        dgate_w = _gemm_dw(dlogits, x2, d)                                   # the router stays fp32 (6 columns)
        _, dgate_b = ops.dy_prep(dlogits, None, False, 1.0, 0, F32, want_dz=False)
        dx = ops.axpy(dx_e, _gemm_dx(dlogits, gate_w.detach().contiguous(), d), 1.0)
        grads = []
        for e in range(E):
            grads += [dW1g[e, :ff], db1g[e, :ff], dW1g[e, ff:], db1g[e, ff:], dW2[e], db2[e]]
        return (dx, dgate_w, dgate_b, None, None, None, None, None, *grads)


class GluBf16Fn(torch.autograd.Function):
    """One GLUExpert over all rows (the shared expert, moe.py:301) on the tcgen05 GEMMs with gradients: x fp32 (rows, d) ->
    linear2((linear1 x) * silu(gate x)) fp32; stacks16 as above with E = 1."""

    @staticmethod
    def forward(ctx, x2, stacks16, w1, b1, wg, bg, w2p, b2p):
        w1g, b1g, w2, b2 = stacks16
        x16 = ops.cast_2d(x2, BF16)
        a = ops.linear(x16, w1g[0], b1g[0], out_dtype=BF16)
        h = ops.swiglu_pair(a)
        ctx.save_for_backward(x16, a, h, w1g, w2)
        return ops.linear(h, w2[0], b2[0], out_dtype=F32)

    @staticmethod
    def backward(ctx, dy):
        x16, a, h, w1g, w2 = ctx.saved_tensors
        ff = h.shape[1]
        d = x16.shape[1]
        dy16, db2 = ops.dy_prep(dy.contiguous(), None, False, 1.0, 0, BF16, want_dz=True)
        dW2 = _gemm_dw(dy16, h, ff)
        dh = _gemm_dx(dy16, w2[0], ff)
        dag = ops.swiglu_pair_bwd(a, dh.contiguous())
        _, db1g = ops.dy_prep(dag, None, False, 1.0, 0, BF16, want_dz=False)
        dW1g = _gemm_dw(dag, x16, d)
        dx = _gemm_dx(dag, w1g[0], d).float()
        return dx, None, dW1g[:ff], db1g[:ff], dW1g[ff:], db1g[ff:], dW2, db2


def glu_expert_bf16_fn(e, x2: torch.Tensor, stacks16) -> torch.Tensor:
    return GluBf16Fn.apply(x2, stacks16, e.linear1.weight, e.linear1.bias, e.gate.weight, e.gate.bias, e.linear2.weight, e.linear2.bias)


def moe_experts_bf16_fn(experts, gate, x2, idx, w, hist, scale, stacks16):
    params = [p for e in experts for p in (e.linear1.weight, e.linear1.bias, e.gate.weight, e.gate.bias, e.linear2.weight, e.linear2.bias)]
    k = idx.shape[-1]
    return MoEExpertsBf16Fn.apply(x2, gate.weight, gate.bias, idx.reshape(-1, k), w.reshape(-1, k), hist, scale, stacks16, *params)


def moe_experts_fn(experts, gate, x2, idx, w, hist, scale, stacks, layer_dropout=None, training=False):
    """layer_dropout: the MoE layer's nn.Dropout (applied to every expert output row, moe.py:197); the experts' own dropout
    (hidden rows, moe.py:48) is read from experts[0] (all experts are clones of one GLUExpert)."""
    params = [p for e in experts for p in (e.linear1.weight, e.linear1.bias, e.gate.weight, e.gate.bias, e.linear2.weight, e.linear2.bias)]
    k = idx.shape[-1]
    drops = None
    if training:
        p_h = float(experts[0].dropout.p)
        p_o = float(layer_dropout.p) if layer_dropout is not None else 0.0
        if p_h > 0 or p_o > 0:
            drops = (p_h, ops.next_dropout_seed(), p_o, ops.next_dropout_seed())
    return MoEExpertsFn.apply(x2, gate.weight, gate.bias, idx.reshape(-1, k), w.reshape(-1, k), hist, scale, stacks, drops, *params)


class GqaAttnFn(torch.autograd.Function):
    """scaled_dot_product_gqa (grouped_query_attention.py:19-170): q (b,n,hq,d), k/v (b,s,hk,d) contiguous fp32 ->
    out (n,b,hq,d) (sequence-first, :159); query head h*g+gi reads kv head h."""

    @staticmethod
    def forward(ctx, q, k, v, causal, q_scale, dropout=None):
        """bf16 tensors (head_dim 64, q pre-scaled: q_scale == 1) run on the tcgen05 forward and the tensor-core backward."""
        b, n, hq, d = q.shape
        s, hk = k.shape[1], k.shape[2]
        assert q.dtype == F32 or (d == 64 and q_scale == 1.0), "bf16 grouped-query attention needs head_dim 64 and pre-scaled queries"
        out = torch.empty((n, b, hq, d), device=q.device, dtype=q.dtype)
        lse = torch.empty((b * hq, n), device=q.device, dtype=F32)
        ops.attention(q, k, v, out, B=b, Hq=hq, Hkv=hk, Lq=n, Lk=s, dh=d, q_strides=(n * hq * d, hq * d),
                      k_strides=(s * hk * d, hk * d), v_strides=(s * hk * d, hk * d), o_strides=(hq * d, b * hq * d),
                      causal=causal, q_scale=q_scale, lse=lse, dropout=dropout)
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.meta = (causal, q_scale)
        ctx.dropout = dropout
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, lse = ctx.saved_tensors
        causal, q_scale = ctx.meta
        b, n, hq, d = q.shape
        s, hk = k.shape[1], k.shape[2]
        dout = dout.contiguous()
        dq = torch.empty_like(q)
        tc = q.dtype == BF16                                               # tensor-core kernels write bf16 dK / dV once (summed over the group)
        dk, dv = (torch.empty_like(k), torch.empty_like(v)) if tc else (torch.zeros_like(k), torch.zeros_like(v))
        ops.attention_bwd(q, k, v, out, dout, lse, None, dq, dk, dv, None, B=b, Hq=hq, Hkv=hk, Lq=n, Lk=s, dh=d,
                          q_strides=(n * hq * d, hq * d), k_strides=(s * hk * d, hk * d), v_strides=(s * hk * d, hk * d),
                          o_strides=(hq * d, b * hq * d), do_strides=(hq * d, b * hq * d), dq_strides=(n * hq * d, hq * d),
                          dkv_strides=(s * hk * d, hk * d), causal=causal, q_scale=q_scale, dropout=ctx.dropout, tensor_core=tc)
        return dq, dk, dv, None, None, None


class MambaCoreFn(torch.autograd.Function):
    """xz (B*L, 2 ED) = in_proj output (x | z)  ->  y (B*L, ED): depthwise conv + SiLU of the x half, x_proj, dt_proj and the
    fused selective scan gated by the z half (mamba.py:264-351).  Backward: selective_scan_bwd (states recomputed, no
    (B, L, ED, N) autograd graph), four strided GEMMs for the two small projections, conv/SiLU backward."""

    @staticmethod
    def forward(ctx, xz, conv_w, conv_b, xproj_w, dtproj_w, dt_bias, A_log, D, B, L, plus):
        ED, N = A_log.shape
        R = dtproj_w.shape[1]
        xc = ops.mamba_conv_silu(xz, ED, conv_w.contiguous(), conv_b, B, L)
        dbc = ops.linear(xc, xproj_w.contiguous())
        draw = ops.linear(dbc[:, :R], dtproj_w.contiguous())
        y = ops.selective_scan(xc, draw, dt_bias, A_log, dbc[:, R:R + N], dbc[:, R + N:], D, xz[:, ED:], B, L, plus=plus)
        ctx.save_for_backward(xz, conv_w, conv_b, xproj_w, dtproj_w, dt_bias, A_log, D, xc, dbc, draw)
        ctx.meta = (B, L, plus)
        return y

    @staticmethod
    def backward(ctx, dy):
        xz, conv_w, conv_b, xproj_w, dtproj_w, dt_bias, A_log, D, xc, dbc, draw = ctx.saved_tensors
        B, L, plus = ctx.meta
        ED, N = A_log.shape
        R = dtproj_w.shape[1]
        M, P = xc.shape[0], dbc.shape[1]
        dy = dy.contiguous()
        dxz = torch.empty_like(xz)
        ddbc = torch.zeros_like(dbc)
        dxc, ddraw, dA_log, dD, ddtb = ops.selective_scan_bwd(xc, draw, dt_bias, A_log, dbc[:, R:R + N], dbc[:, R + N:], D, xz[:, ED:],
                                                              dy, ddbc[:, R:R + N], ddbc[:, R + N:], dxz[:, ED:], B, L, plus=plus)
        wdt, wx = dtproj_w.contiguous(), xproj_w.contiguous()
        ops.gemm_strided(ddraw, ED, 1, wdt, 1, R, M, R, ED, out=ddbc[:, :R])                  # d(dbc[:, :R]) = ddraw @ Wdt
        dwdt = _gemm_dw(ddraw, dbc, R)                                                        # ddraw^T @ dbc[:, :R]
        dxc = ops.axpy(dxc, ops.gemm_strided(ddbc, P, 1, wx, 1, ED, M, ED, P), 1.0)           # + ddbc @ Wx
        dwx = _gemm_dw(ddbc, xc, ED)                                                          # ddbc^T @ xc
        dconv_w, dconv_b = ops.mamba_conv_silu_bwd(xz, ED, conv_w.contiguous(), conv_b, dxc, dxz, B, L)
        return dxz, dconv_w, dconv_b, dwx, dwdt, ddtb, dA_log, dD, None, None, None


class AttnRowsFn(torch.autograd.Function):
    """Attention over separate q (rows_q, E), k, v (rows_k, E) fp32 projections with explicit (batch, row) strides -- the
    (L, B, E) module layout has rows ordered (l, b).  q arrives pre-scaled.  Optional Er (RPR skew) and causal mask; with
    need_p the per-head probabilities (B*H, Lq, Lk) are returned as a second, non-differentiable output."""

    @staticmethod
    def forward(ctx, q, k, v, er, B, Lq, Lk, H, causal, qs, ks, need_p, dropout=None):
        E = q.shape[1]
        ctx.dropout = dropout
        out = torch.empty_like(q)
        lse = torch.empty((B * H, Lq), device=q.device, dtype=F32)
        p_out = torch.empty((B * H, Lq, Lk), device=q.device, dtype=F32) if need_p else None
        ops.attention(q, k, v, out, B=B, Hq=H, Hkv=H, Lq=Lq, Lk=Lk, dh=E // H, q_strides=qs, k_strides=ks, v_strides=ks, o_strides=qs,
                      causal=causal, Er=er, lse=lse, p_out=p_out, dropout=dropout)
        ctx.save_for_backward(q, k, v, out, lse, er)
        ctx.meta = (B, Lq, Lk, H, causal, qs, ks)
        if need_p:
            ctx.mark_non_differentiable(p_out)
            return out, p_out
        return out, None

    @staticmethod
    def backward(ctx, dout, _dp):
        q, k, v, out, lse, er = ctx.saved_tensors
        B, Lq, Lk, H, causal, qs, ks = ctx.meta
        E = q.shape[1]
        dout = dout.contiguous()
        dq, dk, dv = torch.empty_like(q), torch.zeros_like(k), torch.zeros_like(v)
        der = torch.zeros_like(er) if er is not None else None
        ops.attention_bwd(q, k, v, out, dout, lse, er, dq, dk, dv, der, B=B, Hq=H, Hkv=H, Lq=Lq, Lk=Lk, dh=E // H, q_strides=qs,
                          k_strides=ks, v_strides=ks, o_strides=qs, do_strides=qs, dq_strides=qs, dkv_strides=ks, causal=causal,
                          dropout=ctx.dropout)
        return dq, dk, dv, der, None, None, None, None, None, None, None, None, None


def mha_rpr_autograd(module, query, key, value, need_weights, attn_mask):
    """MultiheadAttentionRPR.forward (rpr.py:170-198 -> multi_head_attention_forward_rpr :201-424) with gradients, fp32:
    three projection GEMMs (q pre-scaled in the epilogue), AttnRowsFn (fused RPR attention forward / backward), out-projection."""
    from .rpr import is_causal_mask
    L, B, E = query.shape
    S = key.shape[0]
    H = module.num_heads
    assert E == module.embed_dim and key.shape == value.shape
    assert key is value or torch.equal(key, value), "separate key / value tensors are not on the AMT path"
    if getattr(module, "compute_dtype", F32) != F32:
        raise NotImplementedError("module-level autograd runs on the fp32 path; bf16 training goes through VideoMusicTransformer "
                                  "(amt_forward_autograd)")
    causal = is_causal_mask(attn_mask, L)
    W, bias = module.in_proj_weight, module.in_proj_bias
    xq = rows_f32(query)
    xk = xq if key is query else rows_f32(key)

    def proj(x, lo, alpha=1.0, alpha_cols=0):
        w = W[lo:lo + E]
        return LinearFn.apply(x, w, None if bias is None else bias[lo:lo + E], w, E, False, alpha, alpha_cols, None, 0, F32, None)

    q = proj(xq, 0, float(E // H) ** -0.5, E)                                    # rpr.py:328 scaling
    k, v = proj(xk, E), proj(xk, 2 * E)
    er = None
    if module.Er is not None:
        if L != S or L > module.Er.shape[0]:
            raise RuntimeError("RPR attention needs len_q == len_k <= er_len (got %d, %d, er_len %d); the reference "
                               "fails in _skew for longer inputs (rpr.py:426-450)" % (L, S, module.Er.shape[0]))
        er = module.Er
    drop = (float(module.dropout), ops.next_dropout_seed()) if (module.training and module.dropout > 0) else None   # rpr.py:412
    ctxv, p = AttnRowsFn.apply(q, k, v, er, B, L, S, H, causal, (E, B * E), (E, B * E), bool(need_weights), drop)
    out = linear_fn(ctxv, module.out_proj).view(L, B, E)
    return out, (p.view(B, H, L, S).sum(dim=1) / H if need_weights else None)           # rpr.py:419-422


def custom_mha_autograd(module, query, key, value, need_weights, attn_mask, average_attn_weights):
    """CustomMultiheadAttention.forward (custom_transformer.py:51-321 -> custom_multi_head_attention_forward :864-1218) with gradients,
    fp32: projections (q pre-scaled in the epilogue: the scaling commutes with the rotation), RoPE with the literal reinterpretation,
    AttnRowsFn, out-projection.  Training of the V1 / V2 model zoo."""
    from .rpr import is_causal_mask
    L, B, E = query.shape
    S = key.shape[0]
    H, dh = module.num_heads, module.head_dim
    causal = is_causal_mask(attn_mask, L) if attn_mask is not None else False
    W, bias = module.in_proj_weight, module.in_proj_bias
    xq = rows_f32(query)
    xk = xq if key is query else rows_f32(key)
    xv = xk if value is key else rows_f32(value)

    def proj(x, lo, alpha=1.0, alpha_cols=0):
        w = W[lo:lo + E]
        return LinearFn.apply(x, w, bias[lo:lo + E], w, E, False, alpha, alpha_cols, None, 0, F32, None)

    q = proj(xq, 0, float(dh) ** -0.5, E)
    k, v = proj(xk, E), proj(xv, 2 * E)
    if module.RoPE is not None:                                                       # custom_transformer.py:1047-1050
        cache = module.RoPE.cache
        q = RopeQuirkFn.apply(q, cache[:L].contiguous(), B, H)
        k = RopeQuirkFn.apply(k, cache[:S].contiguous(), B, H)
    drop = (float(module.dropout), ops.next_dropout_seed()) if (module.training and module.dropout > 0) else None
    ctxv, p = AttnRowsFn.apply(q, k, v, None, B, L, S, H, causal, (E, B * E), (E, B * E), bool(need_weights), drop)
    out = linear_fn(ctxv, module.out_proj).view(L, B, E)
    if not need_weights:
        return out, None
    wts = p.view(B, H, L, S)
    return out, (wts.mean(dim=1) if average_attn_weights else wts)


def diff_mha_autograd(module, query, key, value, attn_mask):
    """DifferentialMultiheadAttention.forward (custom_transformer.py:610-832) with gradients, fp32: two attentions over the even / odd
    half-heads sharing the values (AttnRowsFn on the literal batch-first re-view of the projections), their difference weighted by
    the differentiable lambda, per-head RMSNorm scaled by (1 - lambda_init), the literal (B, H, L, d) -> (L, B, E) re-view, out-projection."""
    L, B, E = query.shape
    S = key.shape[0]
    H, dh = module.num_heads, module.head_dim
    if module.training and module.dropout.p > 0:
        raise NotImplementedError("DifferentialMultiheadAttention trains with dropout 0 (its attention-weight dropout is not built)")
    xq = rows_f32(query)
    xk = xq if key is query else rows_f32(key)
    xv = xk if value is key else rows_f32(value)
    wq, wk = module.q_proj.weight, module.k_proj.weight
    q = LinearFn.apply(xq, wq, None, wq, E, False, float(module.scaling), 2 * E, None, 0, F32, None)      # (L*B, 2E), pre-scaled
    k = linear_fn(xk, module.k_proj)
    v = linear_fn(xv, module.v_proj)                                                  # (S*B, E)
    if module.RoPE is not None:
        q = RopeQuirkFn.apply(q, module.RoPE.cache[:L].contiguous(), B, 2 * H)
        k = RopeQuirkFn.apply(k, module.RoPE.cache[:S].contiguous(), B, 2 * H)
    q5, k5 = q.view(B, L, H, 2, dh), k.view(B, S, H, 2, dh)                           # the same memory read batch-first (:786-788)
    causal = attn_mask is not None
    outs = []
    for i in (0, 1):
        qi = q5[:, :, :, i].contiguous().view(B * L, E)
        ki = k5[:, :, :, i].contiguous().view(B * S, E)
        o, _ = AttnRowsFn.apply(qi, ki, v, None, B, L, S, H, causal, (L * E, E), (S * E, E), False, None)
        outs.append(o)
    lam = (torch.exp(torch.sum(module.lambda_q1 * module.lambda_k1)) - torch.exp(torch.sum(module.lambda_q2 * module.lambda_k2))
           + module.lambda_init)                                                      # :811-813, differentiable
    diff = outs[0] - lam * outs[1]
    w = module.subln.weight * (1.0 - module.lambda_init)
    attn = RMSNormFn.apply(diff.view(-1, dh).contiguous(), w.contiguous(), module.subln.eps).view(B, L, H, dh)
    attn = attn.permute(0, 2, 1, 3).contiguous().view(L * B, E)                       # (B, H, L, d) memory read as (L, B, E) rows (:824)
    return linear_fn(attn, module.out_proj).view(L, B, E), None
