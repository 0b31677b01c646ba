"""Drop-in for model/grouped_query_attention.py: `scaled_dot_product_gqa` (:19-170) and `MultiheadGQA`
(:172-358), including the literal behaviours SURVEY.md 7.2-1 lists: the module ignores attn_mask
(:339), is causal only when is_causal=True, reinterprets the (L,B,E) projections as (B,L,E) with
`.view` (:316-326), and the function returns the output sequence-first (n, b, h, d) (:159).
Query head hq reads kv head hq // (Hq/Hkv)  ("b (h g) n d -> b g h n d", :126)."""
from typing import Optional

import torch
import torch.nn as nn

from . import ops


def scaled_dot_product_gqa(query, key, value, num_heads=None, dropout: float = 0.0, scale: Optional[float] = None,
                           attn_mask=None, key_padding_mask=None, is_causal: Optional[bool] = None,
                           need_weights: bool = False, average_attn_weights: bool = False):
    if (attn_mask is not None) and (is_causal is not None):
        raise ValueError("Only one of 'attn_mask' and 'is_causal' should be provided, but got both.")
    elif not query.ndim == key.ndim == value.ndim == 4:
        raise ValueError(f"Expected query, key, and value to be 4-dimensional, but got shapes "
                         f"{query.shape}, {key.shape}, and {value.shape}.")
    if attn_mask is not None or key_padding_mask is not None or need_weights:
        raise NotImplementedError("attn_mask / key_padding_mask / need_weights are never used by the reference's "
                                  "callers (grouped_query_attention.py:333-342)")
    b, n, hq, d = query.shape
    bk, s, hk, dk = key.shape
    if not (b == bk == value.shape[0] and d == dk == value.shape[3]):
        raise ValueError("Expected query, key, and value to have the same batch size (dim=0) and embedding dimension (dim=3)")
    elif key.shape[1:3] != value.shape[1:3]:
        raise ValueError("Expected key and value to have the same size in dimensions 1 and 2")
    elif hq % hk != 0:
        raise ValueError("Expected query heads to be a multiple of key/value heads")
    if scale is None:
        scale = d ** 0.5
    dt = query.dtype
    from . import autograd as ag
    # dropout > 0: F.dropout(attention, p) with its default training=True, i.e. ALWAYS applied (grouped_query_attention.py:152-153;
    # MultiheadGQA never passes it, :333-342)
    drop = (float(dropout), ops.next_dropout_seed()) if dropout > 0.0 else None
    if (ag.tracking(query, key, value) or drop is not None) and dt == torch.float32:   # gradients through our attention backward kernel
        return ag.GqaAttnFn.apply(query.contiguous(), key.contiguous(), value.contiguous(), bool(is_causal), 1.0 / scale, drop), None
    if ag.tracking(query, key, value) and drop is None and dt == torch.bfloat16 and d == 64:
        # tensor-core path: the kernels want pre-scaled queries (the scaling stays in the autograd graph)
        return ag.GqaAttnFn.apply((query * (1.0 / scale)).contiguous(), key.contiguous(), value.contiguous(), bool(is_causal), 1.0, None), None
    if drop is not None:
        raise NotImplementedError("scaled_dot_product_gqa: dropout > 0 is built on the fp32 path")
    q, k, v = (t.detach().contiguous() for t in (query, key, value))
    out = torch.empty((n, b, hq, d), device=q.device, dtype=dt)          # sequence-first, :159
    ops.attention(q, k, v, out, B=b, Hq=hq, Hkv=hk, Lq=n, Lk=s, dh=d,
                  q_strides=(n * hq * d, hq * d), k_strides=(s * hk * d, hk * d), v_strides=(s * hk * d, hk * d),
                  o_strides=(hq * d, b * hq * d), causal=bool(is_causal), q_scale=1.0 / scale)
    return out, None


class MultiheadGQA(nn.Module):
    def __init__(self, embed_dim, query_heads, kv_heads, dropout=0.0, bias=True, layer_norm=True,
                 layer_norm_eps=1e-5, gamma_init=1.0, device=None, dtype=None, RoPE=None):
        super().__init__()
        if RoPE is not None:
            raise NotImplementedError("RoPE inside MultiheadGQA is SURVEY.md 8f row 2")
        self.query_heads = query_heads
        self.kv_heads = kv_heads
        self.dropout = dropout
        self.layer_norm = layer_norm
        self.gamma_init = gamma_init
        self.RoPE = None
        self.embed_dim = embed_dim
        if self.query_heads % self.kv_heads != 0:
            raise ValueError(f"query_heads ({query_heads}) must be divisible by kv_heads ({kv_heads})")
        elif (embed_dim % self.query_heads != 0) or (embed_dim % self.kv_heads != 0):
            raise ValueError(f"embed_dim ({embed_dim}) must be divisible by query_heads ({query_heads}) and kv_heads ({kv_heads})")
        head_dim = embed_dim // query_heads
        if not head_dim % 8 == 0:
            raise ValueError(f"head_dim (embed_dim / num_heads = {head_dim}) must be divisible by 8")
        if not head_dim <= 128:
            raise ValueError(f"head_dim (embed_dim / num_heads = {head_dim}) must be <= 128")
        self.q_proj = nn.Linear(embed_dim, embed_dim, bias=bias, device=device, dtype=dtype)
        kv_embed_dim = embed_dim // query_heads * kv_heads
        self.k_proj = nn.Linear(embed_dim, kv_embed_dim, bias=bias, device=device, dtype=dtype)
        self.v_proj = nn.Linear(embed_dim, kv_embed_dim, bias=bias, device=device, dtype=dtype)
        self.norm = nn.LayerNorm(embed_dim, eps=layer_norm_eps, device=device, dtype=dtype) if layer_norm else None
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=bias, device=device, dtype=dtype)
        self._reset_parameters()

    def _reset_parameters(self):                                       # grouped_query_attention.py:262-283
        nn.init.xavier_normal_(self.q_proj.weight)
        nn.init.xavier_normal_(self.k_proj.weight)
        nn.init.xavier_normal_(self.v_proj.weight, gain=self.gamma_init)
        nn.init.xavier_normal_(self.out_proj.weight, gain=self.gamma_init)
        for lin in (self.q_proj, self.k_proj, self.v_proj, self.out_proj):
            if lin.bias is not None:
                nn.init.constant_(lin.bias, 0)

    def forward(self, query, key, value, need_weights=False, attn_mask=None, key_padding_mask=None,
                is_causal=False, average_attn_weights=False):
        if need_weights or key_padding_mask is not None:
            raise NotImplementedError("need_weights / key_padding_mask are not used by the reference's callers")
        tgt_len, bsz, E = query.shape
        src_len = key.shape[0]
        H, Hk = self.query_heads, self.kv_heads
        dh = E // H

        from . import autograd as ag
        track = ag.tracking(query, key, value, self)

        def lin(m, x):
            if track:
                return ag.linear_fn(ag.rows_f32(x), m)
            x2 = x.detach().reshape(-1, x.shape[-1]).float().contiguous()
            return ops.linear(x2, m.weight.detach(), m.bias.detach() if m.bias is not None else None, k=x2.shape[1])

        bf16 = torch.bfloat16
        tc = (track and getattr(self, "compute_dtype", torch.float32) == bf16 and dh == 64 and E % 8 == 0
              and not getattr(self, "batch_independent", False))
        if tc:
            # tensor-core training path (compute_dtype = bf16; fp32 master weights, fp32 gradients): projections on the tcgen05 GEMM
            # with the 1 / sqrt(d) of scaled_dot_product_gqa (:93-96) folded into the query projection, tcgen05 attention forward,
            # tensor-core attention backward (dK / dV summed over the query heads of a group), bf16 LayerNorm, fp32 output
            rows16 = lambda x: ag.rows_f32(x).to(bf16)
            xq = rows16(query)
            same = key is query and value is query
            xk, xv = (xq, xq) if same else (rows16(key), rows16(value))
            q = ag.linear_bf16_fn(xq, self.q_proj, alpha=float(dh) ** -0.5)
            k, v = ag.linear_bf16_fn(xk, self.k_proj), ag.linear_bf16_fn(xv, self.v_proj)
            q4, k4, v4 = q.view(bsz, tgt_len, H, dh), k.view(bsz, src_len, Hk, dh), v.view(bsz, src_len, Hk, dh)   # literal .view, :324-326
            causal = bool(is_causal) or bool(getattr(self, "force_causal", False))
            x = ag.GqaAttnFn.apply(q4, k4, v4, causal, 1.0, None)          # (n, b, hq, d), :159
            x2 = x.reshape(-1, E)
            if self.layer_norm:
                x2 = ag.LayerNormFn.apply(x2.contiguous(), self.norm.weight, self.norm.bias, self.norm.eps)
            y = ag.linear_bf16_fn(x2, self.out_proj, out_dtype=torch.float32)
            return y.view(x.shape[0], x.shape[1], E), None
        q, k, v = lin(self.q_proj, query), lin(self.k_proj, key), lin(self.v_proj, value)   # :306-308
        # literal `.view(bsz, len, heads*dh)` of the (len, bsz, .) projections (:324-326)
        if getattr(self, "batch_independent", False) and bsz > 1:
            # our extension (off by default): every video of the batch is treated as the reference treats a batch of one --
            # the literal `.view` below interleaves batch and time for bsz > 1, which batched generation of
            # independent videos must not do.  For bsz == 1 the two are the same tensor.
            q4 = q.view(tgt_len, bsz, H, dh).transpose(0, 1).contiguous()
            k4 = k.view(src_len, bsz, Hk, dh).transpose(0, 1).contiguous()
            v4 = v.view(src_len, bsz, Hk, dh).transpose(0, 1).contiguous()
        else:
            q4 = q.view(bsz, tgt_len, H, dh)
            k4 = k.view(bsz, src_len, Hk, dh)
            v4 = v.view(bsz, src_len, Hk, dh)
        # force_causal (our extension, off by default): the reference's wrappers hand the causal mask over as attn_mask, which
        # this module drops (grouped_query_attention.py:339 "attn_mask=None"), so a literal GQA decoder sees future positions;
        # the config-4 shell (VideoMusicTransformer_GQA) turns this on for decoder self-attention
        is_causal = bool(is_causal) or bool(getattr(self, "force_causal", False))
        x, _ = scaled_dot_product_gqa(q4, k4, v4, num_heads=H, is_causal=True if is_causal else None)
        x = x.reshape(x.shape[0], x.shape[1], H * dh)                   # labelled "b n (h d)" by the reference (:343)
        if track:
            x2 = x.reshape(-1, E)
            if self.layer_norm:
                x2 = ag.LayerNormFn.apply(x2.contiguous(), self.norm.weight, self.norm.bias, self.norm.eps)
            return ag.linear_fn(x2, self.out_proj).view(x.shape[0], x.shape[1], E), None
        if self.layer_norm:
            x = ops.layernorm(x.contiguous(), self.norm.weight.detach(), self.norm.bias.detach(), eps=self.norm.eps)
        y = ops.linear(x.reshape(-1, E), self.out_proj.weight.detach(),
                       self.out_proj.bias.detach() if self.out_proj.bias is not None else None, k=E)
        return y.view(x.shape[0], x.shape[1], E), None
