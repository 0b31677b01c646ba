"""Drop-in modules for model/rpr.py of the reference: same class names, constructor signatures,
parameter names/shapes (state_dict keys) and (L, B, E) seq-first tensor layout, with the arithmetic
done by the sm_100a kernels.

  MultiheadAttentionRPR          <- rpr.py:112-198  (+ multi_head_attention_forward_rpr :201-424)
  TransformerDecoderLayerRPR     <- rpr.py:37-70
  TransformerDecoderRPR          <- rpr.py:17-35
  TransformerEncoderLayerRPR/…   <- rpr.py:73-110 (MusicTransformer only; parameter containers + forward)

Not supported (raise, like the rest of the package there is no silent fallback): key_padding_mask,
add_bias_kv / add_zero_attn, kdim/vdim != embed_dim, attention masks other than the causal
`generate_square_subsequent_mask`, dropout > 0 in training mode.
"""
import copy
from typing import Optional

import torch
import torch.nn as nn
from torch.nn import Parameter
import weakref

from torch.nn.init import constant_, xavier_uniform_

from . import ops


def _get_clones(module, n):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(n)])


_mask_cache = {}      # id(mask tensor) -> (weakref to it, version, L, verdict); entries die with the tensor


def is_causal_mask(attn_mask: Optional[torch.Tensor], L: int) -> bool:
    """True for the float mask of nn.Transformer.generate_square_subsequent_mask (0 on/below the
    diagonal, -inf above); None -> False; anything else is rejected.
    The verdict is cached per tensor object (weak reference) and in-place version: a different mask that the caching
    allocator later places at the same address is a different object and is checked again (the reference rebuilds its
    mask every forward, video_music_transformer.py:1033, so an address-keyed cache would go stale)."""
    if attn_mask is None:
        return False
    key = id(attn_mask)
    hit = _mask_cache.get(key)
    if hit is None or hit[0]() is not attn_mask or hit[1] != attn_mask._version or hit[2] != L:
        ok = attn_mask.shape == (L, L)
        if ok:
            ref = torch.triu(torch.full((L, L), float("-inf"), device=attn_mask.device), diagonal=1)
            ok = bool(torch.equal(attn_mask.float(), ref))
        hit = (weakref.ref(attn_mask, lambda _r, k=key: _mask_cache.pop(k, None)), attn_mask._version, L, ok)
        _mask_cache[key] = hit
    if not hit[3]:
        raise NotImplementedError("only the causal square-subsequent attn_mask (or None) is supported")
    return True


def _compute_dtype(module: nn.Module) -> torch.dtype:
    return getattr(module, "compute_dtype", torch.float32)


class _Cast:
    """bf16 copies of fp32 parameters, refreshed when the parameter changes."""

    def __init__(self):
        self.cache = {}

    def get(self, p: torch.Tensor, dtype: torch.dtype, rows: Optional[slice] = None) -> torch.Tensor:
        src = p.detach()
        if rows is not None:
            src = src[rows]
        if dtype == torch.float32:
            return src
        key = (p.data_ptr(), p._version, rows.start if rows else None, rows.stop if rows else None)
        t = self.cache.get(key)
        if t is None:
            if len(self.cache) > 16:
                self.cache.clear()
            t = ops.cast_2d(src, torch.bfloat16, (src.shape[1] + 7) // 8 * 8)
            self.cache[key] = t
        return t


class MultiheadAttentionRPR(nn.Module):
    def __init__(self, embed_dim, num_heads, dropout=0., bias=True, add_bias_kv=False, add_zero_attn=False,
                 kdim=None, vdim=None, er_len=None):
        super().__init__()
        self.embed_dim = embed_dim
        self.kdim = kdim if kdim is not None else embed_dim
        self.vdim = vdim if vdim is not None else embed_dim
        self._qkv_same_embed_dim = self.kdim == embed_dim and self.vdim == embed_dim
        if not self._qkv_same_embed_dim or add_bias_kv or add_zero_attn:
            raise NotImplementedError("kdim/vdim != embed_dim, add_bias_kv and add_zero_attn are not on the AMT path")
        self.num_heads = num_heads
        self.dropout = dropout
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == self.embed_dim, "embed_dim must be divisible by num_heads"
        self.in_proj_weight = Parameter(torch.empty(3 * embed_dim, embed_dim))
        if bias:
            self.in_proj_bias = Parameter(torch.empty(3 * embed_dim))
        else:
            self.register_parameter("in_proj_bias", None)
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.bias_k = self.bias_v = None
        self.add_zero_attn = add_zero_attn
        if er_len is not None:
            self.Er = Parameter(torch.rand((er_len, self.head_dim), dtype=torch.float32))   # rpr.py:147-150
        else:
            self.Er = None
        self._reset_parameters()
        self._cast = _Cast()

    def _reset_parameters(self):                                                            # rpr.py:154-168
        xavier_uniform_(self.in_proj_weight)
        if self.in_proj_bias is not None:
            constant_(self.in_proj_bias, 0.)
            constant_(self.out_proj.bias, 0.)

    def __deepcopy__(self, memo):
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = _Cast() if k == "_cast" else copy.deepcopy(v, memo)
        return new

    def forward(self, query, key, value, key_padding_mask=None, need_weights=True, attn_mask=None, **kwargs):
        if key_padding_mask is not None:
            raise NotImplementedError("key_padding_mask is not used by the AMT path (rpr.py:55-57 passes None)")
        drop = self.training and self.dropout > 0            # dropout of the probabilities (rpr.py:412), training mode
        if drop and _compute_dtype(self) != torch.float32:
            raise NotImplementedError("module-level attention dropout runs on the fp32 path; bf16 training with dropout goes "
                                      "through VideoMusicTransformer (autograd.amt_forward_autograd)")
        if drop or (torch.is_grad_enabled() and (query.requires_grad or self.in_proj_weight.requires_grad)):
            from .autograd import mha_rpr_autograd
            return mha_rpr_autograd(self, query, key, value, need_weights, attn_mask)
        return self._forward_impl(query, key, value, need_weights, attn_mask)

    def _forward_impl(self, query, key, value, need_weights, attn_mask):
        L, B, E = query.shape
        assert E == self.embed_dim
        assert key.shape == value.shape
        S = key.shape[0]
        H, dh = self.num_heads, self.head_dim
        dt = _compute_dtype(self)
        causal = is_causal_mask(attn_mask, L)
        scaling = float(dh) ** -0.5
        bias = self.in_proj_bias.detach() if self.in_proj_bias is not None else None

        def as2d(t):
            t2 = t.detach().reshape(-1, E)
            return t2 if t2.dtype == dt else ops.cast_2d(t2.contiguous().float(), dt)

        xq = as2d(query)
        if key is query and value is query:                                                 # rpr.py:250-253
            qkv = ops.linear(xq, self._cast.get(self.in_proj_weight, dt), bias, k=E, alpha=scaling, alpha_cols=E)
            q, k, v = qkv, qkv[:, E:], qkv[:, 2 * E:]
            ldq = ldk = qkv.stride(0)
        else:                                                                               # rpr.py:255-277
            assert key is value or torch.equal(key, value), "separate key / value tensors are not on the AMT path"
            q = ops.linear(xq, self._cast.get(self.in_proj_weight, dt, slice(0, E)), None if bias is None else bias[:E],
                           k=E, alpha=scaling, alpha_cols=E)
            kv = ops.linear(as2d(key), self._cast.get(self.in_proj_weight, dt, slice(E, 3 * E)),
                            None if bias is None else bias[E:], k=E)
            k, v = kv, kv[:, E:]
            ldq, ldk = q.stride(0), kv.stride(0)
        er = None
        if self.Er is not None:
            if L != S or L > self.Er.shape[0]:
                raise RuntimeError("RPR attention needs len_q == len_k <= er_len (got %d, %d, er_len %d); the reference "
                                   "fails in _skew for longer inputs (rpr.py:426-450)" % (L, S, self.Er.shape[0]))
            er = self.Er.detach() if dt == torch.float32 else ops.cast_2d(self.Er.detach(), dt)
        ctx = torch.empty((L * B, E), device=query.device, dtype=dt)
        p_out = torch.empty((B * H, L, S), device=query.device, dtype=torch.float32) if need_weights else None
        if p_out is not None and dt != torch.float32:
            raise NotImplementedError("need_weights=True is only available on the fp32 path")
        # rows are ordered (l, b): batch stride = row pitch, sequence stride = B * row pitch
        ops.attention(q, k, v, ctx, B=B, Hq=H, Hkv=H, Lq=L, Lk=S, dh=dh, q_strides=(ldq, B * ldq),
                      k_strides=(ldk, B * ldk), v_strides=(ldk, B * ldk), o_strides=(E, B * E), causal=causal, Er=er,
                      p_out=p_out)
        out = ops.linear(ctx, self._cast.get(self.out_proj.weight, dt),
                         self.out_proj.bias.detach() if self.out_proj.bias is not None else None, k=E,
                         out_dtype=torch.float32)
        out = out.view(L, B, E)
        if need_weights:                                                                    # rpr.py:419-422
            return out, p_out.view(B, H, L, S).sum(dim=1) / H
        return out, None


class TransformerDecoderLayerRPR(nn.Module):
    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, er_len=None):
        super().__init__()
        self.self_attn = MultiheadAttentionRPR(d_model, nhead, dropout=dropout, er_len=er_len)
        self.multihead_attn = MultiheadAttentionRPR(d_model, nhead, dropout=dropout, er_len=None)  # stock MHA arithmetic (rpr.py:42)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.dropout3 = nn.Dropout(dropout)
        self._cast = _Cast()

    def __deepcopy__(self, memo):
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = _Cast() if k == "_cast" else copy.deepcopy(v, memo)
        return new

    def _add_ln(self, x, y, norm):
        """norm(x + y) on (L,B,E) fp32 tensors."""
        return ops.layernorm(x.contiguous(), norm.weight.detach(), norm.bias.detach(), res=y.contiguous(), eps=norm.eps)

    def forward(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None,
                memory_key_padding_mask=None, **kwargs):
        if memory_mask is not None or tgt_key_padding_mask is not None or memory_key_padding_mask is not None:
            raise NotImplementedError("memory_mask / key padding masks are not used by the AMT path")
        dt = _compute_dtype(self)
        self.self_attn.compute_dtype = dt
        self.multihead_attn.compute_dtype = dt
        L, B, E = tgt.shape
        from . import autograd as ag
        if ag.tracking(tgt, memory, self) or ag.has_dropout(self):     # stand-alone training of the layer (fp32): autograd Functions
            if dt != torch.float32:
                raise NotImplementedError("module-level autograd runs on the fp32 path; bf16 training goes through "
                                          "VideoMusicTransformer (autograd.amt_forward_autograd)")
            add_ln = lambda x, y, n: ag.LayerNormFn.apply(ag.AddFn.apply(ag.rows_f32(x), ag.rows_f32(y), 1.0), n.weight, n.bias, n.eps)
            tr = self.training
            tgt2 = self.self_attn(tgt, tgt, tgt, attn_mask=tgt_mask, need_weights=False)[0]
            x = add_ln(tgt, ag.drop_any(tgt2, self.dropout1, tr), self.norm1)                        # rpr.py:59
            tgt2 = self.multihead_attn(x.view(L, B, E), memory, memory, need_weights=False)[0]
            x = add_ln(x, ag.drop_any(tgt2, self.dropout2, tr), self.norm2)                           # :65
            hdn = ag.drop_rows(ag.linear_fn(x, self.linear1, relu=True), self.dropout, tr)            # :67
            r = ag.AddFn.apply(x, ag.drop_rows(ag.linear_fn(hdn, self.linear2), self.dropout3, tr), 1.0)   # :68
            return ag.LayerNormFn.apply(r, self.norm3.weight, self.norm3.bias, self.norm3.eps).view(L, B, E)
        tgt2 = self.self_attn(tgt, tgt, tgt, attn_mask=tgt_mask, need_weights=False)[0]                 # rpr.py:56-57
        tgt = self._add_ln(tgt, tgt2, self.norm1)                                                        # :58-59
        tgt2 = self.multihead_attn(tgt, memory, memory, need_weights=False)[0]                           # :62-63
        tgt = self._add_ln(tgt, tgt2, self.norm2)                                                        # :65-66
        x2 = tgt.reshape(-1, E)
        xin = x2 if dt == torch.float32 else ops.cast_2d(x2, dt)
        hdn = ops.linear(xin, self._cast.get(self.linear1.weight, dt), self.linear1.bias.detach(), k=E, relu=True)
        r = ops.linear(hdn, self._cast.get(self.linear2.weight, dt), self.linear2.bias.detach(), k=hdn.shape[1],
                       residual=x2, out_dtype=torch.float32)                                             # :67-68
        return ops.layernorm(r, self.norm3.weight.detach(), self.norm3.bias.detach(), eps=self.norm3.eps).view(L, B, E)


class TransformerDecoderRPR(nn.Module):
    def __init__(self, decoder_layer, num_layers, norm=None):
        super().__init__()
        self.layers = _get_clones(decoder_layer, num_layers)
        self.num_layers = num_layers
        self.norm = norm

    def forward(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None,
                memory_key_padding_mask=None, **kwargs):
        output = tgt
        for mod in self.layers:
            mod.compute_dtype = _compute_dtype(self)
            output = mod(output, memory, tgt_mask=tgt_mask, memory_mask=memory_mask,
                         tgt_key_padding_mask=tgt_key_padding_mask, memory_key_padding_mask=memory_key_padding_mask)
        if self.norm is not None:
            from . import autograd as ag
            if ag.tracking(output, self.norm):
                return ag.LayerNormFn.apply(output.float().contiguous(), self.norm.weight, self.norm.bias, self.norm.eps)
            output = ops.layernorm(output.contiguous(), self.norm.weight.detach(), self.norm.bias.detach(), eps=self.norm.eps)
        return output
