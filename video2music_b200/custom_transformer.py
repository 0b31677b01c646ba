"""Drop-in for the generic layer wrappers of model/custom_transformer.py that the V1-V3 model zoo and
BASELINE config 4 (grouped-query attention + MoE FFN) are assembled from:

    RMSNorm                                  (custom_transformer.py:27-47)
    TransformerEncoderLayer / DecoderLayer   (:1220-1292)   post- or pre-norm, NO dropout on the residuals
    TransformerEncoder / Decoder             (:1371-1401)   n deep copies + optional final norm
    TransformerEncoderShorter / DecoderShorter (:1403-1433) caller-provided layer list

Same constructor signatures and attribute names (`self_attn`, `cross_attn`, `ff`, `norm1..3`, `layers`,
`norm`), so a reference checkpoint loads.  The attention and feed-forward sub-modules are whatever the
caller plugs in (MultiheadGQA, MultiheadAttentionRPR, MoELayer, SharedMoELayer, GLUExpert ...); the
residual add and the normalisation run as one kernel of ours.  The generic wrappers, RMSNorm, MultiheadGQA and the MoE
layers also train (fp32): with grad mode on they chain the autograd Functions of `autograd.py`, whose forward and backward
are our kernels (BASELINE config 4).  CustomMultiheadAttention / DifferentialMultiheadAttention: inference only.  The RoSC layers (:1294-1369) are out of scope.
"""
import math
from copy import deepcopy

import torch
import torch.nn as nn

from . import ops
from .rpr import _get_clones


class RotaryPositionalEmbeddings(nn.Module):
    """Rotation cache of rotate_operation.py:53-113 (theta_i = base^(-2i/dim), cache[pos][i] = (cos, sin)(pos * theta_i));
    buffers are non-persistent as in the reference.  Applying it is `ops.rope_quirk` (see CustomMultiheadAttention)."""

    def __init__(self, dim: int, max_seq_len: int = 4096, base: int = 10_000) -> None:
        super().__init__()
        self.dim, self.base, self.max_seq_len = dim, base, max_seq_len
        theta = 1.0 / (base ** (torch.arange(0, dim, 2)[: (dim // 2)].float() / dim))
        self.register_buffer("theta", theta, persistent=False)
        idx_theta = torch.einsum("i, j -> ij", torch.arange(max_seq_len, dtype=theta.dtype), theta).float()
        self.register_buffer("cache", torch.stack([torch.cos(idx_theta), torch.sin(idx_theta)], dim=-1), persistent=False)


class CustomMultiheadAttention(nn.Module):
    """Drop-in for custom_transformer.py:51-321 (nn.MultiheadAttention + optional RoPE, the attention of the V2 / V3 model
    zoo): packed in-projection, RoPE on q and k with the reference's literal reinterpretations (:1044-1053), scaled
    dot-product attention (causal float mask or none), out-projection.  fp32; with gradients being tracked the same kernels run
    inside autograd Functions (autograd.custom_mha_autograd: RoPE backward = the rotation with negated sines).
    Same parameter names (`in_proj_weight`, `in_proj_bias`, `out_proj.weight`, `out_proj.bias`)."""

    def __init__(self, embed_dim, num_heads, dropout=0., bias=True, add_bias_kv=False, add_zero_attn=False, kdim=None, vdim=None,
                 batch_first=False, device=None, dtype=None, RoPE=None) -> None:
        super().__init__()
        if embed_dim <= 0 or num_heads <= 0:
            raise ValueError(f"embed_dim and num_heads must be greater than 0, got embed_dim={embed_dim} and num_heads={num_heads} instead")
        if add_bias_kv or add_zero_attn or batch_first or (kdim not in (None, embed_dim)) or (vdim not in (None, embed_dim)) or not bias:
            raise NotImplementedError("add_bias_kv / add_zero_attn / batch_first / kdim / vdim / bias=False are never used by the reference's models")
        self.embed_dim, self.num_heads, self.dropout, self.batch_first = embed_dim, num_heads, dropout, False
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == embed_dim, "embed_dim must be divisible by num_heads"
        self.RoPE = deepcopy(RoPE)
        self.in_proj_weight = nn.Parameter(torch.empty((3 * embed_dim, embed_dim)))
        self.in_proj_bias = nn.Parameter(torch.empty(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=True)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.in_proj_bias, 0.)
        nn.init.constant_(self.out_proj.bias, 0.)

    def forward(self, query, key, value, key_padding_mask=None, need_weights=True, attn_mask=None, average_attn_weights=True,
                is_causal=False):
        from .rpr import is_causal_mask
        if key_padding_mask is not None:
            raise NotImplementedError("key_padding_mask is not used by the reference's models")
        from . import autograd as ag
        if ag.tracking(query, key, value, self):                                        # training: the same kernels inside autograd Functions
            return ag.custom_mha_autograd(self, query, key, value, need_weights, attn_mask, average_attn_weights)
        drop = (float(self.dropout), ops.next_dropout_seed()) if (self.training and self.dropout > 0) else None
        L, B, E = query.shape
        S = key.shape[0]
        H, dh = self.num_heads, self.head_dim
        causal = is_causal_mask(attn_mask, L) if attn_mask is not None else False
        w, b = self.in_proj_weight.detach(), self.in_proj_bias.detach()
        flat = lambda t: t.detach().reshape(-1, E).float().contiguous()
        if key is query and value is query:                                            # packed self-attention projection
            qkv = ops.linear(flat(query), w, b)
            q, k, v = qkv[:, :E].contiguous(), qkv[:, E:2 * E].contiguous(), qkv[:, 2 * E:]
        else:
            q = ops.linear(flat(query), w[:E], b[:E])
            kv = ops.linear(flat(key), w[E:], b[E:]) if key is value else None
            k = kv[:, :E].contiguous() if kv is not None else ops.linear(flat(key), w[E:2 * E], b[E:2 * E])
            v = kv[:, E:] if kv is not None else ops.linear(flat(value), w[2 * E:], b[2 * E:])
        if self.RoPE is not None:                                                       # custom_transformer.py:1047-1050
            cache = self.RoPE.cache
            q = ops.rope_quirk(q, cache[:L].contiguous(), B, H)
            k = ops.rope_quirk(k, cache[:S].contiguous(), B, H)
        out = torch.empty((L * B, E), device=q.device, dtype=torch.float32)             # rows (l, b): sequence-first strides
        p_out = torch.empty((B * H, L, S), device=q.device, dtype=torch.float32) if need_weights else None
        ops.attention(q, k, v, out, B=B, Hq=H, Hkv=H, Lq=L, Lk=S, dh=dh, q_strides=(q.stride(0), B * q.stride(0)),
                      k_strides=(k.stride(0), B * k.stride(0)), v_strides=(v.stride(0), B * v.stride(0)), o_strides=(E, B * E),
                      causal=causal, q_scale=float(dh) ** -0.5, p_out=p_out, dropout=drop)   # F.multi_head_attention_forward's dropout_p
        y = ops.linear(out, self.out_proj.weight.detach(), self.out_proj.bias.detach()).view(L, B, E)
        if not need_weights:
            return y, None
        wts = p_out.view(B, H, L, S)
        return y, (wts.mean(dim=1) if average_attn_weights else wts)


def lambda_init_fn(depth):
    return 0.8 - 0.6 * math.exp(-0.3 * depth)                 # custom_transformer.py:607-608


class DifferentialMultiheadAttention(nn.Module):
    """Drop-in for custom_transformer.py:610-832 (Differential Transformer attention of the V3 models), inference and fp32
    training (autograd.diff_mha_autograd).
    q, k are projected to 2 * num_heads heads, v to num_heads heads; head pair (2h, 2h+1) gives
        softmax(q_2h k_2h^T) v_h - lambda * softmax(q_2h+1 k_2h+1^T) v_h,
    followed by a per-head RMSNorm (`subln`), the factor (1 - lambda_init) and the bias-free out-projection.
    Literal layout behaviour kept: RoPE on the [2H, len, B, dh] VIEW of the projections (:777-784), then the same memory
    viewed batch-first as [B, len, 2H, dh] (:786-788), and the (B, H, L, dh) result viewed as (L, B, E) (:824)."""

    def __init__(self, embed_dim, num_heads, dropout=0., batch_first=False, device=None, dtype=None, RoPE=None, depth=2) -> None:
        super().__init__()
        if embed_dim <= 0 or num_heads <= 0:
            raise ValueError(f"embed_dim and num_heads must be greater than 0, got embed_dim={embed_dim} and num_heads={num_heads} instead")
        if batch_first:
            raise NotImplementedError("batch_first is never used by the reference's models")
        self.embed_dim, self.num_heads, self.batch_first = embed_dim, num_heads, False
        self.dropout = nn.Dropout(dropout)
        self.head_dim = embed_dim // num_heads
        self.scaling = self.head_dim ** -0.5
        self.RoPE = deepcopy(RoPE)
        self.k_proj = nn.Linear(embed_dim, embed_dim * 2, bias=False)
        self.q_proj = nn.Linear(embed_dim, embed_dim * 2, bias=False)
        self.v_proj = nn.Linear(embed_dim, embed_dim, bias=False)
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=False)
        for lin in (self.k_proj, self.q_proj, self.v_proj, self.out_proj):
            nn.init.xavier_uniform_(lin.weight)
        self.lambda_init = lambda_init_fn(depth)
        for n in ("lambda_q1", "lambda_k1", "lambda_q2", "lambda_k2"):
            setattr(self, n, nn.Parameter(torch.zeros(self.head_dim, dtype=torch.float32).normal_(mean=0, std=0.1)))
        self.subln = RMSNorm(self.head_dim, eps=1e-5, elementwise_affine=True)
        self._lam = None

    def _lambda_full(self) -> float:
        ps = (self.lambda_q1, self.lambda_k1, self.lambda_q2, self.lambda_k2)
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if self._lam is None or self._lam[0] != key:                     # four 64-vectors: evaluated on the host, cached
            q1, k1, q2, k2 = (p.detach().float().cpu() for p in ps)
            lam = float(torch.exp(torch.sum(q1 * k1)) - torch.exp(torch.sum(q2 * k2))) + self.lambda_init
            self._lam = (key, lam)
        return self._lam[1]

    def forward(self, query, key, value, key_padding_mask=None, need_weights=True, attn_mask=None, average_attn_weights=True,
                is_causal=False):
        if key_padding_mask is not None:
            raise NotImplementedError("key_padding_mask is not used by the reference's models")
        from . import autograd as ag
        if ag.tracking(query, key, value, self):                          # training: the same kernels inside autograd Functions
            return ag.diff_mha_autograd(self, query, key, value, attn_mask)
        L, B, E = query.shape
        S = key.shape[0]
        H, dh = self.num_heads, self.head_dim
        flat = lambda t: t.detach().reshape(-1, E).float().contiguous()
        q = ops.linear(flat(query), self.q_proj.weight.detach())          # (L*B, 2E), rows (l, b)
        k = ops.linear(flat(key), self.k_proj.weight.detach())
        v = ops.linear(flat(value), self.v_proj.weight.detach())          # (S*B, E)
        if self.RoPE is not None:
            q = ops.rope_quirk(q, self.RoPE.cache[:L].contiguous(), B, 2 * H)
            k = ops.rope_quirk(k, self.RoPE.cache[:S].contiguous(), B, 2 * H)
        # the same memory, now read batch-first as [B, len, 2H, dh]; even / odd heads are split into two operands
        q5, k5 = q.view(B, L, H, 2, dh), k.view(B, S, H, 2, dh)
        causal = attn_mask is not None                                    # any mask means the causal one here (:801-809)
        outs = []
        for i in (0, 1):
            qi, ki = q5[:, :, :, i].contiguous(), k5[:, :, :, i].contiguous()         # (B, len, H, dh)
            o = torch.empty((B, L, H, dh), device=q.device, dtype=torch.float32)
            ops.attention(qi, ki, v, o, B=B, Hq=H, Hkv=H, Lq=L, Lk=S, dh=dh, q_strides=(L * E, E), k_strides=(S * E, E),
                          v_strides=(S * E, E), o_strides=(L * E, E), causal=causal, q_scale=self.scaling)
            outs.append(o)
        diff = ops.axpy(outs[0], outs[1], -self._lambda_full())            # a1 v - lambda a2 v
        w = (self.subln.weight.detach() * (1.0 - self.lambda_init)).contiguous()
        attn = ops.rmsnorm(diff.view(-1, dh), w, self.subln.eps).view(B, L, H, dh)
        attn = attn.permute(0, 2, 1, 3).contiguous().view(L * B, E)        # (B, H, L, dh) memory read as (L, B, E) rows (:824)
        return ops.linear(attn, self.out_proj.weight.detach()).view(L, B, E), None


class RMSNorm(nn.Module):
    """x * rsqrt(mean(x^2) + eps) * weight, eps 1e-6 by default (custom_transformer.py:27-47)."""

    def __init__(self, dim: int, eps: float = 1e-6, elementwise_affine=True, memory_efficient=False):
        super().__init__()
        self.dim = dim
        self.eps = eps
        self.elementwise_affine = elementwise_affine
        if self.elementwise_affine:
            self.weight = nn.Parameter(torch.ones(dim))
        else:
            self.register_parameter("weight", None)

    def forward(self, x):
        from . import autograd as ag
        if ag.tracking(x, self):
            return ag.RMSNormFn.apply(x.float().contiguous(), self.weight, self.eps).type_as(x)
        w = self.weight.detach() if self.weight is not None else None
        return ops.rmsnorm(x.detach().float(), w, self.eps).type_as(x)

    def extra_repr(self) -> str:
        return f"dim={self.dim}, eps={self.eps}, elementwise_affine={self.elementwise_affine}"


def _norm(norm, x, res=None):
    """norm(x + res): one fused kernel for LayerNorm, add + RMSNorm kernels otherwise.  With gradients being tracked the
    same kernels run inside autograd Functions (add, then the norm with its backward kernel)."""
    from . import autograd as ag
    if ag.tracking(x, res, norm):
        x = x.float().contiguous()
        if res is not None:
            x = ag.AddFn.apply(x, res.float().contiguous(), 1.0)
        if isinstance(norm, nn.LayerNorm):
            return ag.LayerNormFn.apply(x, norm.weight, norm.bias, norm.eps)
        if isinstance(norm, RMSNorm):
            return ag.RMSNormFn.apply(x, norm.weight, norm.eps)
        raise NotImplementedError("norm layer %s (LayerNorm and RMSNorm are built)" % type(norm).__name__)
    x = x.detach().float().contiguous()
    if isinstance(norm, nn.LayerNorm):
        return ops.layernorm(x, norm.weight.detach(), norm.bias.detach(), res=None if res is None else res.detach().float().contiguous(),
                             eps=norm.eps)
    if res is not None:
        x = ops.axpy(x, res.detach().float().contiguous(), 1.0)
    if isinstance(norm, RMSNorm) or hasattr(norm, "eps") and hasattr(norm, "weight") and not hasattr(norm, "bias"):
        w = norm.weight.detach() if norm.weight is not None else None
        return ops.rmsnorm(x, w, norm.eps)
    raise NotImplementedError("norm layer %s (LayerNorm and RMSNorm are built)" % type(norm).__name__)


def _add(x, y):
    from . import autograd as ag
    if ag.tracking(x, y):
        return ag.AddFn.apply(x.float().contiguous(), y.float().contiguous(), 1.0)
    return ops.axpy(x.detach().float().contiguous(), y.detach().float().contiguous(), 1.0)


class TransformerEncoderLayer(nn.Module):
    def __init__(self, self_att_layer, ff_layer, pre_norm=False, norm=None, dropout=0.1):
        super().__init__()
        self.self_attn = deepcopy(self_att_layer)
        self.ff = deepcopy(ff_layer)
        self.pre_norm = pre_norm
        self.norm1, self.norm2 = _get_clones(norm, 2)

    def forward(self, src, src_mask=None, src_key_padding_mask=None, **kwargs):
        if not self.pre_norm:                                                    # custom_transformer.py:1231-1238
            src2 = self.self_attn(src, src, src, attn_mask=src_mask, key_padding_mask=src_key_padding_mask)[0]
            src = _norm(self.norm1, src, src2)
            return _norm(self.norm2, src, self.ff(src))
        src2 = _norm(self.norm1, src)                                            # :1239-1247
        src = _add(src, self.self_attn(src2, src2, src2, attn_mask=src_mask, key_padding_mask=src_key_padding_mask)[0])
        return _add(src, self.ff(_norm(self.norm2, src)))


class TransformerDecoderLayer(nn.Module):
    def __init__(self, self_att_layer, cross_att_layer, ff_layer, pre_norm=False, norm=None, dropout=0.1):
        super().__init__()
        self.self_attn = deepcopy(self_att_layer)
        self.cross_attn = deepcopy(cross_att_layer)
        self.ff = deepcopy(ff_layer)
        self.pre_norm = pre_norm
        self.norm1, self.norm2, self.norm3 = _get_clones(norm, 3)

    def forward(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None, memory_key_padding_mask=None):
        if not self.pre_norm:                                                    # custom_transformer.py:1263-1277
            tgt2 = self.self_attn(tgt, tgt, tgt, attn_mask=tgt_mask, key_padding_mask=tgt_key_padding_mask)[0]
            tgt = _norm(self.norm1, tgt, tgt2)
            tgt2 = self.cross_attn(tgt, memory, memory, attn_mask=memory_mask, key_padding_mask=memory_key_padding_mask)[0]
            tgt = _norm(self.norm2, tgt, tgt2)
            return _norm(self.norm3, tgt, self.ff(tgt))
        tgt2 = _norm(self.norm1, tgt)                                            # :1278-1291
        tgt = _add(tgt, self.self_attn(tgt2, tgt2, tgt2, attn_mask=tgt_mask, key_padding_mask=tgt_key_padding_mask)[0])
        tgt2 = _norm(self.norm2, tgt)
        tgt = _add(tgt, self.cross_attn(tgt2, memory, memory, attn_mask=memory_mask, key_padding_mask=memory_key_padding_mask)[0])
        return _add(tgt, self.ff(_norm(self.norm3, tgt)))


class TransformerEncoder(nn.Module):
    def __init__(self, encoder_layer, num_layers, norm=None):
        super().__init__()
        self.layers = _get_clones(encoder_layer, num_layers)
        self.num_layers = num_layers
        self.norm = deepcopy(norm)

    def forward(self, src, mask=None, src_key_padding_mask=None, **kwargs):
        output = src
        for mod in self.layers:
            output = mod(output, src_mask=mask, src_key_padding_mask=src_key_padding_mask)
        if self.norm:                                                            # :1383 (truthiness, as the reference)
            output = _norm(self.norm, output)
        return output


class TransformerDecoder(nn.Module):
    def __init__(self, decoder_layer, num_layers, norm=None):
        super().__init__()
        self.layers = _get_clones(decoder_layer, num_layers)
        self.num_layers = num_layers
        self.norm = deepcopy(norm)

    def forward(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None, memory_key_padding_mask=None,
                **kwargs):
        output = tgt
        for mod in self.layers:
            output = mod(output, memory, tgt_mask=tgt_mask, memory_mask=memory_mask, tgt_key_padding_mask=tgt_key_padding_mask,
                         memory_key_padding_mask=memory_key_padding_mask)
        if self.norm is not None:
            output = _norm(self.norm, output)
        return output


class TransformerEncoderShorter(nn.Module):
    def __init__(self, encoder_layers, norm=None):
        super().__init__()
        self.layers = encoder_layers
        self.norm = deepcopy(norm)

    forward = TransformerEncoder.forward


class TransformerDecoderShorter(nn.Module):
    def __init__(self, decoder_layers, norm=None):
        super().__init__()
        self.layers = decoder_layers
        self.norm = deepcopy(norm)

    forward = TransformerDecoder.forward
