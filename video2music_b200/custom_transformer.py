"""Drop-in for the generic layer wrappers of model/custom_transformer.py that the V1-V3 model zoo and
BASELINE config 4 (grouped-query attention + MoE FFN) are assembled from:

    RMSNorm                                  (custom_transformer.py:27-47)
    TransformerEncoderLayer / DecoderLayer   (:1220-1292)   post- or pre-norm, NO dropout on the residuals
    TransformerEncoder / Decoder             (:1371-1401)   n deep copies + optional final norm
    TransformerEncoderShorter / DecoderShorter (:1403-1433) caller-provided layer list

Same constructor signatures and attribute names (`self_attn`, `cross_attn`, `ff`, `norm1..3`, `layers`,
`norm`), so a reference checkpoint loads.  The attention and feed-forward sub-modules are whatever the
caller plugs in (MultiheadGQA, MultiheadAttentionRPR, MoELayer, SharedMoELayer, GLUExpert ...); the
residual add and the normalisation run as one kernel of ours.  Inference only (inputs are detached);
training goes through `autograd.py` for the base AMT.  The RoSC layers (:1294-1369) are out of scope.
"""
from copy import deepcopy

import torch
import torch.nn as nn

from . import ops
from .rpr import _get_clones


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
        w = self.weight.detach() if self.weight is not None else None
        return ops.rmsnorm(x.detach().float(), w, self.eps).type_as(x)

    def extra_repr(self) -> str:
        return f"dim={self.dim}, eps={self.eps}, elementwise_affine={self.elementwise_affine}"


def _norm(norm, x, res=None):
    """norm(x + res): one fused kernel for LayerNorm, add + RMSNorm kernels otherwise."""
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
