"""Mamba blocks with the reference's module API (model/mamba.py, model/bimamba.py) on the B200 kernels.

Class names, constructor arguments, parameter names and shapes follow the reference, so `load_state_dict` of a reference
checkpoint works: MambaConfig (mamba.py:36-79), RMSNorm (:472-489), MambaBlock (:161-257), ResidualBlock (:137-159),
Mamba (:81-104), BiMambaEncoderLayer / BiMambaEncoder (bimamba.py:9-99).  The forward pass (mamba.py:259-351) runs as
GEMMs (in_proj, x_proj, dt_proj, out_proj) + a depthwise conv/SiLU kernel + ONE fused selective-scan kernel that never
materialises the (B, L, ED, N) tensors of `selective_scan` (:333-351).  fp32.  With grad mode on (input or parameters
requiring grad) the same kernels run inside autograd Functions (`autograd.MambaCoreFn`: fused-scan backward with recomputed
states, conv/SiLU backward; LinearFn / LayerNormFn / RMSNormFn around it), so Mamba, Mamba+ and the Bi-Mamba layers train;
dropout must be 0.  The stand-alone differentiable scan is video2music_b200.pscan (model/pscan.py:228).
"""
import math
from dataclasses import dataclass
from typing import Union

import copy

import torch
import torch.nn as nn

from . import ops


@dataclass
class MambaConfig:
    d_model: int
    n_layers: int
    dt_rank: Union[int, str] = 'auto'
    d_state: int = 16
    expand_factor: int = 2
    d_conv: int = 4
    dropout: int = 0.0
    use_KAN: bool = False
    use_version: int = 0            # 0: original mamba, 1: mamba+
    dt_min: float = 0.001
    dt_max: float = 0.1
    dt_init: str = "random"
    dt_scale: float = 1.0
    dt_init_floor = 1e-4
    rms_norm_eps: float = 1e-5
    base_std: float = 0.02
    bias: bool = False
    conv_bias: bool = True
    inner_layernorms: bool = False
    mup: bool = False
    mup_base_width: float = 128
    pscan: bool = True
    use_cuda: bool = False

    def __post_init__(self):
        self.d_inner = self.expand_factor * self.d_model
        if self.dt_rank == 'auto':
            self.dt_rank = math.ceil(self.d_model / 16)
        if self.mup:
            self.mup_width_mult = self.d_model / self.mup_base_width


def _ag():
    from . import autograd
    return autograd


class RMSNorm(nn.Module):
    def __init__(self, d_model: int, eps: float = 1e-5, use_mup: bool = False):
        super().__init__()
        self.use_mup = use_mup
        self.eps = eps
        if not use_mup:
            self.weight = nn.Parameter(torch.ones(d_model))

    def forward(self, x):
        if _ag().tracking(x, self):
            return _ag().RMSNormFn.apply(x.float().contiguous(), None if self.use_mup else self.weight, self.eps)
        return ops.rmsnorm(x.float(), None if self.use_mup else self.weight.detach(), self.eps)


class MambaBlock(nn.Module):
    def __init__(self, config: MambaConfig):
        super().__init__()
        if config.use_KAN:
            raise NotImplementedError("KANLinear experts (efficient_kan, unpinned third-party dependency) are out of scope")
        if config.inner_layernorms:
            raise NotImplementedError("inner_layernorms (jamba variant) is not used by the reference models")
        self.config = config
        self.dropout = nn.Dropout(config.dropout)
        self.in_proj = nn.Linear(config.d_model, 2 * config.d_inner, bias=config.bias)
        self.conv1d = nn.Conv1d(in_channels=config.d_inner, out_channels=config.d_inner, kernel_size=config.d_conv,
                                bias=config.conv_bias, groups=config.d_inner, padding=config.d_conv - 1)
        self.x_proj = nn.Linear(config.d_inner, config.dt_rank + 2 * config.d_state, bias=False)
        self.dt_proj = nn.Linear(config.dt_rank, config.d_inner, bias=True)
        dt_init_std = config.dt_rank ** -0.5 * config.dt_scale                     # mamba.py:188-195
        if config.dt_init == "constant":
            nn.init.constant_(self.dt_proj.weight, dt_init_std)
        elif config.dt_init == "random":
            nn.init.uniform_(self.dt_proj.weight, -dt_init_std, dt_init_std)
        else:
            raise NotImplementedError
        dt = torch.exp(torch.rand(config.d_inner) * (math.log(config.dt_max) - math.log(config.dt_min))
                       + math.log(config.dt_min)).clamp(min=config.dt_init_floor)
        inv_dt = dt + torch.log(-torch.expm1(-dt))                                  # inverse softplus, mamba.py:201-203
        with torch.no_grad():
            self.dt_proj.bias.copy_(inv_dt)
        A = torch.arange(1, config.d_state + 1, dtype=torch.float32).repeat(config.d_inner, 1)
        self.A_log = nn.Parameter(torch.log(A))
        self.A_log._no_weight_decay = True
        self.D = nn.Parameter(torch.ones(config.d_inner))
        self.D._no_weight_decay = True
        self.out_proj = nn.Linear(config.d_inner, config.d_model, bias=config.bias)

    def forward(self, x):
        """x (B, L, D) -> (B, L, D)   (mamba.py:259-291; dropout is defined but never applied by the reference forward)."""
        cfg = self.config
        B, L, D = x.shape
        ED, N, R = cfg.d_inner, cfg.d_state, cfg.dt_rank
        ag = _ag()
        if ag.tracking(x, self):                                   # training: same kernels inside autograd Functions
            # compute_dtype = bf16: the two wide projections (in_proj, out_proj: ~85 % of the block's GEMM flops) run on the tcgen05
            # GEMM with bf16 operands; conv, the small x_proj / dt_proj, the scan, master weights and gradients stay fp32
            tc = getattr(self, "compute_dtype", torch.float32) == torch.bfloat16 and ag.bf16_ok(self.in_proj, self.out_proj)
            if tc:
                xz = ag.linear_bf16_fn(ag.rows_f32(x).to(torch.bfloat16), self.in_proj, out_dtype=torch.float32)
            else:
                xz = ag.linear_fn(ag.rows_f32(x), self.in_proj)
            y = ag.MambaCoreFn.apply(xz, self.conv1d.weight.reshape(ED, -1), self.conv1d.bias, self.x_proj.weight, self.dt_proj.weight,
                                     self.dt_proj.bias, self.A_log, self.D, B, L, cfg.use_version == 1)
            if tc:
                return ag.linear_bf16_fn(y.to(torch.bfloat16), self.out_proj, out_dtype=torch.float32).view(B, L, D)
            return ag.linear_fn(y, self.out_proj).view(B, L, D)
        x2 = x.reshape(B * L, D).float().contiguous()
        det = lambda p: None if p is None else p.detach()
        xz = ops.linear(x2, det(self.in_proj.weight), det(self.in_proj.bias))                       # (B*L, 2ED): x | z
        xc = ops.mamba_conv_silu(xz, ED, det(self.conv1d.weight).reshape(ED, -1), det(self.conv1d.bias), B, L)   # (B*L, ED)
        dbc = ops.linear(xc, det(self.x_proj.weight))                                               # (B*L, R + 2N)
        draw = ops.linear(dbc[:, :R], det(self.dt_proj.weight))                                     # delta before bias / softplus
        y = ops.selective_scan(xc, draw, det(self.dt_proj.bias), det(self.A_log), dbc[:, R:R + N], dbc[:, R + N:], det(self.D),
                               xz[:, ED:], B, L, plus=(cfg.use_version == 1))
        return ops.linear(y, det(self.out_proj.weight), det(self.out_proj.bias)).view(B, L, D)


    @torch.no_grad()
    def step(self, x, cache):
        """Recurrent single-token inference (mamba.py:407-438): x (B, D), cache = (h (B, ED, N) or None, inputs (B, ED,
        d_conv - 1)) -> (output (B, D), new cache).  Literal: the gate is y * silu(z) for both use_version values (:430), the
        caller's cache tensors are left untouched.  Five launches: in_proj GEMM, conv window, x_proj GEMM, fused dt_proj +
        softplus + state update + C contraction + gate, out_proj GEMM."""
        cfg = self.config
        h, inputs = cache
        det = lambda p: None if p is None else p.detach()
        xz = ops.linear(x.float().contiguous(), det(self.in_proj.weight), det(self.in_proj.bias))          # (B, 2 ED)
        out, h_new, in_new = ops.mamba_step(xz, cfg.d_inner, det(self.conv1d.weight).reshape(cfg.d_inner, -1), det(self.conv1d.bias),
                                            det(self.x_proj.weight), det(self.dt_proj.weight), det(self.dt_proj.bias),
                                            det(self.A_log), det(self.D), h, inputs)
        return ops.linear(out, det(self.out_proj.weight), det(self.out_proj.bias)), (h_new, in_new)


class ResidualBlock(nn.Module):
    def __init__(self, config: MambaConfig):
        super().__init__()
        self.mixer = MambaBlock(config)
        self.norm = RMSNorm(config.d_model, config.rms_norm_eps, config.mup)

    def forward(self, x):
        return self.mixer(self.norm(x)) + x                                                         # mamba.py:144-149

    def step(self, x, cache):
        """mamba.py:151-159: x (B, D), cache (h, inputs) -> (mixer.step(norm(x)) + x, new cache)."""
        output, cache = self.mixer.step(self.norm(x), cache)
        return output + x, cache


class Mamba(nn.Module):
    def __init__(self, config: MambaConfig):
        super().__init__()
        self.config = config
        self.layers = nn.ModuleList([ResidualBlock(config) for _ in range(config.n_layers)])

    def step(self, x, caches):
        """mamba.py:100-108: one token through every layer; caches[i] = (h, inputs) of layer i (a new list is returned)."""
        caches = list(caches)
        for i, layer in enumerate(self.layers):
            x, caches[i] = layer.step(x, caches[i])
        return x, caches

    def forward(self, x):
        for layer in self.layers:
            x = layer(x)
        return x


class _FFN(nn.Sequential):
    """nn.Sequential(Linear, ReLU, Dropout, Linear) with the reference's child indices (bimamba.py:50-62)."""

    def __init__(self, d_model, d_ff, dropout):
        super().__init__(nn.Linear(d_model, d_ff), nn.ReLU(), nn.Dropout(dropout), nn.Linear(d_ff, d_model))

    def forward(self, x):
        shp = x.shape
        ag = _ag()
        drop = self.training and self[2].p > 0
        if ag.tracking(x, self) or drop:
            if getattr(self, "compute_dtype", torch.float32) == torch.bfloat16 and ag.bf16_ok(self[0], self[3]):   # tensor-core path
                h = ag.linear_bf16_fn(ag.rows_f32(x).to(torch.bfloat16), self[0], relu=True, out_dtype=torch.float32 if drop else torch.bfloat16)
                if drop:                                       # the element-wise dropout op is fp32
                    h = ag.drop_rows(h, self[2], self.training).to(torch.bfloat16)
                return ag.linear_bf16_fn(h, self[3], out_dtype=torch.float32).view(shp)
            return ag.linear_fn(ag.drop_rows(ag.linear_fn(ag.rows_f32(x), self[0], relu=True), self[2], self.training), self[3]).view(shp)
        x2 = x.reshape(-1, shp[-1]).float().contiguous()
        h = ops.linear(x2, self[0].weight.detach(), self[0].bias.detach(), relu=True)
        return ops.linear(h, self[3].weight.detach(), self[3].bias.detach()).view(shp)


def _sum(a, b):
    """a + b (one kernel; inside an autograd Function when gradients are tracked)."""
    ag = _ag()
    if ag.tracking(a, b):
        return ag.AddFn.apply(a.float().contiguous(), b.float().contiguous(), 1.0)
    return ops.axpy(a.float().contiguous(), b.float().contiguous(), 1.0)


def _add_norm(norm: nn.LayerNorm, a, b):
    shp = a.shape
    ag = _ag()
    if ag.tracking(a, b, norm):
        return ag.LayerNormFn.apply(ag.AddFn.apply(ag.rows_f32(a), ag.rows_f32(b), 1.0), norm.weight, norm.bias, norm.eps).view(shp)
    return ops.layernorm(a.reshape(-1, shp[-1]).contiguous(), norm.weight.detach(), norm.bias.detach(),
                         res=b.reshape(-1, shp[-1]).contiguous(), eps=norm.eps).view(shp)


class BiMambaEncoderLayer(nn.Module):
    """Bi-Mamba4TS layer, bimamba.py:34-99 (literal: the backward branch's FFN reads the FORWARD branch, :91)."""

    def __init__(self, config: MambaConfig, dim_feedforward=1024, dropout=0.2):
        super().__init__()
        self.config = config
        self.mamba_forward = MambaBlock(config)
        self.mamba_backward = MambaBlock(config)
        self.d_ff = dim_feedforward
        self.norm1 = nn.LayerNorm(config.d_model)
        self.norm2 = nn.LayerNorm(config.d_model)
        self.norm3 = nn.LayerNorm(config.d_model)
        self.norm4 = nn.LayerNorm(config.d_model)
        self.dropout = nn.Dropout(dropout)
        self.ffn1 = _FFN(config.d_model, dim_feedforward, dropout)
        self.ffn2 = _FFN(config.d_model, dim_feedforward, dropout)

    def forward(self, x):
        x = x.float()
        dr = lambda t: _ag().drop_any(t, self.dropout, self.training)            # self.dropout(.) of bimamba.py:73,79,87,93
        x_flip = torch.flip(x, dims=[1])
        x_f = _add_norm(self.norm1, dr(self.mamba_forward(x)), x)
        x_f = _add_norm(self.norm2, dr(self.ffn1(x_f)), x_f)
        x_b = torch.flip(self.mamba_backward(x_flip), dims=[1])
        x_b = _add_norm(self.norm3, dr(x_b), x)
        x_b = _add_norm(self.norm4, dr(self.ffn2(x_f)), x_b)
        return _sum(x_f, x_b)


def _norm_only(norm: nn.LayerNorm, a):
    shp = a.shape
    ag = _ag()
    if ag.tracking(a, norm):
        return ag.LayerNormFn.apply(ag.rows_f32(a), norm.weight, norm.bias, norm.eps).view(shp)
    return ops.layernorm(a.reshape(-1, shp[-1]).contiguous(), norm.weight.detach(), norm.bias.detach(), eps=norm.eps).view(shp)


class BiMambaEncoderLayer_V1(nn.Module):
    """Bi-Mamba+ layer, bimamba.py:101-191: two mamba+ blocks (forward / flipped sequence), one feed-forward that is either
    Linear-ReLU-Linear or a deep copy of the caller's MoE layer, post-norm or norm_first."""

    def __init__(self, config: MambaConfig, dim_feedforward=1024, dropout=0.2, moe_layer=None, norm_first=False):
        super().__init__()
        assert config.use_version == 1, "use_version should be 1 to use Mamba+"
        self.config = config
        self.mamba_forward = MambaBlock(config)
        self.mamba_backward = MambaBlock(config)
        self.d_ff = dim_feedforward
        self.dropout = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(config.d_model)
        self.norm2 = nn.LayerNorm(config.d_model)
        self.norm3 = nn.LayerNorm(config.d_model)
        self.norm_first = norm_first
        self.ffn = _FFN(config.d_model, dim_feedforward, dropout) if moe_layer is None else copy.deepcopy(moe_layer)

    def forward(self, x):
        x = x.float().contiguous()
        dr = lambda t: _ag().drop_any(t, self.dropout, self.training)            # self.dropout(.) of bimamba.py:142,151,162,170,178,188
        x_flip = torch.flip(x, dims=[1])
        if self.norm_first:                                                      # bimamba.py:141-165
            x_f = _sum(x, dr(self.mamba_forward(_norm_only(self.norm1, x))))
            x_b = dr(torch.flip(self.mamba_backward(_norm_only(self.norm2, x_flip)), dims=[1]))
            x_b = _sum(x, x_b)
            x = _sum(x_f, x_b)
            return _sum(x, dr(self.ffn(_norm_only(self.norm3, x)).float()))
        x_f = _add_norm(self.norm1, dr(self.mamba_forward(x)), x)                # :168-189
        x_b = _add_norm(self.norm2, dr(torch.flip(self.mamba_backward(x_flip), dims=[1])), x)
        x = _sum(x_f, x_b)
        return _add_norm(self.norm3, dr(self.ffn(x).float()), x)


class BiMambaEncoder(nn.Module):
    def __init__(self, config: MambaConfig, dim_feedforward=1024, n_encoder_layers=2, dropout=0.2, moe_layer=None, norm_first=False):
        super().__init__()
        self.n_encoder_layers = n_encoder_layers
        if config.use_version == 0:                                              # bimamba.py:15-19
            self.layers = nn.ModuleList([BiMambaEncoderLayer(config, dim_feedforward, dropout) for _ in range(n_encoder_layers)])
        else:
            self.layers = nn.ModuleList([BiMambaEncoderLayer_V1(config, dim_feedforward, dropout, moe_layer=moe_layer,
                                                                norm_first=norm_first) for _ in range(n_encoder_layers)])
        self.norm_first = norm_first
        if norm_first:
            self.norm = nn.LayerNorm(config.d_model)

    def forward(self, x):
        for i in range(self.n_encoder_layers):
            x = self.layers[i](x)
        if self.norm_first:
            x = _norm_only(self.norm, x)
        return x
