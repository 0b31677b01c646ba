"""Drop-in for the MoE FFN of model/moe.py: GLUExpert (:36-49), MoELayer (:150-200), SharedMoELayer
(:202-302), TopKScheduler (:66-82), TemperatureScheduler (:84-97) -- same constructor signatures and
parameter names (`experts.N.{linear1,linear2,gate}`, `gate`, `shared_expert`, buffer `bias`).

Routing (gate GEMV + top-k + fp32 softmax + expert histogram) is one fused kernel; token copies are then
permuted into expert-contiguous order by a second kernel and all experts run as two ragged grouped GEMMs
(group bounds read on the device, SwiGLU fused into the first) followed by a weighted combine (see
`_experts_forward`): five launches per MoE layer and no host synchronisation.

Training: when grad mode is on and the input or a parameter requires grad, the layers run through
`autograd.MoEExpertsFn` (fp32): permute, grouped GEMMs, SwiGLU, combine and the softmax-over-top-k router gradient are all
our kernels in both directions (`csrc/moe.cu`: combine_bwd, swiglu_bwd, grouped_dw); dropout must be 0.

`module.compute_dtype = torch.bfloat16` moves the expert GEMMs to the tcgen05 tensor-core path (grouped GEMM over
128-row aligned expert groups, bf16 operands, fp32 accumulation); the router always runs in fp32, so the routing indices
stay bit-exact.  The reference's per-expert Python loop with
`torch.where` host syncs (moe.py:192-199) is gone; the logging side channels
(third_party/log_experts.py, log_maxvio.py) are kept as optional callables (`on_route`).
"""
import copy
from typing import Callable, Optional

import torch
import torch.nn as nn

from . import ops
from .rpr import _get_clones


class GLUExpert(nn.Module):
    def __init__(self, d_model, d_ff=2048, dropout=0.1):
        super().__init__()
        self.linear1 = nn.Linear(d_model, d_ff)
        self.linear2 = nn.Linear(d_ff, d_model)
        self.gate = nn.Linear(d_model, d_ff)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x):
        """x (..., d) -> linear2(dropout((linear1 x) * silu(gate x)))  (moe.py:44-49)."""
        shp = x.shape
        from . import autograd as ag
        drop = self.training and self.dropout.p > 0
        if drop and getattr(self, "compute_dtype", torch.float32) != torch.float32:
            raise NotImplementedError("training-mode dropout of GLUExpert is built on the fp32 path")
        if ag.tracking(x, self) or drop:                                  # training: forward and backward on our kernels
            return ag.glu_expert_fn(self, ag.rows_f32(x)).view(shp[:-1] + (self.linear2.out_features,))
        x2 = x.detach().reshape(-1, shp[-1]).float().contiguous()
        return _glu(self, x2).view(shp[:-1] + (self.linear2.out_features,))


def _glu(e: GLUExpert, x2: torch.Tensor) -> torch.Tensor:
    d = x2.shape[1]
    a = ops.linear(x2, e.linear1.weight.detach(), e.linear1.bias.detach(), k=d)
    g = ops.linear(x2, e.gate.weight.detach(), e.gate.bias.detach(), k=d)
    h = ops.swiglu(a, g)
    return ops.linear(h, e.linear2.weight.detach(), e.linear2.bias.detach(), k=h.shape[1])


class TopKScheduler(nn.Module):
    def __init__(self, n_experts=8, min_n_experts_per_token=2, update_step=16):
        super().__init__()
        self.n_experts = n_experts
        self.min_n_experts_per_token = min_n_experts_per_token
        self.k = n_experts
        self.update_step = update_step
        self.counting_step = 0

    def step(self):
        self.counting_step += 1
        if self.counting_step % self.update_step == 0:
            self.k = max(self.min_n_experts_per_token, self.k - 1)

    def getK(self):
        return self.k


class TemperatureScheduler(nn.Module):
    def __init__(self, temperature_min=0.8, temperature_max=1.1, temperature_step=0.0005):
        super().__init__()
        self.temperature_min = temperature_min
        self.temperature_max = temperature_max
        self.temperature_step = temperature_step
        self.t = self.temperature_min

    def step(self):
        self.t += self.temperature_step
        self.t = min(self.t, self.temperature_max)

    def getT(self):
        return self.t


def _stacked(experts):
    """Expert weights as [E, ...] stacks for the grouped GEMMs, rebuilt only when a parameter changed in place
    (optimizer step, load_state_dict) or moved."""
    ps = [p for e in experts for p in (e.linear1.weight, e.linear1.bias, e.gate.weight, e.gate.bias, e.linear2.weight, e.linear2.bias)]
    key = tuple((p.data_ptr(), p._version) for p in ps)
    cache = getattr(experts, "_v2m_stack", None)
    if cache is None or cache[0] != key:
        def st(f):
            return torch.stack([f(e).detach().float() for e in experts]).contiguous()
        cache = (key, (st(lambda e: e.linear1.weight), st(lambda e: e.linear1.bias), st(lambda e: e.gate.weight),
                       st(lambda e: e.gate.bias), st(lambda e: e.linear2.weight), st(lambda e: e.linear2.bias)))
        object.__setattr__(experts, "_v2m_stack", cache)
    return cache[1]


def _stacked_bf16(experts):
    """bf16 stacks for the tensor-core path: (linear1 | gate) weights as one [E, 2 ff, d] operand, linear2 as [E, d, ff];
    biases stay fp32."""
    ps = [p for e in experts for p in (e.linear1.weight, e.linear1.bias, e.gate.weight, e.gate.bias, e.linear2.weight, e.linear2.bias)]
    key = tuple((p.data_ptr(), p._version) for p in ps)
    holder = experts if isinstance(experts, nn.Module) else experts[0]      # a single (shared) expert caches on itself
    cache = getattr(holder, "_v2m_stack_bf16", None)
    if cache is None or cache[0] != key:
        w1g = torch.stack([torch.cat([e.linear1.weight.detach(), e.gate.weight.detach()], 0) for e in experts]).bfloat16().contiguous()
        b1g = torch.stack([torch.cat([e.linear1.bias.detach(), e.gate.bias.detach()], 0) for e in experts]).float().contiguous()
        w2 = torch.stack([e.linear2.weight.detach() for e in experts]).bfloat16().contiguous()
        b2 = torch.stack([e.linear2.bias.detach() for e in experts]).float().contiguous()
        cache = (key, (w1g, b1g, w2, b2))
        object.__setattr__(holder, "_v2m_stack_bf16", cache)
    return cache[1]


def _tc_ok(d: int, ff: int, d_out: int) -> bool:
    return d % 64 == 0 and ff % 64 == 0 and d_out % 128 == 0


def _glu_bf16(e: GLUExpert, x2: torch.Tensor) -> torch.Tensor:
    """One expert over all tokens (the shared expert) on the tcgen05 GEMM: stacked (linear1 | gate) projection, SwiGLU, linear2."""
    w1g, b1g, w2, b2 = _stacked_bf16([e])
    a = ops.linear(ops.cast_2d(x2, torch.bfloat16), w1g[0], b1g[0], out_dtype=torch.bfloat16)
    return ops.linear(ops.swiglu_pair(a), w2[0], b2[0], out_dtype=torch.float32)


def _experts_forward(experts, x2: torch.Tensor, idx: torch.Tensor, w: torch.Tensor, hist: torch.Tensor,
                     dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """out[t] = sum_r w[t,r] * expert_{idx[t,r]}(x[t])  (moe.py:191-199).
    Token copies are permuted into expert-contiguous order on the device, the experts run as two ragged grouped GEMMs
    (SwiGLU fused into the first), and the weighted results are combined per token in rank order (deterministic).
    No index tensors go through torch and nothing is read back to the host."""
    if dtype == torch.bfloat16:
        if not _tc_ok(x2.shape[1], experts[0].linear1.out_features, experts[0].linear2.out_features):
            raise NotImplementedError("bf16 MoE experts need d_model % 128 == 0 and d_ff % 64 == 0")
        return ops.moe_experts_bf16(x2, idx.reshape(-1, idx.shape[-1]), w.reshape(-1, w.shape[-1]), hist, *_stacked_bf16(experts))
    w1, b1, wg, bg, w2, b2 = _stacked(experts)
    return ops.moe_experts(x2, idx.reshape(-1, idx.shape[-1]), w.reshape(-1, w.shape[-1]), hist, w1, b1, wg, bg, w2, b2)


def balance_update_(bias: torch.Tensor, hist: torch.Tensor, rate: float, group=None) -> torch.Tensor:
    """Loss-free load balancing of SharedMoELayer (moe.py:270-279): bias += rate * (mean(c) - c), c = how many token copies
    each expert received in this batch.  Data parallel (one process per GPU, torch.distributed initialised): the histogram is
    sum-all-reduced first, so every replica applies the update of the GLOBAL batch and the `bias` buffers stay identical
    across ranks; the update is linear in c, so this equals the single-process update."""
    import torch.distributed as dist
    c = hist.to(bias.dtype)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        c = c.clone()
        dist.all_reduce(c, op=dist.ReduceOp.SUM, group=group)
    with torch.no_grad():
        bias += rate * (c.mean() - c).unsqueeze(1)
    return c


class MoELayer(nn.Module):
    def __init__(self, expert, d_model, n_experts=8, n_experts_per_token=2, dropout=0.1, topk_scheduler=None,
                 temperature_scheduler=None):
        super().__init__()
        self.n_experts = n_experts
        self.n_experts_per_token = n_experts_per_token
        self.d_model = d_model
        self.dropout = nn.Dropout(dropout)
        self.experts = _get_clones(expert, n_experts)
        self.gate = nn.Linear(d_model, n_experts)
        if topk_scheduler is not None:
            self.topk_scheduler = topk_scheduler
        if temperature_scheduler is not None:
            self.temperature_scheduler = temperature_scheduler
        self.on_route: Optional[Callable] = None      # hook(selected_experts, histogram, training)
        self.last_selected_experts = None

    def forward(self, x):
        if hasattr(self, "topk_scheduler") and self.training:            # moe.py:168-172
            self.topk_scheduler.step()
            k = self.topk_scheduler.getK()
        else:
            k = self.n_experts_per_token
        if hasattr(self, "temperature_scheduler") and self.training:     # moe.py:174-178
            self.temperature_scheduler.step()
            t = self.temperature_scheduler.getT()
        else:
            t = 1.0
        shp = x.shape
        x2 = x.detach().reshape(-1, shp[-1]).float().contiguous()
        idx, w, hist, _ = ops.moe_route(x2, self.gate.weight.detach(), self.gate.bias.detach(), k, inv_t_pre=1.0 / t)
        self.last_selected_experts = idx.view(shp[:-1] + (k,))
        if self.on_route is not None:
            self.on_route(self.last_selected_experts, hist, self.training)
        from . import autograd as ag
        drop = ag.has_dropout(self)                                      # nn.Dropout sites of moe.py:48,197 in training mode
        if drop and getattr(self, "compute_dtype", torch.float32) != torch.float32:
            raise NotImplementedError("training-mode dropout of the MoE layers is built on the fp32 path")
        if (ag.tracking(x, self) or drop) and getattr(self, "compute_dtype", torch.float32) == torch.float32:   # gradients of moe.py:180-199
            return ag.moe_experts_fn(self.experts, self.gate, ag.rows_f32(x), idx, w, hist, 1.0 / t, _stacked(self.experts),
                                     layer_dropout=self.dropout, training=self.training).view(shp)
        if ag.tracking(x, self) and _tc_ok(shp[-1], self.experts[0].linear1.out_features, self.experts[0].linear2.out_features):
            # bf16 compute dtype with gradients: forward and backward on the grouped tcgen05 GEMMs
            return ag.moe_experts_bf16_fn(self.experts, self.gate, ag.rows_f32(x), idx, w, hist, 1.0 / t, _stacked_bf16(self.experts)).view(shp)
        return _experts_forward(self.experts, x2, idx, w, hist, getattr(self, "compute_dtype", torch.float32)).view(shp)


class SharedMoELayer(nn.Module):
    def __init__(self, expert, d_model, n_experts=8, n_experts_per_token=2, dropout=0.1, balancing=False,
                 topk_scheduler=None, temperature_scheduler=None, use_KAN=False):
        super().__init__()
        if use_KAN:
            raise NotImplementedError("KAN gate needs efficient_kan (out of scope, SURVEY.md 8c)")
        self.n_experts = n_experts
        self.n_experts_per_token = n_experts_per_token
        self.d_model = d_model
        self.dropout = nn.Dropout(dropout)
        self.experts = _get_clones(expert, n_experts)
        self.balancing = balancing
        if topk_scheduler is not None:
            self.topk_scheduler = topk_scheduler
        if temperature_scheduler is not None:
            self.temperature_scheduler = temperature_scheduler
        self.gate = nn.Linear(d_model, n_experts)
        if self.balancing:
            self.register_buffer("bias", torch.zeros((n_experts, 1)))
            self.update_rate = 0.001
        self.shared_expert = _get_clones(expert, 1)[0]
        self.on_route: Optional[Callable] = None
        self.last_selected_experts = None

    def forward(self, x):
        if hasattr(self, "topk_scheduler") and self.training:            # moe.py:232-236
            self.topk_scheduler.step()
            k = self.topk_scheduler.getK()
        else:
            k = self.n_experts_per_token
        if hasattr(self, "temperature_scheduler"):                       # moe.py:238-242 (steps in eval too)
            self.temperature_scheduler.step()
            t = self.temperature_scheduler.getT()
        else:
            t = 1.0
        shp = x.shape
        x2 = x.detach().reshape(-1, shp[-1]).float().contiguous()
        sel_bias = None
        if self.balancing and self.training:                             # moe.py:258-268
            sel_bias = self.bias.detach().reshape(-1).float().contiguous()
        idx, w, hist, _ = ops.moe_route(x2, self.gate.weight.detach(), self.gate.bias.detach(), k, sel_bias=sel_bias,
                                        inv_t_post=1.0 / t)
        if self.balancing and self.training:                             # moe.py:270-279
            balance_update_(self.bias, hist, self.update_rate, getattr(self, "process_group", None))
        self.last_selected_experts = idx.view(shp[:-1] + (k,))
        if self.on_route is not None:
            self.on_route(self.last_selected_experts, hist, self.training)
        from . import autograd as ag
        drop = ag.has_dropout(self)                                      # nn.Dropout sites of moe.py:48,295 in training mode
        if drop and getattr(self, "compute_dtype", torch.float32) != torch.float32:
            raise NotImplementedError("training-mode dropout of the MoE layers is built on the fp32 path")
        if (ag.tracking(x, self) or drop) and getattr(self, "compute_dtype", torch.float32) == torch.float32:   # gradients of moe.py:244-301
            xr = ag.rows_f32(x)
            out = ag.moe_experts_fn(self.experts, self.gate, xr, idx, w, hist, 1.0 / t, _stacked(self.experts),
                                    layer_dropout=self.dropout, training=self.training)
            return ag.AddFn.apply(out, ag.glu_expert_fn(self.shared_expert, xr), 1.0 / k).view(shp[:-1] + (self.d_model,))
        dt = getattr(self, "compute_dtype", torch.float32)
        if ag.tracking(x, self) and _tc_ok(shp[-1], self.experts[0].linear1.out_features, self.experts[0].linear2.out_features):
            xr = ag.rows_f32(x)                                          # bf16 compute dtype with gradients: tensor-core training path
            out = ag.moe_experts_bf16_fn(self.experts, self.gate, xr, idx, w, hist, 1.0 / t, _stacked_bf16(self.experts))
            shared = ag.glu_expert_bf16_fn(self.shared_expert, xr, _stacked_bf16([self.shared_expert]))
            return ag.AddFn.apply(out, shared, 1.0 / k).view(shp[:-1] + (self.d_model,))
        out = _experts_forward(self.experts, x2, idx, w, hist, dt)
        shared = _glu_bf16(self.shared_expert, x2) if dt == torch.bfloat16 else _glu(self.shared_expert, x2)   # moe.py:301
        return ops.axpy(out, shared, 1.0 / k).view(shp[:-1] + (self.d_model,))
