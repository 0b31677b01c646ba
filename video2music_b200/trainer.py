"""Data-parallel training step for the AMT (replaces the body of train_epoch, utilities/run_model_vevo.py:84-124).

One process per GPU (the reference hard-codes cuda:0, utilities/device.py:8-9, so each process sees its GPU as
device 0 or sets the device explicitly), replicated weights, the global batch cut into equal per-rank shards, and ONE
exchange step per iteration: an all-reduce (NCCL over NVLink / NVSwitch on GPUs, gloo in the CPU tests) of the flat
fp32 gradient buffer (32.5 M parameters = 130 MB).  Parameters, gradients and the Adam moments live in flat buffers so
that the optimiser is one kernel launch and the collective one call.
"""
import math
from typing import Dict, Optional

import torch
import torch.distributed as dist


def noam_lr(step: int, d_model: int = 512, warmup: int = 4000, start: float = 1.0) -> float:
    """LrStepTracker.step (utilities/lr_scheduling.py:28-45): Noam / 'Attention is all you need' schedule."""
    step = max(step, 1)
    return start * (d_model ** -0.5) * min(step ** -0.5, step * warmup ** -1.5)


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced [begin, end) slice of n items for `rank` (videos for generation, samples for training)."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_batch(batch: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    n = next(iter(batch.values())).shape[0]
    b, e = shard_range(n, rank, world)
    return {k: v[b:e] for k, v in batch.items()}


def allreduce_mean_(flat: torch.Tensor, group=None) -> float:
    """Sum-all-reduce of the flat gradient buffer; returns the scale (1/world) the optimiser applies."""
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
        if world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        return 1.0 / world
    return 1.0


class FlatParams:
    """Re-homes every parameter of `model` into one contiguous fp32 buffer (and its gradient into another)."""

    def __init__(self, model: torch.nn.Module):
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        params = [p for _, p in named]
        self.params = params
        # every parameter starts on a 16-byte boundary of the bf16 mirror (TMA operands), the total is a multiple of 512
        offs, off = [], 0
        for p in params:
            offs.append(off)
            off += (p.numel() + 7) // 8 * 8
        n = (off + 511) // 512 * 512
        dev = params[0].device
        self.flat_p = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(n, device=dev, dtype=torch.float32)
        self.offsets = {}
        for (name, p), o in zip(named, offs):
            k = p.numel()
            self.flat_p[o:o + k].copy_(p.detach().reshape(-1))
            p.data = self.flat_p[o:o + k].view(p.shape)
            p.grad = self.flat_g[o:o + k].view(p.shape)
            self.offsets[name] = (o, tuple(p.shape))
        self.numel = n


class Trainer:
    def __init__(self, model, lr: Optional[float] = None, betas=(0.9, 0.98), eps: float = 1e-9, warmup: int = 4000,
                 group=None, use_graph: bool = False):
        """use_graph: after two eager steps the whole step (forward, loss, backward, all-reduce, Adam) is captured into ONE
        CUDA graph and replayed -- ~400 launches per step otherwise keep the host busier than a small per-GPU batch keeps the
        GPU (strong scaling over 8 GPUs).  Needs fixed batch shapes; the learning rate, Adam's bias corrections and the
        dropout seeds are read from device memory so that they keep changing across replays."""
        from .autograd import AmtLossFn  # noqa: F401  (fail early if the extension is missing)
        self.model = model
        self.flat = FlatParams(model)
        self.m = torch.zeros_like(self.flat.flat_p)
        self.v = torch.zeros_like(self.flat.flat_p)
        self.lr, self.betas, self.eps, self.warmup = lr, betas, eps, warmup
        self.group = group
        self.step_no = 0
        self.last_parts = None
        self._copy_stream = None
        # bf16 mirror of the flat parameter buffer: ONE cast kernel per optimiser step instead of one per weight matrix;
        # the model's bf16 operands are views into it (engine.AMTWeights.flat16)
        self.flat16 = None
        if self.flat.flat_p.is_cuda:
            self.flat16 = torch.empty(self.flat.numel, device=self.flat.flat_p.device, dtype=torch.bfloat16)
            self._views16 = {name: self.flat16[o:o + math.prod(shp)].view(shp) for name, (o, shp) in self.flat.offsets.items()}
            self._mirror()
        self.use_graph = bool(use_graph) and self.flat.flat_p.is_cuda
        self._graph = None
        self._static = None
        self._static_loss = None
        self.launches_per_step = None
        if self.use_graph:
            from . import ops
            dev = self.flat.flat_p.device
            self._dyn = torch.zeros(3, device=dev, dtype=torch.float32)       # lr, 1 - b1^t, 1 - b2^t
            self._ctr = torch.zeros(1, device=dev, dtype=torch.int32)         # step counter added to the dropout seeds

    def _mirror(self) -> None:
        """flat fp32 masters -> flat bf16 mirror, and hand the views to the model's weight resolvers."""
        from . import ops
        ops.cast_into(self.flat.flat_p.view(-1, 512), self.flat16.view(-1, 512))
        self._handoff()

    def _handoff(self) -> None:
        for w in getattr(self.model, "_weights", {}).values():
            if self.flat16 is not None:
                w.flat16 = self._views16
            w.invalidate()                      # per-weight copies (odd leading dimensions) are re-derived on the next forward

    def _lr(self) -> float:
        return self.lr if self.lr is not None else noam_lr(self.step_no, self.model.d_model, self.warmup)

    def _step_body(self, b: Dict[str, torch.Tensor], dyn: Optional[torch.Tensor] = None) -> torch.Tensor:
        """forward + loss + backward + the one exchange step + Adam (which also writes the bf16 mirror of the parameters and
        clears the gradients) -- everything the GPU does in one iteration, enqueued on the current stream."""
        from . import ops
        from .autograd import AmtLossFn
        m = self.model
        y = m(b["x"], b["x_root"], b["x_attr"], b["feature_semantic_list"], b["feature_key"], b["feature_scene_offset"],
              b["feature_motion"], b["feature_emotion"])
        loss = AmtLossFn.apply(y, b["tgt"], b["tgt_emotion"], 0.1, 0.4, 0.6)          # run_model_vevo.py:101-119
        loss.backward()                                                                  # accumulates into flat_g views
        scale = allreduce_mean_(self.flat.flat_g, self.group)                            # the one exchange step
        ops.adam_step(self.flat.flat_p, self.flat.flat_g, self.m, self.v, self._lr(), self.betas[0], self.betas[1], self.eps,
                      max(self.step_no, 1), grad_scale=scale, dyn=dyn, p16=self.flat16, zero_grad=True,
                      counter=self._ctr if self.use_graph else None)
        return loss.detach()

    def prefetch(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """Starts the host -> device copy of a (pinned) batch on a side stream and returns the device batch; train_step waits
        for it.  Called for step i+1 before train_step(i), the copy overlaps the compute of step i (the reference's DataLoader
        does the same job with pin_memory workers, train.py:160-170)."""
        dev = self.flat.flat_p.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._slots = [None, None]          # two persistent device batches (no allocator traffic in the steady state)
            self._slot_free = [None, None]      # event: the step that last read the slot has finished
            self._slot_i = 0
        i = self._slot_i
        self._slot_i ^= 1
        cs = self._copy_stream
        slot = self._slots[i]
        if slot is None or any(k not in slot or slot[k].shape != v.shape or slot[k].dtype != v.dtype for k, v in batch.items()):
            slot = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in batch.items()}
            self._slots[i] = slot
            cs.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(cs):
            if self._slot_free[i] is not None:
                cs.wait_event(self._slot_free[i])
            for k, v in batch.items():
                slot[k].copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        out = dict(slot)
        out["_ready"] = ev
        out["_slot"] = i
        return out

    def train_step(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        """batch: this rank's shard with the keys of synthetic.make_inputs (run_model_vevo.py:31-45), on the host or a
        device batch returned by prefetch()."""
        from . import ops
        from .autograd import AmtLossFn
        m = self.model
        dev = self.flat.flat_p.device
        ready, slot_i = batch.get("_ready"), batch.get("_slot")
        if ready is not None:
            torch.cuda.current_stream(dev).wait_event(ready)
        b = {k: v.to(dev, non_blocking=True) for k, v in batch.items() if not k.startswith("_")}
        self.step_no += 1
        from . import ops
        prev_seed_dev = ops.DROP_SEED_DEV
        if self.use_graph:
            ops.DROP_SEED_DEV = self._ctr        # only while this trainer's kernels are being enqueued / captured
        try:
            loss = self._run_step(b, dev)
        finally:
            ops.DROP_SEED_DEV = prev_seed_dev
        self._handoff()
        if slot_i is not None:                  # the prefetch slot may be overwritten once this step has run
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            self._slot_free[slot_i] = ev
        return loss.detach()

    def _run_step(self, b: Dict[str, torch.Tensor], dev) -> torch.Tensor:
        if not self.use_graph or self.step_no <= 2:
            loss = self._step_body(b)
        else:
            if self._static is None:
                self._static = {k: torch.empty_like(v) for k, v in b.items()}
            for k, v in b.items():
                self._static[k].copy_(v, non_blocking=True)
            b1, b2 = self.betas
            self._dyn.copy_(torch.tensor([self._lr(), 1.0 - b1 ** self.step_no, 1.0 - b2 ** self.step_no], dtype=torch.float32))
            if self._graph is None:
                from . import _lib
                torch.cuda.synchronize(dev)
                n0 = _lib.launches()
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    self._static_loss = self._step_body(self._static, dyn=self._dyn)
                self.launches_per_step = _lib.launches() - n0
            self._graph.replay()
            loss = self._static_loss
        return loss
