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


def noam_lr(step: int, d_model: int = 512, warmup: int = 4000, start: float = 1.0, init_steps: int = 0) -> float:
    """LrStepTracker.step (utilities/lr_scheduling.py:28-45), literally: the multiplier LambdaLR applies to the base learning
    rate LR_DEFAULT_START = 1.0 -- linear warm-up `step * warmup^-1.5 / sqrt(d_model)` up to `warmup`, then `1 / sqrt(d_model
    * step)`.  step 0 gives 0."""
    step = step + init_steps
    if step <= warmup:
        return start * (d_model ** -0.5) * (warmup ** -1.5) * step
    return start * (d_model ** -0.5) * (step ** -0.5)


def scheduled_lr(opt_step: int, d_model: int = 512, warmup: int = 4000, init_steps: int = 0) -> float:
    """Learning rate of the `opt_step`-th optimiser step (1-based) under the reference's LambdaLR(opt, LrStepTracker.step)
    (train.py:252-253): the scheduler is constructed before the first step (lr = f(0) = 0) and stepped AFTER each
    optimiser step (run_model_vevo.py:121-123), so step t runs with f(t - 1) and the very first update has lr = 0."""
    return noam_lr(opt_step - 1, d_model, warmup, init_steps=init_steps)


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


def global_loss_norm(tgt: torch.Tensor, group=None, ignore: int = 158) -> torch.Tensor:
    """fp32 [n, r] on tgt's device: n = (non-PAD targets over ALL ranks) / world, r = (target rows over all ranks) / world.
    A rank that normalises its CE sum by n and its BCE sum by r * C, followed by the 1/world gradient mean, reproduces the
    reference's single-process loss over the global batch: CrossEntropyLoss(ignore_index) is a mean over non-PAD targets
    (train.py:222), BCEWithLogitsLoss a mean over all elements (train.py:233)."""
    world = dist.get_world_size(group)
    both = torch.empty(2, device=tgt.device, dtype=torch.float32)
    if tgt.is_cuda:
        from . import ops
        both[:1].copy_(ops.count_valid(tgt, ignore))
    else:
        both[0] = float((tgt != ignore).sum())
    both[1:].fill_(float(tgt.numel()))
    dist.all_reduce(both, op=dist.ReduceOp.SUM, group=group)
    return both / world


class GradBuckets:
    """Bucketed gradient all-reduce overlapped with the backward pass (the reference is single-GPU, this is
    the one exchange step data parallelism adds to run_model_vevo.py:84-124).

    The flat gradient buffer is cut at parameter boundaries into contiguous buckets of about `bucket_bytes`.  Backward
    produces gradients roughly in reverse parameter order (output head, decoder layers 5..0, encoder layers 5..0, input
    projections), so buckets complete from the tail of the buffer towards its head.  A post-accumulate hook per parameter
    counts arrivals; when the last expected gradient of a bucket has been accumulated, the bucket's slice is all-reduced
    asynchronously (NCCL runs it on its own stream, ordered after the gradient kernels by an event) while backward continues
    with the earlier layers.  `finish()` waits for all of them and reduces whatever did not fire (parameters without a
    gradient in this step keep zeros, which reduce to zeros).  Which parameters receive gradients is learnt in the first
    step (`expected` = None: that step falls back to one flat all-reduce).  Works under CUDA-graph capture: the async
    collectives become parallel branches of the captured graph."""

    def __init__(self, params, offsets, flat_g: torch.Tensor, group=None, bucket_bytes: int = 16 << 20):
        self.flat_g, self.group = flat_g, group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.bounds = []                      # [begin, end) element ranges of the buckets, in buffer order
        self.bucket_of = []                   # parameter index -> bucket index
        begin, cur = 0, 0
        per = max(1, bucket_bytes // 4)
        for i, (p, o) in enumerate(zip(params, offsets)):
            end = o + (p.numel() + 7) // 8 * 8
            self.bucket_of.append(len(self.bounds))
            cur = end
            if cur - begin >= per:
                self.bounds.append((begin, cur))
                begin = cur
        if cur > begin or not self.bounds:
            self.bounds.append((begin, flat_g.numel()))
        else:
            self.bounds[-1] = (self.bounds[-1][0], flat_g.numel())
        self.bucket_of = [min(b, len(self.bounds) - 1) for b in self.bucket_of]
        self.expected = None                  # per bucket: number of gradient-complete signals per step
        self._seen = [0] * len(self.bounds)
        self._fired = set()
        self._work = []
        self.order = []                       # bucket indices in launch order of the last step (diagnostics / tests)
        self.enabled = self.world > 1
        for i, p in enumerate(params):
            hook = self._make_hook(i)
            p.register_post_accumulate_grad_hook(hook)
            # backward kernels that add straight into the flat buffer (autograd.direct_grad) bypass AccumulateGrad and its
            # hook; they call this instead -- once per use of the parameter, the same number of times every step
            p._v2m_grad_ready = hook

    def _make_hook(self, i):
        def hook(_p=None):
            if not self.enabled:
                return
            b = self.bucket_of[i]
            self._seen[b] += 1
            if self.expected is not None and self._seen[b] == self.expected[b] and b not in self._fired:
                self._launch(b)
        return hook

    def _launch(self, b):
        lo, hi = self.bounds[b]
        self._fired.add(b)
        self.order.append(b)
        self._work.append(dist.all_reduce(self.flat_g[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def start(self):
        self._seen = [0] * len(self.bounds)
        self._fired, self._work, self.order = set(), [], []

    def finish(self) -> float:
        """Waits for the in-flight buckets, reduces the ones that never fired; returns the scale (1 / world) that turns the
        summed gradients into the mean the optimiser applies."""
        if not self.enabled:
            return 1.0
        if self.expected is None:
            # discovery step: one flat all-reduce, remember which parameters produced gradients
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.group)
            self.expected = list(self._seen)
            self.order = [-1]
            return 1.0 / self.world
        for b in range(len(self.bounds) - 1, -1, -1):
            if b not in self._fired:
                self._launch(b)
        for w in self._work:
            w.wait()
        self._work = []
        return 1.0 / self.world


class FlatParams:
    """Re-homes every parameter of `model` into one contiguous fp32 buffer (and its gradient into another)."""

    def __init__(self, model: torch.nn.Module, direct: bool = True):
        """direct: mark the parameters so that the backward kernels accumulate into flat_g directly (autograd.direct_grad);
        requires that whoever owns the buffers clears flat_g after every optimiser step (Trainer's Adam kernel does)."""
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
        self.offset_list = offs
        for (name, p), o in zip(named, offs):
            k = p.numel()
            self.flat_p[o:o + k].copy_(p.detach().reshape(-1))
            p.data = self.flat_p[o:o + k].view(p.shape)
            p.grad = self.flat_g[o:o + k].view(p.shape)
            if direct:
                p._v2m_direct_grad = True
            self.offsets[name] = (o, tuple(p.shape))
        self.numel = n


class Trainer:
    def __init__(self, model, lr: Optional[float] = None, betas=(0.9, 0.98), eps: float = 1e-8, warmup: int = 4000,
                 group=None, use_graph: bool = False, optimizer: str = "Adam", weight_decay: Optional[float] = None,
                 bucket_mb: float = 16.0, overlap: bool = True, broadcast_init: bool = True, global_loss_norm: bool = True,
                 init_steps: int = 0, direct_grads: bool = True):
        """lr None: the reference's schedule (LambdaLR over LrStepTracker, train.py:252; see scheduled_lr), else constant.
        eps: ADAM_EPSILON = 10e-9 = 1e-8 (utilities/constants.py:91).  optimizer: "Adam" (train.py:237-238) or "AdamW"
        (train.py:239-240, the CLI default, argument_funcs.py:17; decoupled weight_decay, torch's default 0.01).
        Data parallel (world > 1): rank 0's parameters are broadcast at construction (replicas must start equal), the
        gradient exchange is bucketed and overlapped with backward (`overlap`, `bucket_mb`; GradBuckets), and with
        `global_loss_norm` the loss normalisers (count of non-PAD targets, row count) are all-reduced first so that the averaged
        gradient is the gradient of the reference's single-process global-batch loss, also with ragged PAD tails and
        unequal shards.
        use_graph: after two eager steps the whole step (forward, loss, backward, all-reduce, Adam) is captured into ONE
        CUDA graph and replayed -- ~400 launches per step otherwise keep the host busier than a small per-GPU batch keeps the
        GPU (strong scaling over 8 GPUs).  Needs fixed batch shapes; the learning rate, Adam's bias corrections and the
        dropout seeds are read from device memory so that they keep changing across replays."""
        from .autograd import AmtLossFn  # noqa: F401  (fail early if the extension is missing)
        self.model = model
        self.flat = FlatParams(model, direct=direct_grads)
        self.m = torch.zeros_like(self.flat.flat_p)
        self.v = torch.zeros_like(self.flat.flat_p)
        self.lr, self.betas, self.eps, self.warmup = lr, betas, eps, warmup
        if optimizer not in ("Adam", "AdamW"):
            raise ValueError("optimizer must be 'Adam' or 'AdamW' (RAdam / RAdanW / Lion of train.py:241-250 are out of the hot path)")
        self.weight_decay = float(weight_decay) if weight_decay is not None else (0.01 if optimizer == "AdamW" else 0.0)
        if optimizer == "Adam" and self.weight_decay != 0.0:
            raise ValueError("Adam with (coupled) weight_decay is not built; use optimizer='AdamW'")
        self.group = group
        self.init_steps = init_steps
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.global_loss_norm = bool(global_loss_norm) and self.world > 1
        if self.world > 1 and broadcast_init:
            dist.broadcast(self.flat.flat_p, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self.buckets = GradBuckets(self.flat.params, self.flat.offset_list, self.flat.flat_g, group,
                                   int(bucket_mb * (1 << 20))) if (overlap and self.world > 1) else None
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
        return self.lr if self.lr is not None else scheduled_lr(self.step_no, self.model.d_model, self.warmup, self.init_steps)

    def _loss_norm(self, tgt: torch.Tensor):
        """[n_valid, rows] normalisers of this rank's loss such that mean-over-ranks == the global-batch loss: one all-reduce
        of two floats [non-PAD targets, rows] over the ranks, divided by world (device-side, no host sync)."""
        if not self.global_loss_norm:
            return None
        both = global_loss_norm(tgt, self.group)
        return both

    def _step_body(self, b: Dict[str, torch.Tensor], dyn: Optional[torch.Tensor] = None) -> torch.Tensor:
        """forward + loss + backward + the one exchange step + Adam (which also writes the bf16 mirror of the parameters and
        clears the gradients) -- everything the GPU does in one iteration, enqueued on the current stream."""
        from . import ops
        from .autograd import AmtLossFn
        m = self.model
        y = m(b["x"], b["x_root"], b["x_attr"], b["feature_semantic_list"], b["feature_key"], b["feature_scene_offset"],
              b["feature_motion"], b["feature_emotion"])
        loss = AmtLossFn.apply(y, b["tgt"], b["tgt_emotion"], 0.1, 0.4, 0.6, self._loss_norm(b["tgt"]))   # run_model_vevo.py:101-119
        if self.buckets is not None:
            self.buckets.start()
            loss.backward()                                   # accumulates into flat_g views; buckets reduce as they complete
            scale = self.buckets.finish()
        else:
            loss.backward()
            scale = allreduce_mean_(self.flat.flat_g, self.group)                        # the one exchange step, unbucketed
        ops.adam_step(self.flat.flat_p, self.flat.flat_g, self.m, self.v, self._lr(), self.betas[0], self.betas[1], self.eps,
                      max(self.step_no, 1), grad_scale=scale, dyn=dyn, p16=self.flat16, zero_grad=True,
                      counter=self._ctr if self.use_graph else None, weight_decay=self.weight_decay)
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
