"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU paths -- gradient all-reduce averaging of
the flat buffer, balanced sharding of samples / videos, Noam schedule.  No kernels are involved."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from video2music_b200.trainer import allreduce_mean_, shard_batch, shard_range
    g = torch.Generator().manual_seed(5)
    full = torch.randn(10, 7, generator=g)                       # per-sample "gradients" of a global batch of 10
    mine = shard_batch({"g": full}, rank, world)["g"]
    flat = mine.mean(0).clone()                                  # local mean over the shard (equal shard sizes)
    scale = allreduce_mean_(flat)
    ok = torch.allclose(flat * scale, full.mean(0), atol=1e-6)
    ranges = [shard_range(13, r, world) for r in range(world)]   # ragged: 13 videos over 2 ranks
    q.put((rank, bool(ok), scale, ranges))
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_and_sharding_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert all(p.exitcode == 0 for p in procs)
    for rank, ok, scale, ranges in res:
        assert ok and scale == 0.5
        assert ranges == [(0, 7), (7, 13)]


def test_noam_schedule_matches_reference_formula():
    from video2music_b200.trainer import noam_lr, scheduled_lr
    # utilities/lr_scheduling.py:38-45: linear warm-up to step 4000, then step^-0.5 decay
    assert abs(noam_lr(4000) - 512 ** -0.5 * 4000 ** -0.5) < 1e-12
    assert noam_lr(100) < noam_lr(200) < noam_lr(4000) > noam_lr(8000)
    assert abs(noam_lr(1) - 512 ** -0.5 * 4000 ** -1.5) < 1e-15
    assert noam_lr(0) == 0.0


def test_schedule_equals_lambdalr_over_the_same_closed_form():
    """train.py:252 + run_model_vevo.py:121-123: LambdaLR is built before the first optimiser step and stepped after each one,
    so optimiser step t runs with f(t - 1) and the first update has lr 0.  Checked against torch's own LambdaLR driving a
    torch.optim.Adam with LR_DEFAULT_START = 1.0."""
    from video2music_b200.trainer import noam_lr, scheduled_lr
    p = torch.nn.Parameter(torch.zeros(3))
    opt = torch.optim.Adam([p], lr=1.0, betas=(0.9, 0.98), eps=1e-8)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: noam_lr(s, 512, 10))
    for t in range(1, 30):
        used = opt.param_groups[0]["lr"]                 # the rate optimiser step t runs with
        assert abs(used - scheduled_lr(t, 512, 10)) < 1e-15, t
        p.grad = torch.ones(3)
        opt.step()
        sched.step()
    assert scheduled_lr(1) == 0.0


def _bucket_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from video2music_b200.trainer import FlatParams, GradBuckets, global_loss_norm
    from video2music_b200.moe import balance_update_
    torch.manual_seed(3)
    net = torch.nn.Sequential(torch.nn.Linear(24, 40), torch.nn.Tanh(), torch.nn.Linear(40, 40), torch.nn.Tanh(),
                              torch.nn.Linear(40, 8))
    unused = torch.nn.Linear(5, 5)                       # parameters that never receive a gradient (like Wout_root of the AMT)
    model = torch.nn.ModuleDict(dict(net=net, unused=unused))
    flat = FlatParams(model)
    buckets = GradBuckets(flat.params, flat.offset_list, flat.flat_g, None, bucket_bytes=4 * 1000)
    g = torch.Generator().manual_seed(11)
    X, Y = torch.randn(12, 24, generator=g), torch.randn(12, 8, generator=g)
    # single-process global-batch gradient
    ref = torch.nn.Sequential(*[type(m)(m.in_features, m.out_features) if isinstance(m, torch.nn.Linear) else torch.nn.Tanh() for m in net])
    ref.load_state_dict(net.state_dict())
    ((ref(X) - Y) ** 2).mean().backward()
    ref_g = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
    orders, ok = [], True
    lo, hi = rank * 6, rank * 6 + 6
    for step in range(3):
        flat.flat_g.zero_()
        buckets.start()
        ((net(X[lo:hi]) - Y[lo:hi]) ** 2).mean().backward()
        scale = buckets.finish()
        mine = torch.cat([p.grad.reshape(-1) for p in net.parameters()]) * scale
        ok = ok and torch.allclose(mine, ref_g, atol=1e-6) and bool((unused.weight.grad == 0).all())
        orders.append(list(buckets.order))
    # loss normalisers: ragged PAD counts per rank -> every rank gets global / world
    tgt = torch.full((4, 10), 7, dtype=torch.int64)
    tgt[:, 10 - 2 * (rank + 1):] = 158
    norm = global_loss_norm(tgt)
    # MoE balancing: ranks see different histograms, every rank applies the global-batch update
    bias = torch.zeros(6, 1)
    hist = torch.tensor([5, 1, 0, 2, 3, 1], dtype=torch.int32) if rank == 0 else torch.tensor([0, 4, 4, 1, 1, 2], dtype=torch.int32)
    balance_update_(bias, hist, 0.001)
    q.put((rank, ok, orders, len(buckets.bounds), norm.tolist(), bias.reshape(-1).tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_overlapped_allreduce_world2():
    """GradBuckets: the first step discovers which parameters get gradients (one flat all-reduce), later steps launch one
    async all-reduce per bucket from the backward hooks, last layers first; the averaged gradient equals the single-process
    global-batch gradient.  Also the global loss normalisers and the MoE balancing update."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_bucket_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert all(p.exitcode == 0 for p in procs)
    hist_sum = torch.tensor([5., 5, 4, 3, 4, 3])
    want_bias = (0.001 * (hist_sum.mean() - hist_sum)).tolist()
    for rank, ok, orders, n_buckets, norm, bias in res:
        assert ok
        assert n_buckets >= 3
        assert orders[0] == [-1]                                     # discovery step: flat
        assert sorted(orders[1]) == list(range(n_buckets)) and orders[1] == orders[2]
        assert orders[1][0] == n_buckets - 2 or orders[1][0] == n_buckets - 1   # the tail of the buffer completes first
        assert orders[1].index(0) > orders[1].index(n_buckets - 2)   # ... the head (first layer) after the later layers
        assert abs(norm[0] - (4 * 8 + 4 * 6) / 2) < 1e-6 and abs(norm[1] - 40.0) < 1e-6
        assert all(abs(a - b) < 1e-7 for a, b in zip(bias, want_bias))
