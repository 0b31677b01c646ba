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
    from video2music_b200.trainer import noam_lr
    # utilities/lr_scheduling.py:38-45: linear warm-up to step 4000, then step^-0.5 decay
    assert abs(noam_lr(4000) - 512 ** -0.5 * 4000 ** -0.5) < 1e-12
    assert noam_lr(100) < noam_lr(200) < noam_lr(4000) > noam_lr(8000)
    assert abs(noam_lr(1) - 512 ** -0.5 * 4000 ** -1.5) < 1e-15
