"""Multi-rank host logic on CPU (gloo, world_size 2): batch sharding and the gradient/loss bucket mean.
The kernels themselves need a GPU; what is checked here is the data-parallel plumbing around them."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from eqnn_jax_b200.train import allreduce_mean_, cosine_decay_lr, shard_batch
    full = torch.arange(8 * 3, dtype=torch.float32).reshape(8, 3)
    mine = shard_batch(full, rank, world)
    # a fake "gradient + loss" bucket that depends on the shard
    bucket = torch.cat([mine.sum(0), mine.mean().reshape(1)])
    allreduce_mean_(bucket, world)
    want = torch.cat([sum(shard_batch(full, r, world).sum(0) for r in range(world)) / world,
                      torch.stack([shard_batch(full, r, world).mean() for r in range(world)]).mean().reshape(1)])
    ok = torch.allclose(bucket, want) and tuple(mine.shape) == (4, 3) and float(mine[0, 0]) == rank * 12.0
    ok = ok and abs(cosine_decay_lr(0, 3e-4, 2000) - 3e-4) < 1e-12 and cosine_decay_lr(2000, 3e-4, 2000) < 1e-12
    out.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_bucket_mean_and_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}


def test_shard_batch_rejects_ragged():
    from eqnn_jax_b200.train import shard_batch
    with pytest.raises(ValueError):
        shard_batch(torch.zeros(7, 2), 0, 2)
