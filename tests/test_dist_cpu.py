"""World-size-2 gloo tests (CPU) of the host-side data-parallel logic: batch sharding as the DDP scripts do it
(P/pretrain_DDP.py:251-290) and the SUM-all-reduce + 1/world folding used by the arena optimiser."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def shard_batch(global_batch: int, world: int, rank: int):
    """ceil(global/world) per rank, the last rank takes the remainder (P/pretrain_DDP.py:251-290)."""
    per = -(-global_batch // world)
    lo = rank * per
    hi = min(global_batch, lo + per) if rank < world - 1 else global_batch
    return lo, max(lo, hi)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    grad = torch.randn(1000, generator=g)
    # arena path: SUM all-reduce, then the AdamW kernel multiplies by gscale = 1/world (clip uses the averaged norm)
    summed = grad.clone()
    dist.all_reduce(summed)
    avg = summed * (1.0 / world)
    gathered = [torch.zeros(1000) for _ in range(world)]
    dist.all_gather(gathered, grad)
    ref = torch.stack(gathered).mean(0)
    ok = torch.allclose(avg, ref, atol=1e-6)
    norm_sq = float((summed * summed).sum())
    coef = min(12.0 / (norm_sq ** 0.5 / world + 1e-6), 1.0) / world        # what adamw_kernel computes from Σg²
    ref_coef = min(12.0 / (float(ref.norm()) + 1e-6), 1.0)
    ok = ok and abs(coef * world - ref_coef) < 1e-6
    # SyncBN statistics: (Σx, Σx², n) ride in one all-reduce; pooled mean/var equal the global-batch statistics
    x = torch.randn(50 + 10 * rank, 8, generator=g)
    pack = torch.cat([x.sum(0), (x * x).sum(0), torch.tensor([float(x.shape[0])])]).double()
    dist.all_reduce(pack)
    n = pack[-1]
    mean, var = pack[:8] / n, pack[8:16] / n - (pack[:8] / n) ** 2
    allx = [torch.zeros(50 + 10 * r, 8) for r in range(world)]
    dist.all_gather_object(obj := [None] * world, x)
    full = torch.cat(obj)
    ok = ok and torch.allclose(mean.float(), full.mean(0), atol=1e-5) and \
        torch.allclose(var.float(), full.var(0, unbiased=False), atol=1e-5)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_shard_batch_like_the_ddp_scripts():
    assert [shard_batch(12, 3, r) for r in range(3)] == [(0, 4), (4, 8), (8, 12)]
    assert [shard_batch(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert [shard_batch(16, 8, r) for r in range(8)] == [(2 * r, 2 * r + 2) for r in range(8)]


def test_world2_gloo_allreduce_semantics():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
