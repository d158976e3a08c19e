"""World-size-2 gloo tests (CPU) of the package's data-parallel host logic — anatomask_b200/parallel.py: batch sharding as
the DDP scripts do it (P/pretrain_DDP.py:251-290), the bucketed gradient exchange over the flat arena (group ranges in
backward-completion order, asynchronous start from backward marks, join), and the SUM-all-reduce + 1/world + global-norm
clip folding the fused AdamW kernel applies (host mirror `clip_scale`)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from anatomask_b200.parallel import GradBuckets, bucket_ranges, clip_scale, shard_batch


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


# a miniature of ParamArena.offsets: named_parameters order = encoder, decoder, densify norms / projs / tokens
OFFSETS = {'sparse_encoder.sp_cnn.conv_blocks_context.0.0.conv1.weight': (0, 100),
           'sparse_encoder.sp_cnn.conv_blocks_context.2.0.conv1.weight': (100, 60),
           'sparse_encoder.sp_cnn.conv_blocks_context.3.0.conv1.weight': (160, 140),
           'sparse_encoder.sp_cnn.conv_blocks_context.4.0.conv2.bias': (300, 20),
           'dense_decoder.dec.0.up_sample.weight': (320, 500), 'dense_decoder.proj.bias': (820, 4),
           'densify_norms.0.weight': (824, 16), 'densify_projs.1.weight': (840, 100), 'mask_tokens.0': (940, 24),
           'densify_norms.4.weight': (1000, 8),                       # dead parameter: beyond n_live
           'dense_decoder.dec.0.conv.1.running_mean': (1100, 16)}     # buffer: beyond n_live
N_LIVE = 964


def test_shard_batch_like_the_ddp_scripts():
    assert [shard_batch(12, 3, r) for r in range(3)] == [(0, 4), (4, 8), (8, 12)]
    assert [shard_batch(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert [shard_batch(16, 8, r) for r in range(8)] == [(2 * r, 2 * r + 2) for r in range(8)]
    assert [shard_batch(2, 4, r) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]        # ragged: empty tail ranks
    with pytest.raises(ValueError):
        shard_batch(4, 2, 2)


def test_bucket_ranges_follow_backward_completion_order():
    r = bucket_ranges(OFFSETS, N_LIVE)
    assert r == [('decoder', 320, 824), ('densify', 824, 964), ('encoder_deep', 160, 320), ('encoder', 0, 160)]
    flat = {n.replace('conv_blocks_context', 'stages'): v for n, v in OFFSETS.items()}      # an encoder without STUNet's stage names
    assert bucket_ranges(flat, N_LIVE) == [('decoder', 320, 824), ('densify', 824, 964), ('encoder', 0, 320)]
    broken = dict(OFFSETS)
    broken['sparse_encoder.late.weight'] = (900, 10)                  # an encoder tensor in the middle of another group
    with pytest.raises(RuntimeError):
        bucket_ranges(broken, N_LIVE)


def test_clip_scale_equals_clip_grad_norm_on_the_mean_gradient():
    g = torch.Generator().manual_seed(0)
    grads = [torch.randn(1000, generator=g) * 3 for _ in range(4)]
    summed = torch.stack(grads).sum(0)
    mean = torch.stack(grads).mean(0)
    for max_norm in (12.0, 200.0, None):
        s = clip_scale(float((summed.double() ** 2).sum()), 4, max_norm)
        p = torch.nn.Parameter(torch.zeros(1000))
        p.grad = mean.clone()
        if max_norm is not None:
            torch.nn.utils.clip_grad_norm_([p], max_norm)
        assert torch.allclose(summed * s, p.grad, rtol=1e-5, atol=1e-7)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        local = torch.randn(N_LIVE, generator=g)
        grad = local.clone()
        gathered = [torch.zeros(N_LIVE) for _ in range(world)]
        dist.all_gather(gathered, local)
        want = torch.stack(gathered).sum(0)
        b = GradBuckets(grad, OFFSETS, N_LIVE, dist.group.WORLD)
        ok = b.order == ['decoder', 'densify', 'encoder_deep', 'encoder']
        # a step: the decoder mark fires, then densify, then the deep encoder stages; the shallow group is started by finish()
        b.begin_step()
        b.start('decoder')
        b.start('decoder')                                            # idempotent within a step
        b.start('densify')
        b.start('encoder_deep')
        b.finish()
        ok = ok and torch.allclose(grad, want, atol=1e-6)
        # next step: no mark fires at all (eager path without marks) → finish() exchanges everything, exactly once
        grad.copy_(local)
        b.begin_step()
        b.finish()
        ok = ok and torch.allclose(grad, want, atol=1e-6)
        # folded clip: every rank derives the same multiplier from the summed gradient
        s = clip_scale(float((grad.double() ** 2).sum()), world, 12.0)
        mean = want / world
        ref = mean * min(12.0 / (float(mean.norm()) + 1e-6), 1.0)
        ok = ok and torch.allclose(grad * s, ref, rtol=1e-5, atol=1e-7)
        # batch sharding: the ranks' slices tile the global batch
        lo, hi = shard_batch(5, world, rank)
        cover = [torch.zeros(5) for _ in range(world)]
        mine = torch.zeros(5)
        mine[lo:hi] = 1
        dist.all_gather(cover, mine)
        ok = ok and bool((torch.stack(cover).sum(0) == 1).all())
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_world2_gloo_bucketed_gradient_exchange():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
