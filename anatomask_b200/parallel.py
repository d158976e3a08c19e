"""Data-parallel host logic of the pre-training step (P/pretrain_DDP.py:188-290).

One process per GPU, batch sharded across ranks, the only data-path collectives being the gradient all-reduce and — with
`sbn=True` — the SyncBN statistics.  Three pieces live here so that CPU (gloo) tests can exercise them without a GPU:

  shard_batch         the per-rank slice of a global batch, as the DDP scripts compute it
  clip_scale          the factor the fused AdamW kernel applies to a SUM-all-reduced gradient (1/world and the global-norm
                      clip folded together) — host mirror of `adamw_dev_kernel`'s arithmetic
  GradBuckets         the gradient arena split into parameter groups in the order the backward pass finishes them
                      (decoder → densify → deep encoder stages → shallow encoder stages); each group's all-reduce is
                      started from a backward mark
                      (ops.backward_mark) so it overlaps the remaining backward work, and joined before the optimiser
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch


def shard_batch(global_batch: int, world: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of this rank's samples: ceil(global/world) per rank, the last rank takes the remainder
    (P/pretrain_DDP.py:251-290)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f'shard_batch: bad world/rank {world}/{rank}')
    per = -(-global_batch // world)
    lo = min(global_batch, rank * per)
    hi = global_batch if rank == world - 1 else min(global_batch, lo + per)
    return lo, max(lo, hi)


def clip_scale(sumsq_of_summed_grads: float, world: int, max_norm: Optional[float]) -> float:
    """Multiplier that turns a SUM-all-reduced gradient into the clipped mean gradient:
    g_mean = g_sum / world;  ‖g_mean‖ = sqrt(Σ g_sum²) / world;  coef = min(max_norm / (‖g_mean‖ + 1e-6), 1)
    (torch.nn.utils.clip_grad_norm_, P/pretrain.py:408).  Returns coef / world."""
    gscale = 1.0 / world
    if max_norm is None or max_norm <= 0:
        return gscale
    norm = (sumsq_of_summed_grads ** 0.5) * gscale
    return min(max_norm / (norm + 1e-6), 1.0) * gscale


# parameter groups in backward-completion order (autograd runs later-created nodes first: the decoder's nodes all precede
# the densify nodes, which precede the encoder's; inside the encoder the deep stages — which hold 94 % of its parameters —
# finish first, so they travel as their own bucket while the shallow, large-extent stages are still in their backward pass).
# A name belongs to the FIRST group whose prefix it matches.
ENCODER_DEEP_FROM_STAGE = 3                   # STUNet conv_blocks_context.{3,4}: 13.4 M of STUNet-B's 14.3 M encoder parameters
GROUP_PREFIXES = (('decoder', ('dense_decoder.',)),
                  ('densify', ('densify_norms.', 'densify_projs.', 'mask_tokens.')),
                  ('encoder_deep', tuple(f'sparse_encoder.sp_cnn.conv_blocks_context.{s}.' for s in range(ENCODER_DEEP_FROM_STAGE, 8))),
                  ('encoder', ('sparse_encoder.',)))
# the last group is complete when backward returns
MARK_OF_GROUP = {'decoder': 'decoder_done', 'densify': 'densify_done', 'encoder_deep': 'encoder_deep_done'}


def bucket_ranges(offsets: Dict[str, Tuple[int, int]], n_live: int) -> List[Tuple[str, int, int]]:
    """[(group, lo, hi)] — element ranges of the live-gradient arena, one per parameter group, in backward-completion order.
    `offsets` is ParamArena.offsets (name -> (offset, numel)); entries at or beyond n_live (dead parameters, buffers) are
    ignored.  A range may contain the arena's alignment padding (zeros) but no tensor of another group; raises if the groups'
    ranges overlap or a live tensor belongs to no group."""
    out = []
    claimed = set()
    for group, prefixes in GROUP_PREFIXES:
        names = [n for n, (o, _) in offsets.items() if n.startswith(prefixes) and o < n_live and n not in claimed]
        claimed.update(names)
        if not names:
            continue
        lo = min(offsets[n][0] for n in names)
        hi = max(offsets[n][0] + offsets[n][1] for n in names)
        for n, (o, k) in offsets.items():
            if o < n_live and n not in names and o < hi and o + k > lo:
                raise RuntimeError(f'gradient arena: parameter group {group} is not contiguous ({n} lies inside it)')
        out.append((group, lo, hi))
    missing = [n for n, (o, _) in offsets.items() if o < n_live and n not in claimed]
    if missing:
        raise RuntimeError(f'gradient arena: {len(missing)} live tensors belong to no parameter group (e.g. {missing[0]})')
    return out


class GradBuckets:
    """Overlapped gradient exchange over the flat gradient arena.

    start(group) is called from the backward pass as soon as a group's gradients are complete: the all-reduce is issued
    asynchronously (NCCL's stream picks up after the main stream AND the side stream that carries the deferred weight
    gradients), finish() makes the main stream wait for all of them.  Works eagerly and under CUDA-graph capture
    (the NCCL kernels become nodes of the step graph)."""

    def __init__(self, grad: torch.Tensor, offsets: Dict[str, Tuple[int, int]], n_live: int, group, side_stream_fn=None):
        self.grad, self.group = grad, group
        self.ranges = {g: (lo, hi) for g, lo, hi in bucket_ranges(offsets, n_live)}
        self.order = [g for g, _, _ in bucket_ranges(offsets, n_live)]
        self._side = side_stream_fn
        self._works = []
        self._started = set()

    def begin_step(self):
        self._works, self._started = [], set()

    def start(self, group_name: str):
        import torch.distributed as dist
        if group_name in self._started or group_name not in self.ranges:
            return
        self._started.add(group_name)
        lo, hi = self.ranges[group_name]
        view = self.grad[lo:hi]
        if view.is_cuda:
            main = torch.cuda.current_stream()
            side = self._side() if self._side is not None else None
            if side is not None:
                side.wait_stream(main)              # the collective must see the main-stream gradients (norm params, dgrads)
                with torch.cuda.stream(side):       # ... and everything the deferred weight-gradient chain has enqueued so far
                    self._works.append(dist.all_reduce(view, group=self.group, async_op=True))
                return
        self._works.append(dist.all_reduce(view, group=self.group, async_op=True))

    def finish(self):
        """Start whatever has not been started (the encoder group; everything when no mark fired) and join."""
        for g in self.order:
            self.start(g)
        for w in self._works:
            w.wait()
        self._works = []
