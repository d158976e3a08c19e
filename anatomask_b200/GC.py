"""Activation-checkpointed variants of the encoder head and the decoder — P/GC.py:61-74,320-329.

The reference keeps a second copy of `LightDecoder` and `STUNet` whose forwards wrap every decoder block / encoder stage in
`torch.utils.checkpoint.checkpoint` so that STUNet-H fits in memory.  Here they are thin subclasses of the regular mirrors
(same constructors, same state-dict keys): each stage's activations are dropped after the forward pass and the stage's
kernels run again during backward.  Non-reentrant checkpointing is used: the reference's call (reentrant, the default of its
torch version) returns no gradients for a stage whose only tensor input does not require grad — the first encoder stage,
which is fed the input volume — while the arithmetic is otherwise identical.  As in the reference, a training-mode
BatchNorm inside a checkpointed block updates its running statistics twice per step (forward + recomputation).
"""
from __future__ import annotations

from typing import List

import torch
from torch.utils.checkpoint import checkpoint

from . import STUNet_head, decoder3D, ops


class LightDecoder(decoder3D.LightDecoder):
    def forward(self, to_dec: List[torch.Tensor]):
        x = None
        for i, d in enumerate(self.dec):
            if i < len(to_dec) and to_dec[i] is not None:
                t = ops.to_internal(to_dec[i])
                x = t if x is None else ops.AddFn.apply(x, t)
            x = checkpoint(d.forward_internal, x, use_reentrant=False)
        if self.proj.out_channels != 1:
            raise NotImplementedError('LightDecoder.proj: only out_channel=1 has an sm_100a kernel')
        return ops.ProjFn.apply(x, self.proj.weight, self.proj.bias)


class STUNet(STUNet_head.STUNet):
    def forward(self, x, hierarchical=False):
        feats = []
        for blocks in self.conv_blocks_context:
            x = checkpoint(blocks, x, use_reentrant=False)
            feats.append(x)
        return feats if hierarchical else x


BasicResBlock = STUNet_head.BasicResBlock
UNetBlock = decoder3D.UNetBlock
