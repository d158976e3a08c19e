"""LightDecoder — P/decoder3D.py:13-84.  Same module tree (dec.{i}.up_sample / conv.{0,1,3,4}, proj) and initialisers; the
forward is ConvTranspose3d(k4,s2,p1) → conv3 → BN(+ReLU6 fused) → conv3 → BN on channels-last bf16 tensor-core kernels.
"""
from __future__ import annotations

import math
from typing import List

import torch
import torch.nn as nn

from . import ops
from ._lib import ACT_NONE, ACT_RELU6


def is_pow2n(x):
    return x > 0 and (x & (x - 1) == 0)


def _bn(bn: nn.Module, x, act, sums=None):
    """BatchNorm3d / SyncBatchNorm forward in the module's current mode."""
    if isinstance(bn, nn.InstanceNorm3d):
        raise NotImplementedError('LightDecoder(use_IN=True) has no sm_100a kernel (no shipped script uses it)')
    if bn.training:
        group = None
        if isinstance(bn, nn.SyncBatchNorm):
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                group = bn.process_group if bn.process_group is not None else dist.group.WORLD
        mom = 0.1 if bn.momentum is None else bn.momentum
        return ops.batch_norm_train(x, bn.weight, bn.bias, bn.eps, act,
                                    (bn.running_mean, bn.running_var, bn.num_batches_tracked), mom, group, sums)
    return ops.batch_norm_eval(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, act)


class UNetBlock(nn.Module):
    def __init__(self, cin, cout, bn3d):
        super().__init__()
        self.up_sample = nn.ConvTranspose3d(cin, cin, kernel_size=4, stride=2, padding=1, bias=True)
        self.conv = nn.Sequential(
            nn.Conv3d(cin, cin, kernel_size=3, stride=1, padding=1, bias=False), bn3d(cin), nn.ReLU6(inplace=True),
            nn.Conv3d(cin, cout, kernel_size=3, stride=1, padding=1, bias=False), bn3d(cout),
        )

    def forward_internal(self, x):
        c0, c3 = self.conv[0], self.conv[3]
        # the ConvTranspose's bias gradient (Σ over voxels of the gradient w.r.t. its output) rides on the input-gradient kernel
        # of the conv that consumes that output (ops.GradSumTap)
        tap = ops.GradSumTap() if (torch.is_grad_enabled() and self.up_sample.bias is not None) else None
        x = ops.conv_transpose3d(x, self.up_sample.weight, self.up_sample.bias, bias_grad_from=tap)
        b1, b4 = self.conv[1], self.conv[4]
        if not torch.is_grad_enabled() and all(isinstance(b, (nn.BatchNorm3d, nn.SyncBatchNorm)) and not b.training
                                               and b.track_running_stats for b in (b1, b4)):
            # inference (the EMA teacher): each BatchNorm (+ReLU6) is folded into the epilogue of the conv that feeds it
            x = ops.conv3d_bn_eval(x, c0.weight, b1.weight, b1.bias, b1.running_mean, b1.running_var, b1.eps, ACT_RELU6)
            return ops.conv3d_bn_eval(x, c3.weight, b4.weight, b4.bias, b4.running_mean, b4.running_var, b4.eps, ACT_NONE)
        # training-mode BN statistics come out of the conv epilogue (Σy, Σy² per channel)
        s0 = ops.new_stats(c0.out_channels, x.device) if (self.conv[1].training and ops.fused_stats_ok(c0.in_channels, c0.out_channels)) else None
        x = ops.conv3d(x, c0.weight, None, 3, 1, stats=s0, dx_sum=tap)
        x = _bn(self.conv[1], x, ACT_RELU6, s0)
        s3 = ops.new_stats(c3.out_channels, x.device) if (self.conv[4].training and ops.fused_stats_ok(c3.in_channels, c3.out_channels)) else None
        x = ops.conv3d(x, c3.weight, None, 3, 1, stats=s3)
        return _bn(self.conv[4], x, ACT_NONE, s3)

    def forward(self, x):
        return ops.to_external(self.forward_internal(ops.to_internal(x)))


class LightDecoder(nn.Module):
    def __init__(self, up_sample_ratio, width=768, sbn=True, use_IN=False, out_channel=1):
        super().__init__()
        self.width = width
        assert is_pow2n(up_sample_ratio)
        n = round(math.log2(up_sample_ratio))
        channels = [self.width // 2 ** i for i in range(n + 1)]
        if sbn:
            bn3d = nn.SyncBatchNorm
        elif use_IN:
            bn3d = nn.InstanceNorm3d
        else:
            bn3d = nn.BatchNorm3d
        self.dec = nn.ModuleList([UNetBlock(cin, cout, bn3d) for (cin, cout) in zip(channels[:-1], channels[1:])])
        self.proj = nn.Conv3d(channels[-1], out_channel, kernel_size=1, stride=1, bias=True)
        self.initialize()

    def forward(self, to_dec: List[torch.Tensor]):
        """to_dec: logical (N,C,D,H,W) tensors, coarse → fine.  Returns rec (N, out_channel, D, H, W) fp32."""
        x = None
        for i, d in enumerate(self.dec):
            if i < len(to_dec) and to_dec[i] is not None:
                t = ops.to_internal(to_dec[i])
                x = t if x is None else ops.AddFn.apply(x, t)
            x = d.forward_internal(x)
        if self.proj.out_channels != 1:
            raise NotImplementedError('LightDecoder.proj: only out_channel=1 has an sm_100a kernel')
        return ops.ProjFn.apply(x, self.proj.weight, self.proj.bias)

    def extra_repr(self) -> str:
        return f'width={self.width}'

    def initialize(self):
        # P/decoder3D.py:68-84 — Conv3d trunc-normal(.02), ConvTranspose3d kaiming-normal(fan_out), BN weight 1 / bias 0
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.ConvTranspose3d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0.)
            elif isinstance(m, (nn.BatchNorm3d, nn.SyncBatchNorm)):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)
