"""Encoder-only STUNet (the `cnn` handed to SparseEncoder) — P/STUNet_head.py:8-103.

Same constructor, attribute and parameter names.  After `SparseEncoder` has swapped its layers, a BasicResBlock runs a
fused schedule on channels-last bf16: the Cin=1 stem computes conv1 and the 1×1 shortcut in one pass, norm+LeakyReLU and
norm+residual+LeakyReLU are single kernels.
"""
from __future__ import annotations

import torch
from torch import nn

from . import encoder3D, ops
from ._lib import ACT_LRELU


class Decoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.deep_supervision = True


class BasicResBlock(nn.Module):
    def __init__(self, input_channels, output_channels, kernel_size=3, padding=1, stride=1, use_1x1conv=False):
        super().__init__()
        self.conv1 = nn.Conv3d(input_channels, output_channels, kernel_size, stride=stride, padding=padding)
        self.norm1 = nn.InstanceNorm3d(output_channels, affine=True)
        self.act1 = nn.LeakyReLU(inplace=True)
        self.conv2 = nn.Conv3d(output_channels, output_channels, kernel_size, padding=padding)
        self.norm2 = nn.InstanceNorm3d(output_channels, affine=True)
        self.act2 = nn.LeakyReLU(inplace=True)
        self.conv3 = nn.Conv3d(input_channels, output_channels, kernel_size=1, stride=stride) if use_1x1conv else None

    def _fusable(self):
        return (isinstance(self.conv1, encoder3D.SparseConv3d) and isinstance(self.norm1, encoder3D.SparseInstanceNorm)
                and isinstance(self.norm2, encoder3D.SparseInstanceNorm))

    def forward(self, x):
        if not self._fusable():                 # un-converted (dense) or exotic head: plain module-by-module path
            y = self.act1(self.norm1(self.conv1(x)))
            y = self.norm2(self.conv2(y))
            if self.conv3:
                x = self.conv3(x)
            y = y + x
            return self.act2(y)
        m = encoder3D._mask_ctx()
        c1, c2, c3 = self.conv1, self.conv2, self.conv3
        stride = c1.stride[0]
        stem = c1.in_channels == 1
        # Engine mode (ops.LEAN_ZERO): which masked voxels of this block's sparse tensors have to be zero.  P = patch edge at the
        # block's output resolution.  With P >= 8 every kernel that reads these tensors walks the active-patch list: the 3x3x3
        # ones read one voxel beyond a visible patch ('shell'), norms / 1x1 convs / the stem weight gradient none at all
        # ('none').  The block OUTPUT feeds the next stage's stride-2 convs, whose patch edge is P/2, hence P >= 16 there.
        P = (x.shape[2] // stride) // m.fd
        lean = ops.LEAN_ZERO and P >= 8 and x.shape[3] // stride // m.fh == P and x.shape[4] // stride // m.fw == P
        z1 = ('shell', 'none' if stem else 'shell', 'full') if lean else ('full', 'full', 'full')
        z2 = ('shell' if P >= 16 else 'full', 'shell', 'none') if lean else ('full', 'full', 'full')
        zero_dx = not ops.LEAN_ZERO            # conv input gradients here are only ever read at visible voxels (norm backward)
        if stem:                                # stem: masked fp32 input → conv1 and shortcut in one kernel
            if c3 is None or stride != 1:
                raise NotImplementedError('in_channels=1 block needs stride 1 and a 1x1 shortcut')
            y, sc = ops.StemFn.apply(x, c1.weight, c1.bias, c3.weight, c3.bias, m, False)
            s1 = None
        else:
            xi = ops.to_internal(x)
            # Σy / Σy² of the visible outputs come out of the conv epilogue (no separate statistics pass)
            s1 = ops.new_stats(c1.out_channels, xi.device) if ops.fused_stats_ok(c1.in_channels, c1.out_channels) else None
            # conv outputs feed pooled masked norms only (visible voxels), so their masked voxels may stay unwritten
            if c3 is not None and ops.LEAN_ZERO:
                y, sc = ops.conv3d_pair(xi, c1.weight, c1.bias, c3.weight, c3.bias, c1.kernel_size[0], stride, m, stats=s1)
            else:
                y = ops.conv3d(xi, c1.weight, c1.bias, c1.kernel_size[0], stride, m, stats=s1, zero_inactive=False,
                               zero_bias_grad=True, zero_dx=zero_dx)
                sc = ops.conv3d(xi, c3.weight, c3.bias, 1, stride, m, zero_inactive=False) if c3 is not None else xi
        y = ops.masked_norm(y, self.norm1.weight, self.norm1.bias, self.norm1.eps, m, ACT_LRELU, sums=s1, zero=z1)
        s2 = ops.new_stats(c2.out_channels, y.device) if ops.fused_stats_ok(c2.in_channels, c2.out_channels) else None
        # conv1 / conv2 feed SparseInstanceNorm only (batch statistics in train AND eval): their bias gradients are
        # identically zero, so no Σdy pass is run for them (ops.conv3d zero_bias_grad)
        y = ops.conv3d(y, c2.weight, c2.bias, c2.kernel_size[0], 1, m, stats=s2, zero_inactive=False, zero_bias_grad=True,
                       zero_dx=zero_dx)
        y = ops.masked_norm(y, self.norm2.weight, self.norm2.bias, self.norm2.eps, m, ACT_LRELU, residual=sc, sums=s2, zero=z2)
        return ops.to_external(y)


class STUNet(nn.Module):
    def __init__(self, input_channels, num_classes, depth=[1, 1, 1, 1, 1, 1], dims=[32, 64, 128, 256, 512, 512],
                 pool_op_kernel_sizes=None, conv_kernel_sizes=None, enable_deep_supervision=True):
        super().__init__()
        n_stages = len(pool_op_kernel_sizes)
        assert n_stages == len(dims) - 1
        self.conv_op, self.input_channels, self.num_classes = nn.Conv3d, input_channels, num_classes
        self.final_nonlin, self.upscale_logits = (lambda x: x), False
        self.decoder = Decoder()
        self.decoder.deep_supervision = enable_deep_supervision
        self.dims, self.pool_op_kernel_sizes, self.conv_kernel_sizes = dims, pool_op_kernel_sizes, conv_kernel_sizes
        self.conv_pad_sizes = [[k // 2 for k in ks] for ks in conv_kernel_sizes]

        def stage(s: int) -> nn.Sequential:
            cin = input_channels if s == 0 else dims[s - 1]
            extra = {} if s == 0 else {'stride': pool_op_kernel_sizes[s - 1]}
            k, p = conv_kernel_sizes[s], self.conv_pad_sizes[s]
            blocks = [BasicResBlock(cin, dims[s], k, p, use_1x1conv=True, **extra)]
            blocks += [BasicResBlock(dims[s], dims[s], k, p) for _ in range(depth[s] - 1)]
            return nn.Sequential(*blocks)

        # five encoder stages (the sixth width/depth entry belongs to the full U-Net bottleneck and is unused here)
        self.conv_blocks_context = nn.ModuleList(stage(s) for s in range(n_stages))

    def get_downsample_ratio(self) -> int:
        return 16

    def get_feature_map_channels(self):
        return self.dims[:5]

    def forward(self, x, hierarchical=False):
        from .parallel import ENCODER_DEEP_FROM_STAGE
        feats = []
        for s, blocks in enumerate(self.conv_blocks_context):
            if s == ENCODER_DEEP_FROM_STAGE:
                # backward reaches this point when the deep stages — most of the encoder's parameters — have their
                # gradients: their all-reduce overlaps the shallow stages' backward pass (parallel.GradBuckets)
                xin, = ops.backward_mark('encoder_deep_done', x)
            else:
                xin = x
            x = blocks(xin)
            feats.append(x)
        return feats if hierarchical else x
