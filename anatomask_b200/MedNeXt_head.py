"""MedNeXt encoder head of the reference (P/MedNeXt_head.py) for the sparse pre-training path — SURVEY.md §8f row 4.

Same class names, constructor arguments and state-dict keys (`stem`, `enc_block_{i}.{j}.conv1|norm|conv2|conv3`,
`down_{i}.…`, `bottleneck.…`, `dummy_tensor`, `out_{i}.conv_out`) so that checkpoints interchange.  The head is built from
plain nn layers exactly like the reference's and becomes sparse through `SparseEncoder.dense_model_to_sparse`
(P/encoder3D.py:298-364): every nn.Conv3d → SparseConv3d (depthwise k³ → `amb_dwconv3d`, 1×1×1 → the tcgen05 implicit GEMM),
nn.GroupNorm → SparseGroupNorm (`amb_voxel_norm_*`).  GELU runs on `amb_gelu` when it sees the internal channels-last
bf16 layout.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.utils.checkpoint as checkpoint

from . import ops


class GELU(nn.GELU):
    """nn.GELU whose CUDA path is the library's kernel (exact erf form, like nn.GELU())."""

    def forward(self, x):
        if x.is_cuda and x.dtype == ops.bf16 and x.ndim == 5 and x.permute(0, 2, 3, 4, 1).is_contiguous():
            return ops.to_external(ops.gelu(x.permute(0, 2, 3, 4, 1)))
        return super().forward(x)


class LayerNorm(nn.Module):
    """P/MedNeXt_head.py:374-396 (channels_last → F.layer_norm; channels_first → the explicit per-voxel formula)."""

    def __init__(self, normalized_shape, eps=1e-5, data_format='channels_last'):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        self.bias = nn.Parameter(torch.zeros(normalized_shape))
        self.eps, self.data_format = eps, data_format
        if data_format not in ('channels_last', 'channels_first'):
            raise NotImplementedError
        self.normalized_shape = (normalized_shape,)

    def forward(self, x, dummy_tensor=False):
        if self.data_format == 'channels_last':
            return F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
        if x.is_cuda and x.ndim == 5:
            y = ops.voxel_norm(ops.to_internal(x), self.weight, self.bias, 1, self.eps, None)
            return ops.to_external(y)
        u = x.mean(1, keepdim=True)
        s = (x - u).pow(2).mean(1, keepdim=True)
        x = (x - u) / torch.sqrt(s + self.eps)
        return self.weight[:, None, None, None] * x + self.bias[:, None, None, None]


class MedNeXtBlock(nn.Module):
    """depthwise k³ conv → norm → 1×1×1 expansion → GELU → 1×1×1 compression (+ residual) — P/MedNeXt_head.py:233-296."""

    def __init__(self, in_channels, out_channels, exp_r=4, kernel_size=7, do_res=True, norm_type='group', n_groups=None,
                 dim='3d'):
        super().__init__()
        if dim != '3d':
            raise NotImplementedError('only the 3-D head is on the pre-training path')
        self.do_res = do_res
        self.conv1 = nn.Conv3d(in_channels, in_channels, kernel_size, 1, kernel_size // 2,
                               groups=in_channels if n_groups is None else n_groups)
        if norm_type == 'group':
            self.norm = nn.GroupNorm(num_groups=in_channels, num_channels=in_channels)
        elif norm_type == 'layer':
            self.norm = LayerNorm(normalized_shape=in_channels, data_format='channels_first')
        self.conv2 = nn.Conv3d(in_channels, exp_r * in_channels, 1, 1, 0)
        self.act = GELU()
        self.conv3 = nn.Conv3d(exp_r * in_channels, out_channels, 1, 1, 0)

    def forward(self, x, dummy_tensor=None):
        x1 = self.conv3(self.act(self.conv2(self.norm(self.conv1(x)))))
        return x + x1 if self.do_res else x1


class MedNeXtDownBlock(MedNeXtBlock):
    """The same block with a stride-2 depthwise conv and an optional 1×1×1 stride-2 residual — P/MedNeXt_head.py:299-357."""

    def __init__(self, in_channels, out_channels, exp_r=4, kernel_size=7, do_res=False, norm_type='group', dim='3d'):
        super().__init__(in_channels, out_channels, exp_r, kernel_size, do_res=False, norm_type=norm_type, dim=dim)
        self.resample_do_res = do_res
        if do_res:
            self.res_conv = nn.Conv3d(in_channels, out_channels, kernel_size=1, stride=2)
        self.conv1 = nn.Conv3d(in_channels, in_channels, kernel_size, 2, kernel_size // 2, groups=in_channels)

    def forward(self, x, dummy_tensor=None):
        x1 = super().forward(x)
        return x1 + self.res_conv(x) if self.resample_do_res else x1


class OutBlock(nn.Module):
    def __init__(self, in_channels, n_classes, dim):
        super().__init__()
        self.conv_out = nn.ConvTranspose3d(in_channels, n_classes, kernel_size=1)

    def forward(self, x, dummy_tensor=None):
        return self.conv_out(x)


class Decoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.deep_supervision = True


class MedNeXt(nn.Module):
    """Encoder half of MedNeXt as the reference's pre-training head keeps it (P/MedNeXt_head.py:11-230): stem, four
    (blocks, down) stages and the bottleneck; `forward(x, hierarchical=True)` returns the five feature maps."""

    def __init__(self, in_channels, n_channels, n_classes, exp_r=4, kernel_size=7, enc_kernel_size=None, dec_kernel_size=None,
                 deep_supervision=False, do_res=False, do_res_up_down=False, checkpoint_style=None,
                 block_counts=(2, 2, 2, 2, 2, 2, 2, 2, 2), enc_norm_type='group', dec_norm_type='group', dim='3d'):
        super().__init__()
        self.decoder = Decoder()
        self.decoder.deep_supervision = deep_supervision
        self.do_ds = deep_supervision
        self.n_channels = n_channels
        assert checkpoint_style in [None, 'outside_block']
        self.inside_block_checkpointing = False
        self.outside_block_checkpointing = checkpoint_style == 'outside_block'
        assert dim == '3d', 'only the 3-D head is on the pre-training path'
        if kernel_size is not None:
            enc_kernel_size = dec_kernel_size = kernel_size
        self.stem = nn.Conv3d(in_channels, n_channels, kernel_size=1)
        if isinstance(exp_r, int):
            exp_r = [exp_r] * len(block_counts)
        for i in range(4):
            c = n_channels * 2 ** i
            setattr(self, f'enc_block_{i}', nn.Sequential(*[
                MedNeXtBlock(c, c, exp_r[i], enc_kernel_size, do_res, enc_norm_type, dim=dim) for _ in range(block_counts[i])]))
            setattr(self, f'down_{i}', MedNeXtDownBlock(c, 2 * c, exp_r[i + 1], enc_kernel_size, do_res_up_down, enc_norm_type, dim=dim))
        self.bottleneck = nn.Sequential(*[
            MedNeXtBlock(n_channels * 16, n_channels * 16, exp_r[4], dec_kernel_size, do_res, enc_norm_type, dim=dim)
            for _ in range(block_counts[4])])
        self.dummy_tensor = nn.Parameter(torch.tensor([1.]), requires_grad=True)
        if self.do_ds:
            for i in range(1, 5):
                setattr(self, f'out_{i}', OutBlock(n_channels * 2 ** i, n_classes, dim))
        self.block_counts = block_counts

    def get_downsample_ratio(self) -> int:
        return 16

    def get_feature_map_channels(self):
        return [self.n_channels * 2 ** i for i in range(5)]

    def iterative_checkpoint(self, sequential_block, x):
        for blk in sequential_block:
            x = checkpoint.checkpoint(blk, x, self.dummy_tensor, use_reentrant=True)
        return x

    def forward(self, x, hierarchical=False):
        feats = []
        x = self.stem(x)
        for i in range(4):
            blocks, down = getattr(self, f'enc_block_{i}'), getattr(self, f'down_{i}')
            if self.outside_block_checkpointing:
                x = self.iterative_checkpoint(blocks, x)
                feats.append(x)
                x = checkpoint.checkpoint(down, x, self.dummy_tensor, use_reentrant=True)
            else:
                x = blocks(x)
                feats.append(x)
                x = down(x)
        x = self.iterative_checkpoint(self.bottleneck, x) if self.outside_block_checkpointing else self.bottleneck(x)
        feats.append(x)
        return feats if hierarchical else x
