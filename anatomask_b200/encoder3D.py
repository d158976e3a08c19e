"""Sparse layer API of the reference (P/encoder3D.py) backed by sm_100a kernels.

Same names, constructor signatures, parameter/buffer names and error behaviour as the reference classes; the arithmetic
goes through anatomask_b200.ops (C ABI → CUDA).  Tensors crossing a module boundary are logical (N, C, D, H, W);
channels-last bf16 tensors pass through zero-copy, anything else is converted once by a layout kernel.

`_cur_active` is kept as the settable module global the reference uses (P/encoder3D.py:5, set at P/spark3D.py:103):
assign the (B,1,f,f,f) bool mask and every sparse layer picks it up.  The device work-list derived from it is cached
per mask tensor, so it is built once per forward instead of `nonzero()`-ing 30 times (P/encoder3D.py:7-10).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from ._lib import ACT_NONE

_cur_active: torch.Tensor = None            # B1fff bool, True = visible
_ctx_cache = (None, -1, None)


def _mask_ctx() -> ops.MaskCtx:
    global _ctx_cache
    if _cur_active is None:
        raise RuntimeError('encoder3D._cur_active is not set (SparK.forward sets it; P/spark3D.py:103)')
    if isinstance(_cur_active, ops.MaskCtx):
        return _cur_active
    # keyed on the tensor AND its version counter: a static mask buffer updated in place (the usual pattern next to CUDA
    # graphs) must rebuild the uint8 copy and the device work-list, like the reference recomputes from _cur_active per call
    if _ctx_cache[0] is not _cur_active or _ctx_cache[1] != _cur_active._version:
        _ctx_cache = (_cur_active, _cur_active._version, ops.MaskCtx(_cur_active))
    return _ctx_cache[2]


def _single(v):
    return v[0] if isinstance(v, (tuple, list)) else v


def _uniform(v, what):
    if isinstance(v, (tuple, list)):
        if len(set(v)) != 1:
            raise NotImplementedError(f'anisotropic {what}={v} is not supported by the sm_100a kernels')
        return v[0]
    return v


class SparseConv3d(nn.Conv3d):
    """(conv3d(x, W) + b) · mask at the output resolution — P/encoder3D.py:12-15,27-28."""

    def forward(self, x: torch.Tensor):
        k, s, p = _uniform(self.kernel_size, 'kernel_size'), _uniform(self.stride, 'stride'), _uniform(self.padding, 'padding')
        depthwise = self.groups > 1 and self.groups == self.in_channels == self.out_channels
        ok_k = (3, 5, 7) if depthwise else (1, 3)
        if k not in ok_k or s not in (1, 2) or p != k // 2 or _uniform(self.dilation, 'dilation') != 1 \
                or (self.groups != 1 and not depthwise) or self.padding_mode != 'zeros':
            raise NotImplementedError(f'SparseConv3d(k={k}, s={s}, p={p}, groups={self.groups}) has no sm_100a kernel yet')
        m = _mask_ctx()
        if depthwise:       # ConvNeXt / MedNeXt blocks: k³ depthwise, P/encoder3D.py:259, P/MedNeXt_head.py:255-262
            return ops.to_external(ops.depthwise_conv3d(ops.to_internal(x), self.weight, self.bias, k, s, m))
        if self.in_channels == 1:
            if k == 3 and s == 1:
                w3 = torch.zeros(self.out_channels, 1, 1, 1, 1, device=x.device)
                b = self.bias if self.bias is not None else torch.zeros(self.out_channels, device=x.device)
                y, _ = ops.StemFn.apply(x, self.weight, b, w3, torch.zeros_like(b), m)
                return ops.to_external(y)
            if k == 1 and s == 1:
                w1 = torch.zeros(self.out_channels, 1, 3, 3, 3, device=x.device)
                b = self.bias if self.bias is not None else torch.zeros(self.out_channels, device=x.device)
                _, y = ops.StemFn.apply(x, w1, torch.zeros_like(b), self.weight, b, m)
                return ops.to_external(y)
            raise NotImplementedError('SparseConv3d with in_channels=1 supports stride 1 only')
        y = ops.conv3d(ops.to_internal(x), self.weight, self.bias, k, s, m)
        return ops.to_external(y)


def _pool_args(self, what):
    k = _uniform(self.kernel_size, 'kernel_size')
    s = k if self.stride is None else _uniform(self.stride, 'stride')
    p = _uniform(self.padding, 'padding')
    if self.ceil_mode or getattr(self, 'return_indices', False) or _uniform(getattr(self, 'dilation', 1), 'dilation') != 1:
        raise NotImplementedError(f'{what}: ceil_mode / return_indices / dilation have no sm_100a kernel')
    return k, s, p


class SparseMaxPooling(nn.MaxPool3d):
    """max_pool3d(x) · mask at the output resolution — P/encoder3D.py:12-15,31-32."""

    def forward(self, x):
        k, s, p = _pool_args(self, 'SparseMaxPooling')
        return ops.to_external(ops.pool3d(ops.to_internal(x), k, s, p, 0, _mask_ctx()))


class SparseAvgPooling(nn.AvgPool3d):
    """avg_pool3d(x) · mask at the output resolution — P/encoder3D.py:12-15,35-36."""

    def forward(self, x):
        k, s, p = _pool_args(self, 'SparseAvgPooling')
        return ops.to_external(ops.pool3d(ops.to_internal(x), k, s, p, 1, _mask_ctx(), self.count_include_pad,
                                          self.divisor_override))


def _sp_bn_forward(self, x: torch.Tensor, group=None):
    """P/encoder3D.py:17-25: BatchNorm1d over the visible voxels (pooled over the local batch, or all ranks when Sync)."""
    m = _mask_ctx()
    xi = ops.to_internal(x)
    if self.training or not self.track_running_stats:
        running, mom = None, 0.0
        if self.track_running_stats and self.training:
            running = (self.running_mean, self.running_var, self.num_batches_tracked)
            mom = 0.1 if self.momentum is None else self.momentum
        y = ops.masked_norm(xi, self.weight, self.bias, self.eps, m, ACT_NONE, None, running, mom, group)
    else:
        y = ops.norm_eval(xi, self.weight, self.bias, self.running_mean, self.running_var, self.eps, ACT_NONE, m)
    return ops.to_external(y)


class SparseBatchNorm3d(nn.BatchNorm1d):
    forward = _sp_bn_forward


class SparseSyncBatchNorm3d(nn.SyncBatchNorm):
    def forward(self, x):
        import torch.distributed as dist
        group = None
        if self.training and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            group = self.process_group if self.process_group is not None else dist.group.WORLD
        return _sp_bn_forward(self, x, group)


class SparseInstanceNorm(nn.InstanceNorm1d):
    """NOT an instance norm in the reference: the visible voxels of the whole local batch are fed to InstanceNorm1d as
    one unbatched (C, N_active) sample → per-channel statistics pooled over the batch, biased variance, no running
    stats, same in train and eval (P/encoder3D.py:149-158)."""

    def __init__(self, num_features, eps=1e-6, sparse=True):
        super().__init__(num_features, eps, affine=True)
        self.sparse = sparse

    def forward(self, x):
        if x.ndim == 5:
            if self.sparse:
                y = ops.masked_norm(ops.to_internal(x), self.weight, self.bias, self.eps, _mask_ctx())
                return ops.to_external(y)
            return super().forward(x)
        if self.sparse:
            raise NotImplementedError
        return super().forward(x)

    def __repr__(self):
        return super().__repr__()[:-1] + f', sp={self.sparse})'


class SparseGroupNorm(nn.GroupNorm):
    """nn.GroupNorm over the (N_active, C) matrix of visible voxels: every visible voxel is normalised on its own over each
    channel group, masked voxels stay zero (P/encoder3D.py:47-78)."""

    def __init__(self, num_groups, num_channels, eps=1e-6, sparse=True):
        super().__init__(num_groups, num_channels, eps)
        self.sparse = sparse

    def forward(self, x):
        if x.ndim == 5:
            if self.sparse:
                y = ops.voxel_norm(ops.to_internal(x), self.weight, self.bias, self.num_groups, self.eps, _mask_ctx())
                return ops.to_external(y)
            return super().forward(x)
        if self.sparse:
            raise NotImplementedError
        return super().forward(x)

    def __repr__(self):
        return super().__repr__()[:-1] + f', sp={self.sparse})'


class SparseGRN(nn.Module):
    """Declared by the reference (P/encoder3D.py:102-137) but instantiated by none of its heads (the MedNeXt block's GRN is
    commented out, P/MedNeXt_head.py:276,292): constructor and parameters kept, no kernel."""

    def __init__(self, dim, use_bias=True, sparse=True):
        super().__init__()
        self.use_bias, self.sparse = use_bias, sparse
        self.gamma = nn.Parameter(torch.zeros(1, dim))
        if use_bias:
            self.beta = nn.Parameter(torch.zeros(1, dim))

    def forward(self, x):
        raise NotImplementedError('SparseGRN is not used by any head of the reference; no sm_100a kernel')


class SparseAdaptiveAvgPooling(nn.AdaptiveAvgPool3d):
    """Mean over the visible voxels of each sample: Σ x·mask / (Σ mask + 1e-6) → (B, C, 1, 1, 1) (P/encoder3D.py:181-190;
    the reference ignores output_size in its forward, so does this)."""

    def __init__(self, output_size, sparse=True):
        super().__init__(output_size)
        self.output_size, self.sparse = output_size, sparse

    def forward(self, x):
        mean = ops.MaskedMeanFn.apply(ops.to_internal(x), _mask_ctx())
        return mean.to(x.dtype).view(x.shape[0], x.shape[1], 1, 1, 1)


class SparseConvNeXtLayerNorm(nn.LayerNorm):
    """LayerNorm over the channel axis of every (visible) voxel, channels_last (B,H,W,D,C) or channels_first (B,C,H,W,D) —
    P/encoder3D.py:193-243.  The dense channels_first branch of the reference is the same per-voxel formula and runs on
    the same kernel; dense channels_last is nn.LayerNorm itself."""

    def __init__(self, normalized_shape, eps=1e-6, data_format='channels_last', sparse=True):
        if data_format not in ['channels_last', 'channels_first']:
            raise NotImplementedError
        super().__init__(normalized_shape, eps, elementwise_affine=True)
        self.data_format, self.sparse = data_format, sparse

    def forward(self, x):
        if x.ndim == 5:
            if self.data_format == 'channels_last':
                if not self.sparse:
                    return super().forward(x)
                ops.require_cuda(x)
                return ops.voxel_norm(x.to(ops.bf16).contiguous(), self.weight, self.bias, 1, self.eps, _mask_ctx())
            y = ops.voxel_norm(ops.to_internal(x), self.weight, self.bias, 1, self.eps, _mask_ctx() if self.sparse else None)
            return ops.to_external(y)
        if self.sparse:
            raise NotImplementedError
        return super().forward(x)

    def __repr__(self):
        return super().__repr__()[:-1] + f', ch={self.data_format.split("_")[-1]}, sp={self.sparse})'


class _DropPath(nn.Module):
    """timm.models.layers.DropPath (stochastic depth per sample; timm is not in the reference tree): identity in eval."""

    def __init__(self, p):
        super().__init__()
        self.drop_prob = p

    def forward(self, x):
        if self.drop_prob == 0. or not self.training:
            return x
        keep = 1 - self.drop_prob
        r = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x * (r / keep)


class SparseConvNeXtBlock(nn.Module):
    """DwConv k³ → LayerNorm → Linear(C, 4C) → GELU → Linear(4C, C) → layer scale γ → · mask → + input
    (P/encoder3D.py:246-282).  Same parameter names as the reference (`dwconv`, `norm`, `pwconv1`, `pwconv2`, `gamma`).
    The two Linears run as 1×1×1 implicit GEMMs on the tensor cores; inside a sparse block they visit visible patches only
    (the reference computes them on masked voxels too and multiplies by the mask afterwards — same result)."""

    def __init__(self, dim, drop_path=0., layer_scale_init_value=1e-6, sparse=True, ks=7):
        super().__init__()
        self.dwconv = nn.Conv3d(dim, dim, kernel_size=ks, padding=ks // 2, groups=dim)
        self.norm = SparseConvNeXtLayerNorm(dim, eps=1e-6, sparse=sparse)
        self.pwconv1 = nn.Linear(dim, 4 * dim)
        self.act = nn.GELU()
        self.pwconv2 = nn.Linear(4 * dim, dim)
        self.gamma = nn.Parameter(layer_scale_init_value * torch.ones((dim)), requires_grad=True) \
            if layer_scale_init_value > 0 else None
        self.drop_path = _DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.sparse = sparse

    def forward(self, x):
        m = _mask_ctx() if self.sparse else None
        xi = ops.to_internal(x)
        dw = self.dwconv
        k = _uniform(dw.kernel_size, 'kernel_size')
        h = ops.depthwise_conv3d(xi, dw.weight, dw.bias, k, 1, _mask_ctx() if isinstance(dw, SparseConv3d) else None)
        h = self.norm(h)                                                  # channels_last: already the internal layout
        h = ops.conv3d(h, self.pwconv1.weight, self.pwconv1.bias, 1, 1, m)
        h = ops.gelu(h)
        h = ops.conv3d(h, self.pwconv2.weight, self.pwconv2.bias, 1, 1, m)
        if not isinstance(self.drop_path, nn.Identity):
            h = self.drop_path(h)
        return ops.to_external(ops.LayerScaleFn.apply(xi, h, self.gamma, m))

    def __repr__(self):
        return super().__repr__()[:-1] + f', sp={self.sparse})'


class SparseEncoder(nn.Module):
    """Wraps any CNN following the `get_downsample_ratio / get_feature_map_channels / forward(x, hierarchical)` protocol
    and swaps its dense layers for the sparse ones — P/encoder3D.py:293-367."""

    def __init__(self, cnn, input_size, sbn=False, verbose=False):
        super().__init__()
        self.sp_cnn = SparseEncoder.dense_model_to_sparse(m=cnn, verbose=verbose, sbn=sbn)
        self.input_size, self.downsample_ratio, self.enc_feat_map_chs = \
            input_size, cnn.get_downsample_ratio(), cnn.get_feature_map_channels()

    @staticmethod
    def dense_model_to_sparse(m: nn.Module, verbose=False, sbn=False):
        """Recursive dense → sparse swap with the rules (and the errors) of P/encoder3D.py:298-364, written as a table:
        (dense type, sparse type, constructor arguments taken from the dense layer, state to carry over)."""
        bn_cls = SparseSyncBatchNorm3d if sbn else SparseBatchNorm3d
        rules = (
            (nn.Conv3d, SparseConv3d,
             lambda d: dict(in_channels=d.in_channels, out_channels=d.out_channels, kernel_size=d.kernel_size, stride=d.stride,
                            padding=d.padding, dilation=d.dilation, groups=d.groups, bias=d.bias is not None,
                            padding_mode=d.padding_mode), ('weight', 'bias')),
            (nn.MaxPool3d, SparseMaxPooling,
             lambda d: dict(kernel_size=d.kernel_size, stride=d.stride, padding=d.padding, dilation=d.dilation,
                            return_indices=d.return_indices, ceil_mode=d.ceil_mode), ()),
            (nn.AvgPool3d, SparseAvgPooling,
             lambda d: dict(kernel_size=d.kernel_size, stride=d.stride, padding=d.padding, ceil_mode=d.ceil_mode,
                            count_include_pad=d.count_include_pad, divisor_override=d.divisor_override), ()),
            (nn.GroupNorm, SparseGroupNorm, lambda d: dict(num_groups=d.num_groups, num_channels=d.num_channels, eps=d.eps), ()),
            (nn.AdaptiveAvgPool3d, SparseAdaptiveAvgPooling, lambda d: dict(output_size=(1, 1, 1)), ()),
            (nn.InstanceNorm3d, SparseInstanceNorm, lambda d: dict(num_features=d.num_features, eps=d.eps), ('weight', 'bias')),
            ((nn.BatchNorm3d, nn.SyncBatchNorm), bn_cls,
             lambda d: dict(num_features=d.weight.shape[0], eps=d.eps, momentum=d.momentum, affine=d.affine,
                            track_running_stats=d.track_running_stats),
             ('weight', 'bias', 'running_mean', 'running_var', 'num_batches_tracked')),
            (nn.LayerNorm, SparseConvNeXtLayerNorm, lambda d: dict(normalized_shape=d.weight.shape[0], eps=d.eps), ('weight', 'bias')),
        )
        already_sparse = (SparseConv3d, SparseGroupNorm, SparseSyncBatchNorm3d, SparseConvNeXtLayerNorm)
        if isinstance(m, nn.Conv1d):
            raise NotImplementedError
        out = m
        skip = isinstance(m, nn.Conv3d) and getattr(m, 'skip_sparse_conversion', False)      # honoured for convs only
        if not isinstance(m, already_sparse) and not skip:
            for dense_t, sparse_t, ctor_args, carried in rules:
                if isinstance(m, dense_t):
                    out = sparse_t(**ctor_args(m))
                    for attr in carried:
                        src = getattr(m, attr, None)
                        if src is not None:
                            getattr(out, attr).data.copy_(src.data)
                    if hasattr(m, 'qconfig') and sparse_t is bn_cls:
                        out.qconfig = m.qconfig
                    break
        for name, child in m.named_children():
            out.add_module(name, SparseEncoder.dense_model_to_sparse(child, verbose=verbose, sbn=sbn))
        return out

    def forward(self, x):
        return self.sp_cnn(x, hierarchical=True)
