"""SparK pre-training module — P/spark3D.py:30-204.  Same constructor, attributes, parameter names and
`forward(inp_bchwd, active_b1ff=None, vis=False) -> loss` contract; the arithmetic runs in sm_100a kernels.

Differences a caller can observe: none in results (within bf16 tolerance); the dead finest densify level
(`densify_*[4]`, `mask_tokens[4]`, never consumed by the 4-block decoder — P/decoder3D.py:57-60) is not computed, so its
parameters keep `.grad = None` exactly as in the reference.
"""
from __future__ import annotations

import sys
from pprint import pformat
from typing import List

import torch
import torch.nn as nn

from . import encoder3D, ops
from .decoder3D import LightDecoder


class SparK(nn.Module):
    def __init__(self, sparse_encoder: encoder3D.SparseEncoder, dense_decoder: LightDecoder, mask_ratio=0.6,
                 densify_norm='in', sbn=False):
        super().__init__()
        input_size, downsample_ratio = sparse_encoder.input_size, sparse_encoder.downsample_ratio
        self.downsample_ratio = downsample_ratio
        self.fmap_h, self.fmap_w, self.fmap_d = (input_size[0] // downsample_ratio, input_size[1] // downsample_ratio,
                                                 input_size[2] // downsample_ratio)
        self.mask_ratio = mask_ratio
        self.len_keep = round(self.fmap_h * self.fmap_w * self.fmap_d * (1 - mask_ratio))
        self.sparse_encoder = sparse_encoder
        self.dense_decoder = dense_decoder
        self.sbn = sbn
        self.hierarchy = len(sparse_encoder.enc_feat_map_chs)
        self.densify_norm_str = densify_norm.lower()
        self.densify_norms = nn.ModuleList()
        self.densify_projs = nn.ModuleList()
        self.mask_tokens = nn.ParameterList()
        e_widths, d_width = list(self.sparse_encoder.enc_feat_map_chs), self.dense_decoder.width
        for i in range(self.hierarchy):          # from the smallest feature map to the largest
            e_width = e_widths.pop()
            p = nn.Parameter(torch.zeros(1, e_width, 1, 1, 1))
            nn.init.trunc_normal_(p, mean=0, std=.02, a=-.02, b=.02)
            self.mask_tokens.append(p)
            if self.densify_norm_str == 'bn':
                norm = (encoder3D.SparseSyncBatchNorm3d if self.sbn else encoder3D.SparseBatchNorm3d)(e_width)
            elif self.densify_norm_str == 'ln':
                norm = encoder3D.SparseConvNeXtLayerNorm(e_width, data_format='channels_first', sparse=True)
            elif self.densify_norm_str == 'gn':
                norm = encoder3D.SparseGroupNorm(e_width, e_width, sparse=True)
            elif self.densify_norm_str == 'in':
                norm = encoder3D.SparseInstanceNorm(e_width, sparse=True)
            else:
                norm = nn.Identity()
            self.densify_norms.append(norm)
            if i == 0 and e_width == d_width:
                proj = nn.Identity()
            else:
                ks = 1 if i <= 0 else 3
                proj = nn.Conv3d(e_width, d_width, kernel_size=ks, stride=1, padding=ks // 2, bias=True)
            self.densify_projs.append(proj)
            d_width //= 2

    # ---- masking -------------------------------------------------------------------------------------------
    def mask(self, B: int, device, generator=None):
        """P/spark3D.py:92-96 — CPU default generator, so seeding behaves exactly like the reference."""
        h, w, d = self.fmap_h, self.fmap_w, self.fmap_d
        idx = torch.rand(B, h * w * d, generator=generator).argsort(dim=1)
        idx = idx[:, :self.len_keep].to(device)
        return torch.zeros(B, h * w * d, dtype=torch.bool, device=device).scatter_(dim=1, index=idx, value=True) \
            .view(B, 1, h, w, d)

    # ---- encode → densify → decode ---------------------------------------------------------------------------
    def _densify_level(self, i: int, fea: torch.Tensor, m: ops.MaskCtx) -> torch.Tensor:
        norm, proj, token = self.densify_norms[i], self.densify_projs[i], self.mask_tokens[i]
        xi = ops.to_internal(fea)
        if isinstance(norm, encoder3D.SparseInstanceNorm):
            y = ops.densify_norm_fill(xi, norm.weight, norm.bias, token, norm.eps, m)
        elif isinstance(norm, (encoder3D.SparseBatchNorm3d, encoder3D.SparseSyncBatchNorm3d)):
            if norm.training:
                group = None
                if isinstance(norm, encoder3D.SparseSyncBatchNorm3d):
                    import torch.distributed as dist
                    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                        group = dist.group.WORLD
                y = ops.densify_norm_fill(xi, norm.weight, norm.bias, token, norm.eps, m,
                                          (norm.running_mean, norm.running_var, norm.num_batches_tracked),
                                          0.1 if norm.momentum is None else norm.momentum, group)
            else:
                y = ops.norm_eval(xi, norm.weight, norm.bias, norm.running_mean, norm.running_var, norm.eps, 0, m, token)
        else:
            raise NotImplementedError(f"densify_norm='{self.densify_norm_str}' has no sm_100a kernel "
                                      "(every shipped script uses 'in')")
        if isinstance(proj, nn.Conv3d):
            y = ops.conv3d(y, proj.weight, proj.bias, proj.kernel_size[0], 1)
        return ops.to_external(y)

    def reconstruct(self, inp_bchwd: torch.Tensor, active_b1ff: torch.Tensor) -> torch.Tensor:
        """steps 1-4 of P/spark3D.py:98-127: returns rec (B,1,D,H,W) fp32."""
        ops.require_cuda(inp_bchwd)
        encoder3D._cur_active = active_b1ff
        m = encoder3D._mask_ctx()
        m.frac_hint = self.len_keep / float(self.fmap_h * self.fmap_w * self.fmap_d)
        # the stem kernel applies the visibility mask to the raw input itself (P/spark3D.py:104-107 folded in)
        fea_bcffs: List[torch.Tensor] = self.sparse_encoder(inp_bchwd)
        fea_bcffs = list(reversed(fea_bcffs))
        n_live = len(self.dense_decoder.dec)
        to_dec = [self._densify_level(i, f, m) if i < n_live else None for i, f in enumerate(fea_bcffs)]
        return self.dense_decoder(to_dec)

    def forward(self, inp_bchwd: torch.Tensor, active_b1ff=None, vis=False):
        if active_b1ff is None:
            active_b1ff = self.mask(inp_bchwd.shape[0], inp_bchwd.device)
        rec_bchwd = self.reconstruct(inp_bchwd, active_b1ff)
        if vis:
            return self._visualise(inp_bchwd, rec_bchwd, active_b1ff)
        m = encoder3D._mask_ctx()
        loss, _ = ops.PatchLossFn.apply(inp_bchwd, rec_bchwd, m.active, True)
        return loss

    def _visualise(self, inp_bchwd, rec_bchwd, active_b1ff):
        """P/spark3D.py:140-144 — host-side convenience path (plain torch, not on the training step)."""
        r = self.downsample_ratio
        active = active_b1ff.repeat_interleave(r, 2).repeat_interleave(r, 3).repeat_interleave(r, 4)
        inp, rec = self.patchify(inp_bchwd.float()), self.patchify(rec_bchwd.float())
        mean = inp.mean(dim=-1, keepdim=True)
        var = (inp.var(dim=-1, keepdim=True) + 1e-6) ** .5
        rec_bchwd = self.unpatchify(rec * var + mean)
        return inp_bchwd, inp_bchwd * active, torch.where(active, inp_bchwd, rec_bchwd)

    def patchify(self, bchwd):
        p = self.downsample_ratio
        h, w, d = self.fmap_h, self.fmap_w, self.fmap_d
        B, C = bchwd.shape[:2]
        bchwd = bchwd.reshape(shape=(B, C, h, p, w, p, d, p))
        bchwd = torch.einsum('bchpwqdg->bhwdpqgc', bchwd)
        return bchwd.reshape(shape=(B, h * w * d, C * p ** 3))

    def unpatchify(self, bln):
        p = self.downsample_ratio
        h, w, d = self.fmap_h, self.fmap_w, self.fmap_d
        B, C = bln.shape[0], bln.shape[-1] // p ** 3
        bln = bln.reshape(shape=(B, h, w, d, p, p, p, C))
        bln = torch.einsum('bhwdpqgc->bchpwqdg', bln)
        return bln.reshape(shape=(B, C, h * p, w * p, d * p))

    def __repr__(self):
        return (f'\n[SparK.config]: {pformat(self.get_config(), indent=2, width=250)}\n'
                f'[SparK.structure]: {super(SparK, self).__repr__().replace(SparK.__name__, "")}')

    def get_config(self):
        return {'mask_ratio': self.mask_ratio, 'densify_norm_str': self.densify_norm_str, 'sbn': self.sbn,
                'hierarchy': self.hierarchy, 'sparse_encoder.input_size': self.sparse_encoder.input_size,
                'dense_decoder.width': self.dense_decoder.width}

    def state_dict(self, destination=None, prefix='', keep_vars=False, with_config=False):
        state = super(SparK, self).state_dict(destination=destination, prefix=prefix, keep_vars=keep_vars)
        if with_config:
            state['config'] = self.get_config()
        return state

    def load_state_dict(self, state_dict, strict=True):
        config: dict = state_dict.pop('config', None)
        incompatible_keys = super(SparK, self).load_state_dict(state_dict, strict=strict)
        if config is not None:
            for k, v in self.get_config().items():
                ckpt_v = config.get(k, None)
                if ckpt_v != v:
                    err = f'[SparseMIM.load_state_dict] config mismatch:  this.{k}={v} (ckpt.{k}={ckpt_v})'
                    if strict:
                        raise AttributeError(err)
                    print(err, file=sys.stderr)
        return incompatible_keys
