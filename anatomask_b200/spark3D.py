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


def _densify_norm(kind: str, width: int, sbn: bool) -> nn.Module:
    """densify_norm option → sparse norm layer (P/spark3D.py:61-72)."""
    if kind == 'bn':
        return (encoder3D.SparseSyncBatchNorm3d if sbn else encoder3D.SparseBatchNorm3d)(width)
    if kind == 'ln':
        return encoder3D.SparseConvNeXtLayerNorm(width, data_format='channels_first', sparse=True)
    if kind == 'gn':
        return encoder3D.SparseGroupNorm(width, width, sparse=True)
    if kind == 'in':
        return encoder3D.SparseInstanceNorm(width, sparse=True)
    return nn.Identity()


class SparK(nn.Module):
    def __init__(self, sparse_encoder: encoder3D.SparseEncoder, dense_decoder: LightDecoder, mask_ratio=0.6,
                 densify_norm='in', sbn=False):
        super().__init__()
        r = sparse_encoder.downsample_ratio
        self.downsample_ratio = r
        self.fmap_h, self.fmap_w, self.fmap_d = (s // r for s in sparse_encoder.input_size)
        self.mask_ratio = mask_ratio
        self.len_keep = round(self.fmap_h * self.fmap_w * self.fmap_d * (1 - mask_ratio))
        self.sparse_encoder, self.dense_decoder = sparse_encoder, dense_decoder
        self.sbn = sbn
        self.densify_norm_str = densify_norm.lower()
        enc_widths = list(sparse_encoder.enc_feat_map_chs)[::-1]           # deepest feature map first
        self.hierarchy = len(enc_widths)
        dec_widths = [dense_decoder.width >> i for i in range(self.hierarchy)]
        tokens, norms, projs = [], [], []
        for level, (e_w, d_w) in enumerate(zip(enc_widths, dec_widths)):
            tok = nn.Parameter(torch.zeros(1, e_w, 1, 1, 1))
            nn.init.trunc_normal_(tok, mean=0, std=.02, a=-.02, b=.02)
            tokens.append(tok)
            norms.append(_densify_norm(self.densify_norm_str, e_w, sbn))
            if level == 0 and e_w == d_w:
                projs.append(nn.Identity())                                 # STUNet: decoder width == deepest encoder width
            else:
                ks = 3 if level > 0 else 1
                projs.append(nn.Conv3d(e_w, d_w, kernel_size=ks, stride=1, padding=ks // 2, bias=True))
        self.densify_norms = nn.ModuleList(norms)
        self.densify_projs = nn.ModuleList(projs)
        self.mask_tokens = nn.ParameterList(tokens)

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
        # backward reaches 'densify_done' once every live densify level has produced its gradient w.r.t. the encoder
        # features, and 'decoder_done' once dec.0 — the last decoder block of the backward pass — has (ops.backward_mark)
        fea_bcffs[:n_live] = ops.backward_mark('densify_done', *fea_bcffs[:n_live])
        to_dec = [self._densify_level(i, f, m) if i < n_live else None for i, f in enumerate(fea_bcffs)]
        to_dec[0], = ops.backward_mark('decoder_done', to_dec[0])
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
        """(B,C,H,W,D) → (B, L, p³·C): patch index l = (h·fw + w)·fd + d, element index ((p·r + q)·r + g)·C + c."""
        r = self.downsample_ratio
        fh, fw, fd = self.fmap_h, self.fmap_w, self.fmap_d
        B, C = bchwd.shape[:2]
        x = bchwd.reshape(B, C, fh, r, fw, r, fd, r).permute(0, 2, 4, 6, 3, 5, 7, 1)
        return x.reshape(B, fh * fw * fd, r ** 3 * C)

    def unpatchify(self, bln):
        r = self.downsample_ratio
        fh, fw, fd = self.fmap_h, self.fmap_w, self.fmap_d
        B, C = bln.shape[0], bln.shape[-1] // r ** 3
        x = bln.reshape(B, fh, fw, fd, r, r, r, C).permute(0, 7, 1, 4, 2, 5, 3, 6)
        return x.reshape(B, C, fh * r, fw * r, fd * r)

    def __repr__(self):
        body = super().__repr__().replace(type(self).__name__, '', 1)
        return f'\n[SparK.config]: {pformat(self.get_config(), indent=2, width=250)}\n[SparK.structure]: {body}'

    def get_config(self):
        return {'mask_ratio': self.mask_ratio, 'densify_norm_str': self.densify_norm_str, 'sbn': self.sbn,
                'hierarchy': self.hierarchy, 'sparse_encoder.input_size': self.sparse_encoder.input_size,
                'dense_decoder.width': self.dense_decoder.width}

    def state_dict(self, destination=None, prefix='', keep_vars=False, with_config=False):
        sd = super().state_dict(destination=destination, prefix=prefix, keep_vars=keep_vars)
        if with_config:
            sd['config'] = self.get_config()
        return sd

    def load_state_dict(self, state_dict, strict=True):
        saved_cfg = state_dict.pop('config', None)
        result = super().load_state_dict(state_dict, strict=strict)
        mismatches = [] if saved_cfg is None else \
            [(k, v, saved_cfg.get(k)) for k, v in self.get_config().items() if saved_cfg.get(k) != v]
        for k, mine, theirs in mismatches:                     # same message / exception type as P/spark3D.py:199-201
            msg = f'[SparseMIM.load_state_dict] config mismatch:  this.{k}={mine} (ckpt.{k}={theirs})'
            if strict:
                raise AttributeError(msg)
            print(msg, file=sys.stderr)
        return result
