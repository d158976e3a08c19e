"""AnatoMask module — P/AnatoMask.py:75-219: SparK whose forward returns (inp, rec) patches, plus `forward_loss` and the
teacher-guided `generate_mask`.

`generate_mask` keeps the reference contract bit-exactly: the hard set (top `len_loss` teacher losses) is selected by
the device top-k kernel; the random fill then replays numpy's global-RNG shuffles on the host exactly like the reference
(`mask_rng='numpy'`, the default, drop-in) or uses the device counter-based RNG with no host round trip
(`mask_rng='device'`, what the throughput step uses).
"""
from __future__ import annotations

import numpy as np
import torch

from . import encoder3D, ops
from .spark3D import SparK as _SparKBase


class SparK(_SparKBase):
    mask_rng = 'numpy'            # 'numpy' (reference RNG replay, host sync) | 'device' (no host sync)
    _rng_offset = 0
    rng_counter = None            # optional device uint64 step counter (CUDA-graph replay: the RNG stream must not be baked in)

    @torch.no_grad()
    def generate_mask(self, loss_pred, guide=True, epoch=0, total_epoch=200, generator=None, original_mask=None):
        h, w, d = self.fmap_h, self.fmap_w, self.fmap_d
        B, L = loss_pred.shape
        keep_ratio = float((epoch + 1) / total_epoch) * 0.5 if guide else 2 / 3
        nm = L - self.len_keep
        len_loss = int(nm * keep_ratio)
        dev = loss_pred.device
        if len_loss <= 0:
            if self.mask_rng == 'device':
                type(self)._rng_offset += 1
                off = 0 if self.rng_counter is not None else type(self)._rng_offset * B * L
                _, mk = ops.hard_mask(loss_pred, 0, self.len_keep, seed=0x5EED, offset=off, offset_dev=self.rng_counter)
                mk = mk.bool().view(B, 1, h, w, d)
                return mk, mk
            noise = torch.randn(B, L, device=dev)            # P/AnatoMask.py:99-103 (device RNG in the reference too)
            keep = torch.argsort(noise, dim=1)[:, :self.len_keep]
            mk = torch.zeros(B, L, dtype=torch.bool, device=dev).scatter_(1, keep, True).view(B, 1, h, w, d)
            return mk, mk
        easy_len = nm - len_loss
        if self.mask_rng == 'device':
            type(self)._rng_offset += 1
            off = 0 if self.rng_counter is not None else type(self)._rng_offset * B * L
            _, mk = ops.hard_mask(loss_pred, len_loss, self.len_keep, seed=0x5EED, offset=off,
                                  offset_dev=self.rng_counter)
            mk = mk.bool().view(B, 1, h, w, d)
            return mk, mk
        # parity mode: device top-k, then numpy's shuffles replayed in the reference's order (two per sample)
        _, _, order = ops.hard_mask(loss_pred, len_loss, self.len_keep, want_mask=False, want_order=True)
        order = order.cpu().numpy().astype(np.int64)          # the one host sync the reference semantics force
        mask = np.zeros((B, L), dtype=bool)
        ids2 = np.zeros(L, dtype=np.int64)
        for i in range(B):
            deleted = np.delete(np.arange(L), order[i, L - len_loss:])
            np.random.shuffle(deleted)
            mask[i, deleted[:self.len_keep]] = True
            ids2 = np.zeros(L, dtype=np.int64)                 # re-zeroed per sample like the reference (:116)
            ids2[L - len_loss - easy_len:L - len_loss] = order[i, L - len_loss - easy_len:L - len_loss]
            deleted2 = np.delete(np.arange(L), ids2[L - len_loss - easy_len:L - len_loss])
            np.random.shuffle(deleted2)
            ids2[:L - easy_len] = deleted2
        # easy_mask with the reference's quirk (P/AnatoMask.py:116): `ids_shuffle2` is re-created inside the loop, so only
        # the LAST sample's row is filled and all other rows argsort a constant.  No caller uses it
        # (P/pretrain_AntoMask.py:427-430); ties are resolved with a stable sort here.
        ids2_all = np.zeros((B, L), dtype=np.int64)
        ids2_all[B - 1] = ids2
        restore2 = np.argsort(ids2_all, axis=1, kind='stable')
        easy_row = np.zeros((B, L), dtype=bool)
        easy_row[:, :self.len_keep + len_loss] = True
        easy = np.take_along_axis(easy_row, restore2, axis=1)
        mk = torch.from_numpy(mask).to(dev).view(B, 1, h, w, d)
        ek = torch.from_numpy(easy).to(dev).view(B, 1, h, w, d)
        return mk, ek

    def forward(self, inp_bchwd: torch.Tensor, active_b1ff=None, vis=False, return_feat=False):
        if active_b1ff is None:
            active_b1ff = self.mask(inp_bchwd.shape[0], inp_bchwd.device)
        if return_feat:
            raise NotImplementedError('return_feat=True is not used by any shipped script')
        rec_bchwd = self.reconstruct(inp_bchwd, active_b1ff)
        if vis:
            return self._visualise(inp_bchwd, rec_bchwd, active_b1ff)
        # (B, L, p³) patches exactly like the reference; forward_loss / teacher_loss also take the volumes directly
        return self.patchify(inp_bchwd), self.patchify(rec_bchwd)

    def forward_volumes(self, inp_bchwd, active_b1ff):
        """Fast path of the same forward: (inp, rec) as (B,1,D,H,W) volumes, no patchified copies."""
        return inp_bchwd, self.reconstruct(inp_bchwd, active_b1ff)

    def _as_volume(self, t):
        return self.unpatchify(t) if t.dim() == 3 else t

    def forward_loss(self, inp, rec, active_b1ff):
        """P/AnatoMask.py:190-202 → (recon_loss, rec_loss (B,L)); one fused kernel."""
        inp, rec = self._as_volume(inp), self._as_volume(rec)
        active = active_b1ff[:, 0].to(torch.uint8).contiguous()
        return ops.PatchLossFn.apply(inp, rec, active, True)

    @torch.no_grad()
    def teacher_loss(self, inp, rec, active_b1ff):
        """P/pretrain_AntoMask.py:423-425: raw per-patch MSE × non-active, (B, L)."""
        inp, rec = self._as_volume(inp), self._as_volume(rec)
        active = active_b1ff[:, 0].to(torch.uint8).contiguous()
        return ops.PatchLossFn.apply(inp, rec, active, False)[1]

    def forward_learning_loss(self, loss_pred, loss_target):
        mean = loss_target.mean(dim=1, keepdim=True)
        var = loss_target.var(dim=1, keepdim=True)
        loss_target = (loss_target - mean) / (var + 1.e-6) ** .5
        return ((loss_pred - loss_target) ** 2).mean()
