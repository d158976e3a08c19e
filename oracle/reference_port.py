"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product package `anatomask_b200`.

A CPU restatement (plain PyTorch fp32, functional style over a flat state-dict) of the reference's
SparK / AnatoMask pre-training step.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import this file, and only as the checker / the CPU baseline.

Pinning: the reference has NO tests or golden vectors for this path, so this port is pinned
against outputs of the UNMODIFIED reference modules executed in the build container by
`oracle/make_golden.py` (fixtures under `tests/golden/`, checked by `tests/test_oracle_golden.py`).
The EMA arithmetic follows `timm.utils.ModelEma` (third-party, unpinned, absent from /root/reference) — that one
piece is "parity unpinned" by any reference artefact; it is restated from timm's published semantics.

Reference files followed (P = nnunetv2/training/nnUNetTrainer/variants/pretrain):
  P/encoder3D.py:7-44,138-169   masked conv / pooled masked norm
  P/STUNet_head.py:8-103        encoder stages and BasicResBlock
  P/decoder3D.py:13-84          UNetBlock / LightDecoder
  P/spark3D.py:30-164           densify, patchify, loss
  P/AnatoMask.py:75-202         generate_mask, forward_loss
  P/pretrain.py:388-411, P/pretrain_AntoMask.py:383-441   step glue
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------------------------
# configuration + parameter inventory (checkpoint-key contract)
# ----------------------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Cfg:
    base: int = 32                      # STUNet-S 16, -B 32, -L 64, -H 96  (STUNetTrainer.py:215-283)
    depth: int = 1                      # blocks per stage: S/B 1, L 2, H 3
    input_size: Tuple[int, int, int] = (128, 128, 128)
    mask_ratio: float = 0.6
    in_ch: int = 1
    ratio: int = 16                     # P/STUNet_head.py:56

    @property
    def dims(self) -> List[int]:
        return [self.base * m for m in (1, 2, 4, 8, 16)]

    @property
    def fmap(self) -> Tuple[int, int, int]:
        return tuple(s // self.ratio for s in self.input_size)

    @property
    def L(self) -> int:
        f = self.fmap
        return f[0] * f[1] * f[2]

    @property
    def len_keep(self) -> int:          # P/spark3D.py:39
        return round(self.L * (1 - self.mask_ratio))

    @property
    def width(self) -> int:             # decoder width == deepest encoder width (P/pretrain.py:208)
        return self.dims[4]


CONFIGS = {
    'tiny': Cfg(base=8, depth=1, input_size=(32, 32, 32)),
    'S64': Cfg(base=16, depth=1, input_size=(64, 64, 64)),
    'B64': Cfg(base=32, depth=1, input_size=(64, 64, 64)),
    'B128': Cfg(base=32, depth=1, input_size=(128, 128, 128)),
    'L64': Cfg(base=64, depth=2, input_size=(64, 64, 64)),
    'L128': Cfg(base=64, depth=2, input_size=(128, 128, 128)),
    # non-cubic volume with odd mask-grid extents (the authors train at 112x112x128 -> fmap 7x7x8)
    'S_aniso': Cfg(base=16, depth=1, input_size=(48, 80, 32)),
    'L32': Cfg(base=64, depth=2, input_size=(32, 32, 32)),
}


def param_shapes(cfg: Cfg) -> Dict[str, Tuple[Tuple[int, ...], str]]:
    """name -> (shape, kind) for every state_dict entry of reference SparK(SparseEncoder(STUNet), LightDecoder)."""
    out: Dict[str, Tuple[Tuple[int, ...], str]] = {}
    dims = cfg.dims
    for s in range(5):
        for b in range(cfg.depth):
            cin = (cfg.in_ch if s == 0 else dims[s - 1]) if b == 0 else dims[s]
            cout = dims[s]
            p = f'sparse_encoder.sp_cnn.conv_blocks_context.{s}.{b}.'
            out[p + 'conv1.weight'] = ((cout, cin, 3, 3, 3), 'conv_w')
            out[p + 'conv1.bias'] = ((cout,), 'bias')
            out[p + 'norm1.weight'] = ((cout,), 'gamma')
            out[p + 'norm1.bias'] = ((cout,), 'beta')
            out[p + 'conv2.weight'] = ((cout, cout, 3, 3, 3), 'conv_w')
            out[p + 'conv2.bias'] = ((cout,), 'bias')
            out[p + 'norm2.weight'] = ((cout,), 'gamma')
            out[p + 'norm2.bias'] = ((cout,), 'beta')
            if b == 0:
                out[p + 'conv3.weight'] = ((cout, cin, 1, 1, 1), 'conv_w')
                out[p + 'conv3.bias'] = ((cout,), 'bias')
    W = cfg.width
    chans = [W // 2 ** i for i in range(5)]
    for i in range(4):
        cin, cout = chans[i], chans[i + 1]
        p = f'dense_decoder.dec.{i}.'
        out[p + 'up_sample.weight'] = ((cin, cin, 4, 4, 4), 'convT_w')
        out[p + 'up_sample.bias'] = ((cin,), 'bias')
        out[p + 'conv.0.weight'] = ((cin, cin, 3, 3, 3), 'conv_w')
        for j, c in ((1, cin), (4, cout)):
            out[p + f'conv.{j}.weight'] = ((c,), 'gamma')
            out[p + f'conv.{j}.bias'] = ((c,), 'beta')
            out[p + f'conv.{j}.running_mean'] = ((c,), 'rmean')
            out[p + f'conv.{j}.running_var'] = ((c,), 'rvar')
            out[p + f'conv.{j}.num_batches_tracked'] = ((), 'nbt')
        out[p + 'conv.3.weight'] = ((cout, cin, 3, 3, 3), 'conv_w')
    out['dense_decoder.proj.weight'] = ((1, chans[4], 1, 1, 1), 'conv_w')
    out['dense_decoder.proj.bias'] = ((1,), 'bias')
    d_width = W
    for i in range(5):
        e = dims[4 - i]
        out[f'densify_norms.{i}.weight'] = ((e,), 'gamma')
        out[f'densify_norms.{i}.bias'] = ((e,), 'beta')
        if not (i == 0 and e == d_width):
            out[f'densify_projs.{i}.weight'] = ((d_width, e, 3, 3, 3), 'conv_w')
            out[f'densify_projs.{i}.bias'] = ((d_width,), 'bias')
        out[f'mask_tokens.{i}'] = ((1, e, 1, 1, 1), 'token')
        d_width //= 2
    return out


BUFFER_KINDS = ('rmean', 'rvar', 'nbt')


def make_state(cfg: Cfg, seed: int) -> Dict[str, Tensor]:
    """Deterministic non-degenerate weights (per-tensor seeded CPU generators; independent of the reference's RNG use).

    Scales chosen so that `rec` is far from zero (SURVEY.md §7 hard part 6) and every norm sees O(1) inputs.
    """
    st: Dict[str, Tensor] = {}
    for idx, (name, (shape, kind)) in enumerate(sorted(param_shapes(cfg).items())):
        g = torch.Generator().manual_seed(seed * 100003 + idx)
        if kind in ('conv_w', 'convT_w'):
            fan = shape[1] * shape[2] * shape[3] * shape[4]
            if kind == 'convT_w':
                fan = shape[0] * 8           # each output voxel of a k4/s2 transposed conv sees 8 taps per in-channel
            t = torch.randn(shape, generator=g) * (1.0 / math.sqrt(fan))
        elif kind == 'bias':
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == 'gamma':
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == 'beta':
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == 'token':
            t = 0.5 * torch.randn(shape, generator=g)
        elif kind == 'rmean':
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == 'rvar':
            t = 1.0 + 0.2 * torch.rand(shape, generator=g)
        elif kind == 'nbt':
            t = torch.zeros((), dtype=torch.long)
        else:
            raise KeyError(kind)
        st[name] = t
    return st


def make_input(cfg: Cfg, batch: int, seed: int) -> Tensor:
    g = torch.Generator().manual_seed(seed * 7919 + 17)
    return torch.randn(batch, cfg.in_ch, *cfg.input_size, generator=g)


def random_mask(cfg: Cfg, batch: int, generator: Optional[torch.Generator] = None) -> Tensor:
    """P/spark3D.py:92-96 — `len_keep` smallest of B×L uniform draws are active."""
    f = cfg.fmap
    idx = torch.rand(batch, cfg.L, generator=generator).argsort(dim=1)[:, :cfg.len_keep]
    m = torch.zeros(batch, cfg.L, dtype=torch.bool).scatter_(1, idx, True)
    return m.view(batch, 1, *f)


# ----------------------------------------------------------------------------------------------------------------
# layers
# ----------------------------------------------------------------------------------------------------------------
def upsample_mask(active: Tensor, size: Tuple[int, int, int]) -> Tensor:
    """P/encoder3D.py:7-10 — nearest up-sampling of the (B,1,f,f,f) bool mask to a feature-map size."""
    r = [size[i] // active.shape[2 + i] for i in range(3)]
    return active.repeat_interleave(r[0], 2).repeat_interleave(r[1], 3).repeat_interleave(r[2], 4)


def masked_conv(x: Tensor, w: Tensor, b: Optional[Tensor], stride: int, active: Tensor) -> Tensor:
    """P/encoder3D.py:12-15 — dense conv (+bias) then zero the inactive OUTPUT positions."""
    y = F.conv3d(x, w, b, stride=stride, padding=w.shape[-1] // 2)
    return y * upsample_mask(active, y.shape[2:])


def pooled_masked_norm(x: Tensor, gamma: Tensor, beta: Tensor, eps: float, active: Tensor) -> Tensor:
    """P/encoder3D.py:149-158 — statistics pooled over ALL active voxels of the local batch (biased variance);
    exact zeros at inactive voxels.  (Identical formula to sp_bn_forward in training mode, :17-25.)"""
    m = upsample_mask(active, x.shape[2:]).to(x.dtype)                        # (B,1,D,H,W)
    n = m.sum()
    mean = (x * m).sum(dim=(0, 2, 3, 4), keepdim=True) / n
    var = (((x - mean) ** 2) * m).sum(dim=(0, 2, 3, 4), keepdim=True) / n
    y = (x - mean) / torch.sqrt(var + eps) * gamma.view(1, -1, 1, 1, 1) + beta.view(1, -1, 1, 1, 1)
    return y * m


def res_block(P: Dict[str, Tensor], pre: str, x: Tensor, stride: int, has_1x1: bool, active: Tensor) -> Tensor:
    """P/STUNet_head.py:96-103 with the sparse layer swap of P/encoder3D.py:298-364 (encoder eps = 1e-5)."""
    y = masked_conv(x, P[pre + 'conv1.weight'], P[pre + 'conv1.bias'], stride, active)
    y = F.leaky_relu(pooled_masked_norm(y, P[pre + 'norm1.weight'], P[pre + 'norm1.bias'], 1e-5, active), 0.01)
    y = masked_conv(y, P[pre + 'conv2.weight'], P[pre + 'conv2.bias'], 1, active)
    y = pooled_masked_norm(y, P[pre + 'norm2.weight'], P[pre + 'norm2.bias'], 1e-5, active)
    if has_1x1:
        x = masked_conv(x, P[pre + 'conv3.weight'], P[pre + 'conv3.bias'], stride, active)
    return F.leaky_relu(y + x, 0.01)


def encoder(P: Dict[str, Tensor], cfg: Cfg, x: Tensor, active: Tensor) -> List[Tensor]:
    """P/STUNet_head.py:67-76 (hierarchical=True)."""
    feats = []
    for s in range(5):
        for b in range(cfg.depth):
            pre = f'sparse_encoder.sp_cnn.conv_blocks_context.{s}.{b}.'
            x = res_block(P, pre, x, stride=(2 if (s > 0 and b == 0) else 1), has_1x1=(b == 0), active=active)
        feats.append(x)
    return feats


def densify(P: Dict[str, Tensor], cfg: Cfg, feats: List[Tensor], active: Tensor) -> List[Tensor]:
    """P/spark3D.py:111-126 — coarse→fine: norm (eps 1e-6) → mask-token fill → 3³ proj (Identity at level 0)."""
    to_dec = []
    cur = active
    for i, fea in enumerate(reversed(feats)):
        y = pooled_masked_norm(fea, P[f'densify_norms.{i}.weight'], P[f'densify_norms.{i}.bias'], 1e-6, active)
        y = torch.where(cur.expand_as(y), y, P[f'mask_tokens.{i}'].expand_as(y))
        if f'densify_projs.{i}.weight' in P:
            y = F.conv3d(y, P[f'densify_projs.{i}.weight'], P[f'densify_projs.{i}.bias'], padding=1)
        to_dec.append(y)
        cur = cur.repeat_interleave(2, 2).repeat_interleave(2, 3).repeat_interleave(2, 4)
    return to_dec


def batch_norm(P: Dict[str, Tensor], pre: str, x: Tensor, training: bool,
               new_buffers: Optional[Dict[str, Tensor]]) -> Tensor:
    """nn.BatchNorm3d semantics (P/decoder3D.py:46): eps 1e-5, momentum 0.1, unbiased var into running_var."""
    g, b = P[pre + 'weight'].view(1, -1, 1, 1, 1), P[pre + 'bias'].view(1, -1, 1, 1, 1)
    if training:
        mean = x.mean(dim=(0, 2, 3, 4))
        var = x.var(dim=(0, 2, 3, 4), unbiased=False)
        if new_buffers is not None:
            n = x.numel() // x.shape[1]
            with torch.no_grad():
                new_buffers[pre + 'running_mean'] = 0.9 * P[pre + 'running_mean'] + 0.1 * mean.detach()
                new_buffers[pre + 'running_var'] = 0.9 * P[pre + 'running_var'] + 0.1 * var.detach() * n / (n - 1)
                new_buffers[pre + 'num_batches_tracked'] = P[pre + 'num_batches_tracked'] + 1
    else:
        mean, var = P[pre + 'running_mean'], P[pre + 'running_var']
    return (x - mean.view(1, -1, 1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1, 1) + 1e-5) * g + b


def decoder(P: Dict[str, Tensor], to_dec: List[Tensor], training: bool,
            new_buffers: Optional[Dict[str, Tensor]]) -> Tensor:
    """P/decoder3D.py:55-63 — only 4 blocks, so to_dec[4] (finest densify level) is never consumed."""
    x = 0
    for i in range(4):
        pre = f'dense_decoder.dec.{i}.'
        x = x + to_dec[i]
        x = F.conv_transpose3d(x, P[pre + 'up_sample.weight'], P[pre + 'up_sample.bias'], stride=2, padding=1)
        x = F.conv3d(x, P[pre + 'conv.0.weight'], None, padding=1)
        x = batch_norm(P, pre + 'conv.1.', x, training, new_buffers)
        x = F.relu6(x)
        x = F.conv3d(x, P[pre + 'conv.3.weight'], None, padding=1)
        x = batch_norm(P, pre + 'conv.4.', x, training, new_buffers)
    return F.conv3d(x, P['dense_decoder.proj.weight'], P['dense_decoder.proj.bias'])


def patchify(cfg: Cfg, x: Tensor) -> Tensor:
    """P/spark3D.py:148-155 — (B,C,D,H,W) → (B, L, p³·C); l=(h·f+w)·f+d, n=((p·16+q)·16+g)·C+c."""
    p = cfg.ratio
    h, w, d = cfg.fmap
    B, C = x.shape[:2]
    x = x.reshape(B, C, h, p, w, p, d, p).permute(0, 2, 4, 6, 3, 5, 7, 1)
    return x.reshape(B, h * w * d, p ** 3 * C)


def patch_loss(cfg: Cfg, inp: Tensor, rec: Tensor, active: Tensor) -> Tuple[Tensor, Tensor]:
    """P/spark3D.py:130-138 == P/AnatoMask.py:190-202.  Returns (scalar loss, per-patch masked loss (B,L))."""
    t, r = patchify(cfg, inp), patchify(cfg, rec)
    mean = t.mean(dim=-1, keepdim=True)
    var = t.var(dim=-1, keepdim=True)               # unbiased
    t = (t - mean) / (var + 1e-6) ** .5
    l2 = ((r - t) ** 2).mean(dim=2)
    non_active = active.logical_not().int().view(active.shape[0], -1)
    per_patch = l2 * non_active
    return per_patch.sum() / (non_active.sum() + 1e-8), per_patch


def teacher_patch_loss(cfg: Cfg, inp: Tensor, rec: Tensor, active: Tensor) -> Tensor:
    """P/pretrain_AntoMask.py:423-425 — RAW (un-normalised) per-patch MSE × non-active."""
    t, r = patchify(cfg, inp), patchify(cfg, rec)
    l2 = ((r - t) ** 2).mean(dim=2)
    return l2 * active.logical_not().int().view(active.shape[0], -1)


def forward(P: Dict[str, Tensor], cfg: Cfg, inp: Tensor, active: Tensor, training: bool = True,
            new_buffers: Optional[Dict[str, Tensor]] = None, keep: Optional[dict] = None) -> Tensor:
    """P/spark3D.py:98-127 up to the reconstruction `rec_bchwd`."""
    x = inp * upsample_mask(active, inp.shape[2:])
    feats = encoder(P, cfg, x, active)
    to_dec = densify(P, cfg, feats, active)
    rec = decoder(P, to_dec, training, new_buffers)
    if keep is not None:
        keep['feats'], keep['to_dec'] = feats, to_dec
    return rec


# ----------------------------------------------------------------------------------------------------------------
# AnatoMask hard-mask generation (P/AnatoMask.py:81-135), RNG-replaying parity mode
# ----------------------------------------------------------------------------------------------------------------
def hard_mask_lengths(cfg: Cfg, epoch: int, total_epoch: int, guide: bool = True) -> Tuple[int, int]:
    keep_ratio = float((epoch + 1) / total_epoch) * 0.5 if guide else 2 / 3
    nm = cfg.L - cfg.len_keep
    len_loss = int(nm * keep_ratio)
    return len_loss, nm - len_loss


def generate_mask(cfg: Cfg, loss_pred: Tensor, epoch: int, total_epoch: int, guide: bool = True,
                  np_rng=np.random) -> Tensor:
    """Bit-exact restatement of the hard-mask branch, consuming the numpy global RNG exactly like the reference
    (two shuffles per sample; the second one only advances RNG state).  Returns (B,1,f,f,f) bool, True = active."""
    B, L = loss_pred.shape
    len_loss, easy_len = hard_mask_lengths(cfg, epoch, total_epoch, guide)
    if len_loss <= 0:
        noise = torch.randn(B, L, device=loss_pred.device)
        keep = torch.argsort(noise, dim=1)[:, :cfg.len_keep]
        m = torch.zeros(B, L, dtype=torch.bool).scatter_(1, keep, True)
        return m.view(B, 1, *cfg.fmap)
    order = torch.argsort(loss_pred, dim=1)
    m = torch.zeros(B, L, dtype=torch.bool)
    for b in range(B):
        hard = order[b, L - len_loss:].cpu().numpy()
        rest = np.delete(np.arange(L), hard)
        np_rng.shuffle(rest)
        m[b, torch.from_numpy(rest[:cfg.len_keep])] = True
        # second shuffle (easy_mask bookkeeping of the reference) — result unused, RNG state consumed
        easy = order[b, L - len_loss - easy_len:L - len_loss].cpu().numpy()
        rest2 = np.delete(np.arange(L), easy)
        np_rng.shuffle(rest2)
    return m.view(B, 1, *cfg.fmap)


def hard_set(cfg: Cfg, loss_pred: Tensor, epoch: int, total_epoch: int) -> Tensor:
    """The RNG-free part of generate_mask: indices of the `len_loss` highest-loss patches per sample, sorted."""
    len_loss, _ = hard_mask_lengths(cfg, epoch, total_epoch)
    order = torch.argsort(loss_pred, dim=1)
    return order[:, cfg.L - len_loss:].sort(dim=1).values


# ----------------------------------------------------------------------------------------------------------------
# steps
# ----------------------------------------------------------------------------------------------------------------
def ema_update(ema: Dict[str, Tensor], model: Dict[str, Tensor], decay: float) -> None:
    """timm ModelEma.update: every state_dict entry, `ema.copy_(ema*d + (1-d)*model)` in the EMA dtype
    (so int64 num_batches_tracked is truncated)."""
    with torch.no_grad():
        for k, v in ema.items():
            v.copy_(v * decay + (1. - decay) * model[k].detach())


def ema_decay(epoch: int, epochs: int) -> float:
    """P/pretrain_AntoMask.py:383-386."""
    q = epochs // 4
    return 0.999 + epoch / q * (0.9999 - 0.999) if epoch < q else 0.9999


def split_state(st: Dict[str, Tensor], cfg: Cfg):
    kinds = param_shapes(cfg)
    params = {k: v for k, v in st.items() if kinds[k][1] not in BUFFER_KINDS}
    buffers = {k: v for k, v in st.items() if kinds[k][1] in BUFFER_KINDS}
    return params, buffers


def live_param_names(cfg: Cfg) -> List[str]:
    """Parameters that receive a gradient: everything except the dead finest densify level."""
    kinds = param_shapes(cfg)
    dead = ('densify_norms.4.', 'densify_projs.4.', 'mask_tokens.4')
    return [k for k, (_, kind) in kinds.items() if kind not in BUFFER_KINDS and not k.startswith(dead)]


def spark_loss_and_grads(st: Dict[str, Tensor], cfg: Cfg, inp: Tensor, active: Tensor):
    """One SparK forward/backward (P/pretrain.py:404-406).  Returns dict with rec, per_patch, loss, grads, new buffers."""
    P = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and k in live_param_names(cfg) else v)
         for k, v in st.items()}
    newbuf: Dict[str, Tensor] = {}
    keep: dict = {}
    rec = forward(P, cfg, inp, active, True, newbuf, keep)
    loss, per_patch = patch_loss(cfg, inp, rec, active)
    names = live_param_names(cfg)
    grads = torch.autograd.grad(loss, [P[k] for k in names])
    return dict(rec=rec.detach(), per_patch=per_patch.detach(), loss=loss.detach(),
                grads=dict(zip(names, grads)), new_buffers=newbuf,
                feats=[f.detach() for f in keep['feats']], to_dec=[t.detach() for t in keep['to_dec']])


class RefTrainer:
    """Step glue of P/pretrain.py:349-411 and P/pretrain_AntoMask.py:221-441 over the functional port."""

    def __init__(self, cfg: Cfg, state: Dict[str, Tensor], lr: float = 1e-4, weight_decay: float = 1e-5,
                 clip: float = 12., epochs: int = 1000, anatomask: bool = True):
        self.cfg, self.clip, self.epochs, self.anatomask = cfg, clip, epochs, anatomask
        self.state = {k: v.clone() for k, v in state.items()}
        self.names = live_param_names(cfg)
        for k in self.names:
            self.state[k].requires_grad_(True)
        # both groups get the same weight decay (SURVEY.md §9 "Optimiser")
        self.opt = torch.optim.AdamW([self.state[k] for k in self.names], lr=lr, betas=(0.9, 0.999),
                                     weight_decay=weight_decay)
        self.ema = {k: v.detach().clone() for k, v in self.state.items()} if anatomask else None

    def spark_step(self, inp: Tensor, active: Tensor) -> float:
        newbuf: Dict[str, Tensor] = {}
        rec = forward(self.state, self.cfg, inp, active, True, newbuf)
        loss, _ = patch_loss(self.cfg, inp, rec, active)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([self.state[k] for k in self.names], self.clip)
        self.opt.step()
        with torch.no_grad():
            for k, v in newbuf.items():
                self.state[k].copy_(v)
        return float(loss)

    def anatomask_step(self, inp: Tensor, mask1: Tensor, epoch: int, np_rng=np.random):
        cfg = self.cfg
        with torch.no_grad():                              # teacher: eval mode → BN running stats
            rec1 = forward(self.ema, cfg, inp, mask1, training=False)
            recon = teacher_patch_loss(cfg, inp, rec1, mask1)
        mask = generate_mask(cfg, recon, epoch, self.epochs - 1, True, np_rng)
        loss = self.spark_step(inp, mask)
        ema_update(self.ema, self.state, ema_decay(epoch, self.epochs))
        return loss, mask, recon
