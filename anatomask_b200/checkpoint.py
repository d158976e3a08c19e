"""Checkpoint compatibility with the reference scripts.

Writer: the `<model>_head_{latest,best}.pt` dictionary of P/pretrain.py:450-463 / P/pretrain_AntoMask.py:472-479 — keys
prefixed `module.` because the scripts save the LocalDDP/DDP wrapper's state_dict.  The EMA teacher and the arena
optimiser state are saved too (the reference drops the teacher; resuming needs it).
Reader-side contract: `load_stunet_ssl_weights` (nnunetv2/run/load_pretrained_weights.py:66-105) keeps entries whose key
contains 'encoder' and strips everything up to 'sp_cnn.'; `encoder_state_for_finetune` applies the same mapping so a
fine-tuning STUNet can `load_state_dict(..., strict=False)` the result."""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional

import torch


def head_checkpoint(model: torch.nn.Module, engine=None, train_loss=None, val_loss=None, epoch: int = 0) -> Dict:
    # .contiguous(): engine parameters may be strided views of taps-major arena storage; the file holds the reference's layout
    sd = OrderedDict(('module.' + k, v.detach().clone().contiguous()) for k, v in model.state_dict().items())
    ckpt = {'network_weights': sd, 'optimizer_state': None, 'grad_scaler_state': None, 'train_loss': train_loss,
            'val_loss': val_loss, 'current_epoch': epoch}
    if engine is not None:
        from . import AnatoMask as _am
        # Adam moments per parameter, in the parameter's own shape and element order (independent of the arena's storage order)
        moments = lambda flat: OrderedDict((n, v.detach().clone().contiguous()) for n, v in engine.arena.views(flat).items())
        ckpt['optimizer_state'] = {'exp_avg': moments(engine.m), 'exp_avg_sq': moments(engine.v), 'step': engine.t,
                                   'layout': dict(engine.arena.offsets), 'n_live': engine.arena.n_live,
                                   # device-RNG stream positions: without them a resumed run replays the masks of step 0
                                   'rng': {'step_counter': int(engine.step_counter.item()), 'rng_calls': engine._rng_calls,
                                           'mask_rng_offset': _am.SparK._rng_offset}}
        if engine.teacher is not None:
            ckpt['ema_weights'] = OrderedDict((k, v.detach().clone().contiguous()) for k, v in engine.teacher.state_dict().items())
    return ckpt


def save_head_checkpoint(path: str, model, engine=None, **kw) -> None:
    torch.save(head_checkpoint(model, engine, **kw), path)


def encoder_state_for_finetune(network_weights: Dict[str, torch.Tensor]) -> 'OrderedDict[str, torch.Tensor]':
    out = OrderedDict()
    for k, v in network_weights.items():
        if 'encoder' in k:
            out[k.split('sp_cnn.')[-1]] = v
    return out


def resume(engine, ckpt: Dict) -> None:
    """Restores student, teacher and optimiser state written by head_checkpoint (the reference has no resume path)."""
    sd = OrderedDict((k[len('module.'):] if k.startswith('module.') else k, v) for k, v in ckpt['network_weights'].items())
    engine.model.load_state_dict(sd)
    if engine.teacher is not None and 'ema_weights' in ckpt:
        engine.teacher.load_state_dict(ckpt['ema_weights'])
    opt: Optional[Dict] = ckpt.get('optimizer_state')
    if opt:
        if opt.get('layout') is not None and (dict(opt['layout']) != dict(engine.arena.offsets)
                                              or int(opt.get('n_live', -1)) != engine.arena.n_live):
            raise RuntimeError('resume: the checkpoint\'s parameter-arena layout differs from this engine\'s (different '
                               'model size or dead-parameter set) — the Adam moments would land on the wrong parameters')
        rng = opt.get('rng')
        if rng:
            from . import AnatoMask as _am
            engine.step_counter.fill_(int(rng['step_counter']))
            engine._rng_calls = int(rng['rng_calls'])
            _am.SparK._rng_offset = int(rng['mask_rng_offset'])
        for flat, saved in ((engine.m, opt['exp_avg']), (engine.v, opt['exp_avg_sq'])):
            if isinstance(saved, dict):
                views = engine.arena.views(flat)
                if set(views) != set(saved):
                    raise RuntimeError('resume: the checkpoint\'s Adam moments name other parameters than this engine\'s')
                for n, v in saved.items():
                    views[n].copy_(v)
            else:                                  # flat moments of an older checkpoint: only valid for the same storage order
                if getattr(engine.arena, 'taps_major', None):
                    raise RuntimeError('resume: flat Adam moments were saved in stock parameter order; this engine stores '
                                       'conv weights taps-major — rebuild it with PretrainEngine(..., taps_major=False)')
                flat.copy_(saved)
        engine.t = int(opt['step'])
