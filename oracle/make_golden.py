"""ORACLE — TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.pt by executing the UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python oracle/make_golden.py
The reference modules are imported from where they lie (P = /root/reference/nnunetv2/training/nnUNetTrainer/
variants/pretrain) with `oracle/timm_stub` standing in for the absent `timm`.  The four training *scripts* cannot
be imported (module-level dataset paths / `cuda:4`), so their step bodies are restated here verbatim in meaning:
P/pretrain.py:404-409 (SparK) and P/pretrain_AntoMask.py:419-440 (AnatoMask).

Fixtures are compact: weights/inputs are regenerated from seeds by oracle.reference_port.make_state/make_input
(torch CPU generators), so only results are stored: rec, per-patch loss, loss, per-tensor gradient digests
(norm, sum, first 64 values), BN buffers after the step, hard masks, post-step state digests.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/nnunetv2/training/nnUNetTrainer/variants/pretrain'
sys.path[:0] = [os.path.join(HERE, 'timm_stub'), REF, os.path.dirname(HERE)]

from oracle import reference_port as rp  # noqa: E402


def build_reference(cfg: rp.Cfg, anatomask: bool):
    import encoder3D  # noqa: F401  (reference)
    from encoder3D import SparseEncoder
    from decoder3D import LightDecoder
    from STUNet_head import STUNet
    import spark3D
    import AnatoMask
    head = STUNet(cfg.in_ch, 1, depth=[cfg.depth] * 6, dims=[cfg.base * x for x in (1, 2, 4, 8, 16, 16)],
                  pool_op_kernel_sizes=[[2, 2, 2]] * 4 + [[1, 1, 1]], conv_kernel_sizes=[[3, 3, 3]] * 6)
    with contextlib.redirect_stdout(io.StringIO()):
        enc = SparseEncoder(head, input_size=cfg.input_size, sbn=False)
        dec = LightDecoder(enc.downsample_ratio, sbn=False, width=cfg.width, out_channel=1)
        mod = (AnatoMask if anatomask else spark3D).SparK(enc, dec, mask_ratio=cfg.mask_ratio, densify_norm='in')
    return mod


def digest(t: torch.Tensor):
    t = t.detach().double().flatten()
    return dict(norm=float(t.norm()), sum=float(t.sum()), head=t[:64].float().clone(), numel=t.numel())


def golden_spark(name: str, cfg: rp.Cfg, batch: int, seed: int):
    torch.manual_seed(seed)
    model = build_reference(cfg, anatomask=False)
    shapes = rp.param_shapes(cfg)
    sd = model.state_dict()
    assert set(sd.keys()) == set(shapes.keys()), (set(sd) ^ set(shapes))
    for k, v in sd.items():
        assert tuple(v.shape) == shapes[k][0], (k, v.shape, shapes[k][0])
    state = rp.make_state(cfg, seed)
    model.load_state_dict(state)
    model.train()
    inp = rp.make_input(cfg, batch, seed)
    active = rp.random_mask(cfg, batch, torch.Generator().manual_seed(seed + 1))
    loss = model(inp, active_b1ff=active)
    loss.backward()
    grads = {k: digest(p.grad) for k, p in model.named_parameters() if p.grad is not None}
    dead = sorted(k for k, p in model.named_parameters() if p.grad is None)
    # the same module in AnatoMask form gives (inp,rec) patches and the per-patch loss
    am = build_reference(cfg, anatomask=True)
    am.load_state_dict(state)
    am.train()
    with torch.no_grad():
        inp_p, rec_p = am(inp, active_b1ff=active)
        loss2, per_patch = am.forward_loss(inp_p, rec_p, active)
        rec = am.unpatchify(rec_p)
    new_sd = model.state_dict()
    buffers = {k: new_sd[k].clone() for k, (_, kind) in shapes.items() if kind in rp.BUFFER_KINDS}
    out = dict(cfg=cfg.__dict__, batch=batch, seed=seed, active=active, loss=float(loss), loss_anatomask=float(loss2),
               per_patch=per_patch.clone(), rec=rec.clone() if rec.numel() <= (1 << 17) else None,
               rec_digest=digest(rec), grads=grads, dead=dead, buffers=buffers,
               param_names=sorted(shapes.keys()))
    torch.save(out, os.path.join(HERE, '..', 'tests', 'golden', f'spark_{name}.pt'))
    print(f'[golden] spark_{name}: loss={float(loss):.6f} (anatomask form {float(loss2):.6f}), '
          f'{len(grads)} grads, dead={dead}')


def golden_anatomask(name: str, cfg: rp.Cfg, batch: int, seed: int, epochs: int, epoch_list, lr=1e-3):
    """A few AnatoMask training steps with the reference modules + AdamW + ModelEma (P/pretrain_AntoMask.py:419-440)."""
    from timm.utils import ModelEma
    from utils.lr_control import get_param_groups
    torch.manual_seed(seed)
    np.random.seed(seed)
    model = build_reference(cfg, anatomask=True)
    state = rp.make_state(cfg, seed)
    model.load_state_dict(state)
    model_ema = ModelEma(model, decay=0.999, device='', resume='')
    with contextlib.redirect_stdout(io.StringIO()):
        groups = get_param_groups(model, nowd_keys={'cls_token', 'pos_embed', 'mask_token', 'gamma'})
    opt = torch.optim.AdamW(params=groups, lr=lr, betas=(0.9, 0.999), weight_decay=1e-5)
    steps = []
    for it, ep in enumerate(epoch_list):
        model.train()
        model_ema.decay = rp.ema_decay(ep, epochs)
        inp = rp.make_input(cfg, batch, seed + 10 + it)
        mask1 = rp.random_mask(cfg, batch, torch.Generator().manual_seed(seed + 100 + it))
        with torch.no_grad():
            inp1, rec1 = model_ema.ema(inp, active_b1ff=mask1)
            l2 = ((rec1 - inp1) ** 2).mean(dim=2, keepdim=False)
            non_active = mask1.logical_not().int().view(mask1.shape[0], -1)
            recon_loss = l2 * non_active
        torch.manual_seed(seed + 1000 + it)      # pins the len_loss<=0 (torch.randn) branch
        mask, _easy = model_ema.ema.generate_mask(recon_loss, guide=True, epoch=ep, total_epoch=epochs - 1)
        inpp, recc = model(inp, active_b1ff=mask, vis=False)
        loss, _ = model.forward_loss(inpp, recc, mask)
        opt.zero_grad()
        loss.backward()
        gn = torch.nn.utils.clip_grad_norm_(model.parameters(), 12).item()
        opt.step()
        model_ema.update(model)
        steps.append(dict(epoch=ep, mask1=mask1, teacher_loss=recon_loss.clone(), mask=mask.clone(),
                          loss=float(loss), grad_norm=gn))
        print(f'[golden] anatomask_{name} step {it} epoch {ep}: loss={float(loss):.6f} gn={gn:.4f} '
              f'active={int(mask.sum())}')
    out = dict(cfg=cfg.__dict__, batch=batch, seed=seed, epochs=epochs, lr=lr, steps=steps,
               student={k: digest(v) for k, v in model.state_dict().items()},
               teacher={k: digest(v) for k, v in model_ema.ema.state_dict().items()})
    torch.save(out, os.path.join(HERE, '..', 'tests', 'golden', f'anatomask_{name}.pt'))


if __name__ == '__main__':
    torch.set_num_threads(8)
    golden_spark('tiny', rp.CONFIGS['tiny'], batch=2, seed=3)
    golden_spark('S64', rp.CONFIGS['S64'], batch=2, seed=5)      # BASELINE config 1
    golden_spark('S_aniso', rp.CONFIGS['S_aniso'], batch=2, seed=4)   # non-cubic, odd mask grid (3,5,2)
    golden_spark('L32', rp.CONFIGS['L32'], batch=2, seed=2)      # STUNet-L: depth 2 (identity-shortcut blocks), width 1024
    golden_spark('B64', rp.CONFIGS['B64'], batch=2, seed=5)      # STUNet-B (the bench model) at 64^3, decoder width 512
    golden_anatomask('tiny', rp.CONFIGS['tiny'], batch=2, seed=7, epochs=20, epoch_list=[0, 9, 18])
    golden_anatomask('S64', rp.CONFIGS['S64'], batch=2, seed=9, epochs=1000, epoch_list=[0, 500, 998])
