"""Per-kernel parity checks (CUDA path through the C ABI vs fp32 torch / the oracle on identical bf16-rounded inputs).

Used by tests/test_kernels_gpu.py (pytest -m gpu) and by tests/gpu_bringup.py, which runs every check in its own
process so that one trapped kernel cannot poison the rest.   CLI:  python -m tests.kernel_checks <check> [json-kwargs]
"""
from __future__ import annotations

import json
import sys

import numpy as np
import torch
import torch.nn.functional as F

bf16 = torch.bfloat16


def _dev():
    assert torch.cuda.is_available(), 'GPU checks need a B200'
    return torch.device('cuda:0')


def _rel(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _maxerr(a, b) -> float:
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _rand_mask(N, f, keep=0.4, seed=0):
    g = torch.Generator().manual_seed(seed)
    Lp = f ** 3
    k = max(1, round(Lp * keep))
    idx = torch.rand(N, Lp, generator=g).argsort(1)[:, :k]
    m = torch.zeros(N, Lp, dtype=torch.bool).scatter_(1, idx, True)
    return m.view(N, 1, f, f, f)


def _up(mask, size):
    r = size // mask.shape[-1]
    return mask.repeat_interleave(r, 2).repeat_interleave(r, 3).repeat_interleave(r, 4)


def check_conv(Cin=64, Cout=64, k=3, stride=1, S=16, N=2, impl=2, masked=False, f=2, bias=True, seed=0, tol=1.5e-2,
               keep=0.4):
    """conv fwd + dgrad + wgrad (+bias grad) vs fp32 torch on the same bf16-rounded operands."""
    from anatomask_b200 import ops
    dev = _dev()
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, Cin, S, S, S, generator=g).to(bf16)
    w = (torch.randn(Cout, Cin, k, k, k, generator=g) / (Cin * k ** 3) ** 0.5)
    b = torch.randn(Cout, generator=g) * 0.1 if bias else None
    mask = _rand_mask(N, f, keep=keep, seed=seed + 1) if masked else None
    if masked:
        x = x * _up(mask, S).to(bf16)
    wq = w.to(bf16).float()
    xr = x.float().to(dev).requires_grad_(True)
    wr = wq.to(dev).requires_grad_(True)
    br = b.to(dev).requires_grad_(True) if bias else None
    yr = F.conv3d(xr, wr, br, stride=stride, padding=k // 2)
    if masked:
        yr = yr * _up(mask, S // stride).to(dev)
    gy = torch.randn(yr.shape, generator=g).to(bf16)
    if masked:
        gy = gy * _up(mask, S // stride).to(bf16)
    yr.backward(gy.float().to(dev))

    xi = x.to(dev).permute(0, 2, 3, 4, 1).contiguous().requires_grad_(True)
    wp = wq.to(dev).requires_grad_(True)
    bp = b.to(dev).requires_grad_(True) if bias else None
    m = ops.MaskCtx(mask.to(dev)) if masked else None
    y = ops.conv3d(xi, wp, bp, k, stride, m, impl)
    y.backward(gy.to(dev).permute(0, 2, 3, 4, 1).contiguous())
    torch.cuda.synchronize()
    res = {'fwd': _rel(y.permute(0, 4, 1, 2, 3), yr.detach()), 'wgrad': _rel(wp.grad, wr.grad)}
    dx, dxr = xi.grad.permute(0, 4, 1, 2, 3).float(), xr.grad
    if masked:      # only the visible voxels of dx are consumed downstream
        mm = _up(mask, S).to(dev)
        dx, dxr = dx * mm, dxr * mm
    res['dgrad'] = _rel(dx, dxr)
    if bias:
        res['bgrad'] = _rel(bp.grad, br.grad)
    bad = {k_: v for k_, v in res.items() if not (v < tol)}
    assert not bad, f'conv Cin={Cin} Cout={Cout} k={k} s={stride} S={S} impl={impl} masked={masked}: {res}'
    return res


def check_convT(Cin=64, Cout=64, S=8, N=2, impl=2, seed=0, tol=1.5e-2, fwd_kernel=None):
    from anatomask_b200 import ops, _lib as L
    dev = _dev()
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, Cin, S, S, S, generator=g).to(bf16)
    w = (torch.randn(Cin, Cout, 4, 4, 4, generator=g) / (Cin * 8) ** 0.5).to(bf16).float()
    b = torch.randn(Cout, generator=g) * 0.1
    xr = x.float().to(dev).requires_grad_(True)
    wr = w.to(dev).requires_grad_(True)
    br = b.to(dev).requires_grad_(True)
    yr = F.conv_transpose3d(xr, wr, br, stride=2, padding=1)
    gy = torch.randn(yr.shape, generator=g).to(bf16)
    yr.backward(gy.float().to(dev))
    xi = x.to(dev).permute(0, 2, 3, 4, 1).contiguous().requires_grad_(True)
    wp = w.to(dev).requires_grad_(True)
    bp = b.to(dev).requires_grad_(True)
    y = ops.conv_transpose3d(xi, wp, bp, impl)
    if fwd_kernel is not None:          # the dispatcher must have picked this kernel for the forward pass
        got = (L.load().amb_last_conv_kernel() or b'').decode()
        assert got == fwd_kernel, f'convT Cin={Cin} Cout={Cout} S={S}: forward ran on {got}, expected {fwd_kernel}'
    y.backward(gy.to(dev).permute(0, 2, 3, 4, 1).contiguous())
    torch.cuda.synchronize()
    res = {'fwd': _rel(y.permute(0, 4, 1, 2, 3), yr.detach()), 'dgrad': _rel(xi.grad.permute(0, 4, 1, 2, 3), xr.grad),
           'wgrad': _rel(wp.grad, wr.grad), 'bgrad': _rel(bp.grad, br.grad)}
    bad = {k_: v for k_, v in res.items() if not (v < tol)}
    assert not bad, f'convT Cin={Cin} Cout={Cout} S={S} impl={impl}: {res}'
    return res


def check_conv_stats(Cin=32, Cout=64, S=16, N=2, seed=0):
    """fused Σy / Σy² epilogue of the tcgen05 kernel vs sums over its own bf16 output's fp32 source."""
    import ctypes as C
    from anatomask_b200 import ops, _lib as L
    dev = _dev()
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, S, S, S, Cin, generator=g).to(bf16).to(dev)
    w = (torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (Cin * 27) ** 0.5).to(dev)
    wp = ops._pack(w, 27, Cout, Cin, 1, Cin * 27, 27)
    y = torch.empty(N, S, S, S, Cout, dtype=bf16, device=dev)
    stats = torch.zeros(2 * Cout, dtype=torch.float64, device=dev)
    ops._conv_call(L.OP_CONV, L.IMPL_TCGEN05, (N, S, S, S), Cin, Cout, 3, 1, x, y, wp, stats=stats)
    yr = F.conv3d(x.permute(0, 4, 1, 2, 3).float(), w.to(bf16).float(), padding=1)
    s1, s2 = yr.sum((0, 2, 3, 4)).double(), (yr * yr).sum((0, 2, 3, 4)).double()
    torch.cuda.synchronize()
    res = {'sum': float((stats[:Cout] - s1).abs().max() / s1.abs().max()), 'sumsq': _rel(stats[Cout:], s2)}
    assert res['sum'] < 2e-3 and res['sumsq'] < 2e-3, res
    return res


def check_norm(Cc=64, S=16, N=2, f=2, mode='sparse', act=1, residual=True, seed=0, tol=1.2e-2):
    """pooled masked norm (+act, +residual) / dense BatchNorm (+running stats) / densify fill: fwd + bwd vs the oracle formula."""
    from anatomask_b200 import ops
    from oracle import reference_port as rp
    dev = _dev()
    g = torch.Generator().manual_seed(seed)
    mask = _rand_mask(N, f, seed=seed + 1)
    mu = _up(mask, S).float()
    x = (torch.randn(N, Cc, S, S, S, generator=g) * 1.5 + 0.3).to(bf16)
    r = torch.randn(N, Cc, S, S, S, generator=g).to(bf16) if residual else None
    if mode != 'dense':
        x = x * mu.to(bf16)
        if residual:
            r = r * mu.to(bf16)
    gamma = (1 + 0.2 * torch.randn(Cc, generator=g))
    beta = 0.2 * torch.randn(Cc, generator=g)
    token = 0.5 * torch.randn(1, Cc, 1, 1, 1, generator=g) if mode == 'fill' else None
    gy = torch.randn(N, Cc, S, S, S, generator=g).to(bf16)
    if mode == 'sparse':
        gy = gy * mu.to(bf16)
    eps = 1e-5
    # reference (fp32, CPU)
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rr = r.float().requires_grad_(True) if residual else None
    tr = token.clone().requires_grad_(True) if token is not None else None
    rm0, rv0 = torch.zeros(Cc), torch.ones(Cc)
    if mode == 'dense':
        P = {'w.weight': gr, 'w.bias': br, 'w.running_mean': rm0, 'w.running_var': rv0,
             'w.num_batches_tracked': torch.zeros((), dtype=torch.long)}
        nb = {}
        yr = rp.batch_norm(P, 'w.', xr, True, nb)
    else:
        yr = rp.pooled_masked_norm(xr, gr, br, eps, mask)
    if mode == 'fill':
        yr = torch.where(_up(mask, S).expand_as(yr), yr, tr.expand_as(yr))
    if residual:
        yr = yr + rr
    yr = F.leaky_relu(yr, 0.01) if act == 1 else (F.relu6(yr) if act == 2 else yr)
    yr.backward(gy.float())
    # CUDA
    m = ops.MaskCtx(mask.to(dev)) if mode != 'dense' else None
    xi = x.to(dev).permute(0, 2, 3, 4, 1).contiguous().requires_grad_(True)
    ri = r.to(dev).permute(0, 2, 3, 4, 1).contiguous().requires_grad_(True) if residual else None
    gp, bp = gamma.to(dev).requires_grad_(True), beta.to(dev).requires_grad_(True)
    tp = token.to(dev).requires_grad_(True) if token is not None else None
    running = None
    if mode == 'dense':
        running = (rm0.to(dev), rv0.to(dev), torch.zeros((), dtype=torch.long, device=dev))
    y = ops.NormFn.apply(xi, gp, bp, ri, tp, eps, act, m, running, 0.1)
    y.backward(gy.to(dev).permute(0, 2, 3, 4, 1).contiguous())
    torch.cuda.synchronize()
    res = {'fwd': _rel(y.permute(0, 4, 1, 2, 3).cpu(), yr.detach()),
           'dx': _rel(xi.grad.permute(0, 4, 1, 2, 3).cpu(), xr.grad),
           'dgamma': _rel(gp.grad.cpu(), gr.grad), 'dbeta': _rel(bp.grad.cpu(), br.grad)}
    if residual:
        res['dres'] = _rel(ri.grad.permute(0, 4, 1, 2, 3).cpu(), rr.grad)
    if token is not None:
        res['dtoken'] = _rel(tp.grad.cpu(), tr.grad)
    if mode == 'dense':
        res['rmean'] = _maxerr(running[0].cpu(), nb['w.running_mean'])
        res['rvar'] = _maxerr(running[1].cpu(), nb['w.running_var'])
        assert int(running[2]) == 1
    bad = {k_: v for k_, v in res.items() if not (v < tol)}
    assert not bad, f'norm C={Cc} S={S} mode={mode} act={act} residual={residual}: {res}'
    return res


def check_stem(Cc=32, S=32, N=2, f=2, seed=0, tol=1.2e-2):
    from anatomask_b200 import ops
    dev = _dev()
    g = torch.Generator().manual_seed(seed)
    mask = _rand_mask(N, f, seed=seed + 1)
    inp = torch.randn(N, 1, S, S, S, generator=g)
    w1 = torch.randn(Cc, 1, 3, 3, 3, generator=g) / 27 ** 0.5
    b1, w3, b3 = 0.1 * torch.randn(Cc, generator=g), torch.randn(Cc, 1, 1, 1, 1, generator=g), 0.1 * torch.randn(Cc, generator=g)
    mu = _up(mask, S).float()
    ps = [t.clone().requires_grad_(True) for t in (w1, b1, w3, b3)]
    xm = inp * mu
    y1 = F.conv3d(xm, ps[0], ps[1], padding=1) * mu
    y3 = F.conv3d(xm, ps[2], ps[3]) * mu
    g1 = (torch.randn(y1.shape, generator=g) * mu).to(bf16)
    g3 = (torch.randn(y3.shape, generator=g) * mu).to(bf16)
    (y1 * g1.float() + y3 * g3.float()).sum().backward()
    m = ops.MaskCtx(mask.to(dev))
    pc = [t.to(dev).requires_grad_(True) for t in (w1, b1, w3, b3)]
    o1, o3 = ops.StemFn.apply(inp.to(dev), *pc, m)
    torch.autograd.backward([o1, o3], [g1.to(dev).permute(0, 2, 3, 4, 1).contiguous(),
                                       g3.to(dev).permute(0, 2, 3, 4, 1).contiguous()])
    torch.cuda.synchronize()
    res = {'y1': _rel(o1.permute(0, 4, 1, 2, 3).cpu(), y1.detach()), 'y3': _rel(o3.permute(0, 4, 1, 2, 3).cpu(), y3.detach())}
    for nme, a, b in zip(('dw1', 'db1', 'dw3', 'db3'), pc, ps):
        res[nme] = _rel(a.grad.cpu(), b.grad)
    bad = {k_: v for k_, v in res.items() if not (v < tol)}
    assert not bad, f'stem: {res}'
    return res


def check_repack(seed=0):
    """amb_pack_weight / amb_unpack_wgrad in the four layouts ops.py uses, bit-exact against torch."""
    from anatomask_b200 import ops, _lib as L
    dev = _dev()
    g = torch.Generator().manual_seed(seed)
    out = {}
    for (co, ci, T) in [(64, 64, 27), (24, 16, 27), (136, 72, 27), (32, 64, 1), (64, 32, 64), (8, 8, 27), (512, 256, 27)]:
        w = torch.randn(co, ci, T, generator=g).to(dev)             # conv layout (Cout, Cin, taps)
        fwd = ops._pack(w, T, co, ci, 1, ci * T, T)                 # [t][co][ci]
        dgr = ops._pack(w, T, ci, co, 1, T, ci * T)                 # [t][ci][co]
        assert torch.equal(fwd, w.permute(2, 0, 1).to(bf16)), ('fwd', co, ci, T)
        assert torch.equal(dgr, w.permute(2, 1, 0).to(bf16)), ('dgrad', co, ci, T)
        wt = torch.randn(ci, co, T, generator=g).to(dev)            # ConvTranspose layout (Cin, Cout, taps)
        tf = ops._pack(wt, T, co, ci, 1, T, co * T)                 # [t][co][ci]
        assert torch.equal(tf, wt.permute(2, 1, 0).to(bf16)), ('convT', co, ci, T)
        dwp = torch.randn(T, co, ci, generator=g).to(dev)           # packed gradient → parameter layout
        dw = torch.full((co, ci, T), float('nan'), device=dev)
        L.call('amb_unpack_wgrad', ops._p(dwp), ops._p(dw), T, co, ci, 1, ci * T, T, ops._stream())
        assert torch.equal(dw, dwp.permute(1, 2, 0)), ('unpack', co, ci, T)
        dwt = torch.full((ci, co, T), float('nan'), device=dev)
        L.call('amb_unpack_wgrad', ops._p(dwp), ops._p(dwt), T, co, ci, 1, T, co * T, ops._stream())
        assert torch.equal(dwt, dwp.permute(2, 1, 0)), ('unpack convT', co, ci, T)
        out[f'{co}x{ci}x{T}'] = 'exact'
    # taps-major master weights (trainer.ParamArena): the module-shaped strided view packs to the same operands as the stock
    # tensor — forward form by plain conversion, input-gradient form by the per-tap transpose kernel; through the per-call
    # entry point and through the batched job table; a misaligned (odd-offset) source takes the scalar path
    from anatomask_b200.trainer import _taps_major_view
    for (co, ci, k, tr) in [(64, 64, 3, False), (24, 40, 3, False), (136, 72, 3, False), (32, 64, 1, False), (64, 32, 4, True),
                            (8, 8, 3, False), (512, 256, 3, False), (40, 24, 4, True)]:
        T = k ** 3
        shape = (ci, co, k, k, k) if tr else (co, ci, k, k, k)
        stock = torch.randn(*shape, generator=g).to(dev)
        for misalign in (0, 1):
            store = torch.zeros(stock.numel() + 4, device=dev)
            wv = _taps_major_view(store[misalign:misalign + stock.numel()], shape, tr)
            wv.copy_(stock)
            al = ops.packed_alias(wv, tr)
            assert al is not None and al.data_ptr() == store.data_ptr() + 4 * misalign
            ops.PACK_RECORD = rec = []
            try:
                fwd, dgr = ops._pack_conv(wv, tr, False), ops._pack_conv(wv, tr, True)
            finally:
                ops.PACK_RECORD = None
            want_f, want_d = ops._pack_conv(stock, tr, False), ops._pack_conv(stock, tr, True)
            assert torch.equal(fwd, want_f) and torch.equal(dgr, want_d), ('taps-major', co, ci, k, tr, misalign)
            assert torch.equal(fwd, al.to(bf16)) and torch.equal(dgr, al.transpose(1, 2).to(bf16))
            plan = ops.PackPlan(rec)
            assert plan.n_jobs == 2, (plan.n_jobs, co, ci, k)
            plan.run()
            got = list(plan.cache.values())
            assert torch.equal(got[0], want_f) and torch.equal(got[1], want_d), ('taps-major batched', co, ci, k, tr, misalign)
            # gradient: a [tap][Cout][Cin] buffer viewed in the parameter's shape without a copy
            dwp = torch.randn(T, co, ci, generator=g).to(dev)
            back = ops._from_packed(dwp, wv, tr)
            assert back.data_ptr() == dwp.data_ptr() and tuple(back.shape) == shape
            assert torch.equal(back.contiguous(), ops._from_packed(dwp, stock, tr))
        out[f'taps-major {co}x{ci}x{k}{"T" if tr else ""}'] = 'exact'
    torch.cuda.synchronize()
    return out


def check_proj(Cc=32, S=16, N=2, seed=0, tol=1e-2):
    from anatomask_b200 import ops
    dev = _dev()
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, Cc, S, S, S, generator=g).to(bf16)
    w, b = torch.randn(1, Cc, 1, 1, 1, generator=g) / Cc ** 0.5, torch.randn(1, generator=g)
    xr, wr, br = x.float().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.conv3d(xr, wr, br)
    gy = torch.randn(yr.shape, generator=g)
    yr.backward(gy)
    xi = x.to(dev).permute(0, 2, 3, 4, 1).contiguous().requires_grad_(True)
    wp, bp = w.to(dev).requires_grad_(True), b.to(dev).requires_grad_(True)
    y = ops.ProjFn.apply(xi, wp, bp)
    y.backward(gy.to(dev))
    torch.cuda.synchronize()
    res = {'fwd': _rel(y.cpu(), yr.detach()), 'dx': _rel(xi.grad.permute(0, 4, 1, 2, 3).cpu(), xr.grad),
           'dw': _rel(wp.grad.cpu(), wr.grad), 'db': _rel(bp.grad.cpu(), br.grad)}
    bad = {k_: v for k_, v in res.items() if not (v < tol)}
    assert not bad, f'proj: {res}'
    return res


def check_loss(S=32, N=2, seed=0):
    from anatomask_b200 import ops
    from oracle import reference_port as rp
    dev = _dev()
    cfg = rp.Cfg(base=8, input_size=(S, S, S))
    g = torch.Generator().manual_seed(seed)
    inp = torch.randn(N, 1, S, S, S, generator=g) * 2 + 0.5
    rec = torch.randn(N, 1, S, S, S, generator=g)
    mask = rp.random_mask(cfg, N, g)
    rr = rec.clone().requires_grad_(True)
    loss_r, pp_r = rp.patch_loss(cfg, inp, rr, mask)
    (loss_r * 1.7).backward()
    raw_r = rp.teacher_patch_loss(cfg, inp, rec, mask)
    rc = rec.to(dev).requires_grad_(True)
    au8 = mask[:, 0].to(torch.uint8).contiguous().to(dev)
    loss, pp = ops.PatchLossFn.apply(inp.to(dev), rc, au8, True)
    (loss * 1.7).backward()
    _, raw = ops.PatchLossFn.apply(inp.to(dev), rec.to(dev), au8, False)
    torch.cuda.synchronize()
    res = {'loss': abs(float(loss) - float(loss_r)) / abs(float(loss_r)), 'per_patch': _rel(pp.cpu(), pp_r.detach()),
           'drec': _rel(rc.grad.cpu(), rr.grad), 'raw': _rel(raw.cpu(), raw_r)}
    assert res['loss'] < 1e-5 and res['per_patch'] < 1e-5 and res['drec'] < 1e-5 and res['raw'] < 1e-5, res
    return res


def check_hard_mask(B=4, Lp=512, len_keep=205, seed=0):
    """hard set bit-exact vs argsort; random fill = exactly len_keep visible, none of them hard, varies with offset."""
    from anatomask_b200 import ops
    from oracle import reference_port as rp
    dev = _dev()
    g = torch.Generator().manual_seed(seed)
    cfg = rp.CONFIGS['B128'] if Lp == 512 else rp.Cfg(base=8, input_size=(32, 32, 32))
    for epoch, total in ((0, 999), (100, 999), (500, 999), (998, 999)):
        loss = torch.rand(B, Lp, generator=g) * 3
        m1 = torch.rand(B, Lp, generator=g) < 0.4
        loss = loss * (~m1)                       # visible patches carry exactly-zero teacher loss
        len_loss, _ = rp.hard_mask_lengths(cfg, epoch, total)
        hard, mask = ops.hard_mask(loss.to(dev), len_loss, len_keep, seed=7, offset=epoch)
        hard2, mask2 = ops.hard_mask(loss.to(dev), len_loss, len_keep, seed=7, offset=epoch + 1)
        torch.cuda.synchronize()
        order = torch.argsort(loss, dim=1)
        assert torch.equal(hard.cpu().long(), order[:, Lp - len_loss:]), f'hard set differs at epoch {epoch}'
        mk = mask.cpu().bool()
        assert (mk.sum(1) == len_keep).all()
        if len_loss:
            assert not mk.gather(1, hard.cpu().long()).any(), 'a hard patch was left visible'
        assert not torch.equal(mask.cpu(), mask2.cpu()), 'RNG offset ignored'
    return {'ok': 1}


def check_ema_adamw(n=1000003, seed=0):
    from anatomask_b200 import ops
    dev = _dev()
    g = torch.Generator().manual_seed(seed)
    n_al = (n + 3) // 4 * 4
    e, mdl = torch.randn(n_al, generator=g), torch.randn(n_al, generator=g)
    d = 0.999 + 3 / 250 * (0.9999 - 0.999)
    ref = e * d + (1. - d) * mdl
    ec = e.to(dev)
    ops.ema_update_(ec, mdl.to(dev), d)
    torch.cuda.synchronize()
    assert torch.equal(ec.cpu(), ref), f'EMA not bit-exact: {float((ec.cpu() - ref).abs().max())}'
    # AdamW with global-norm clipping, 3 steps vs torch.optim.AdamW + clip_grad_norm_
    p0 = torch.randn(n_al, generator=g)
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pr], lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-2)
    pc = p0.to(dev)
    m, v = torch.zeros_like(pc), torch.zeros_like(pc)
    for step in range(1, 4):
        gr = torch.randn(n_al, generator=g) * (5.0 if step == 2 else 0.001)
        pr.grad = gr.clone()
        torch.nn.utils.clip_grad_norm_([pr], 12.0)
        opt.step()
        ops.adamw_step_(pc, gr.to(dev), m, v, 1e-3, (0.9, 0.999), 1e-8, 1e-2, step, 12.0)
    torch.cuda.synchronize()
    err = float((pc.cpu() - pr.detach()).abs().max())
    assert err < 2e-6, f'AdamW differs: {err}'
    return {'ema': 0.0, 'adamw': err}


def check_layout(Cc=24, S=12, N=2):
    from anatomask_b200 import ops
    dev = _dev()
    x = torch.randn(N, Cc, S, S + 2, S + 4)
    xi = ops.to_internal(x.to(dev))
    back = ops.to_ncdhw_f32(xi)
    torch.cuda.synchronize()
    assert torch.equal(xi.permute(0, 4, 1, 2, 3).float().cpu(), x.to(bf16).float())
    assert torch.equal(back.cpu(), x.to(bf16).float())
    return {'ok': 1}


def check_sparse_bn(Cc=32, S=16, N=2, f=2, seed=0, tol=1.2e-2):
    """SparseBatchNorm3d as a MODULE (P/encoder3D.py:17-25,39-41): training mode normalises with the statistics of the
    visible voxels only and updates running_mean / running_var (unbiased) / num_batches_tracked from them; eval mode uses
    the running statistics on visible voxels and leaves masked voxels exactly zero.  Oracle: the reference's own recipe —
    gather the visible voxels, run torch's BatchNorm1d over them, scatter back into zeros."""
    from anatomask_b200 import encoder3D
    dev = _dev()
    g = torch.Generator().manual_seed(seed)
    mask = _rand_mask(N, f, seed=seed + 1)
    up = _up(mask, S)
    res = {}
    bn_ref = torch.nn.BatchNorm1d(Cc)
    bn = encoder3D.SparseBatchNorm3d(Cc).to(dev)
    with torch.no_grad():
        bn_ref.weight.copy_(1 + 0.2 * torch.randn(Cc, generator=g)); bn_ref.bias.copy_(0.2 * torch.randn(Cc, generator=g))
        bn_ref.running_mean.copy_(0.1 * torch.randn(Cc, generator=g)); bn_ref.running_var.copy_(1 + 0.3 * torch.rand(Cc, generator=g))
    bn.load_state_dict({k: v.to(dev) for k, v in bn_ref.state_dict().items()})

    def reference(x, train):
        bn_ref.train(train)
        ii = up[:, 0].nonzero(as_tuple=True)                                    # (b, d, h, w) of the visible voxels
        bhwc = x.permute(0, 2, 3, 4, 1)
        nc = bn_ref(bhwc[ii])
        out = torch.zeros_like(bhwc)
        out[ii] = nc
        return out.permute(0, 4, 1, 2, 3)

    for step, train in enumerate((True, True, False)):
        x = ((torch.randn(N, Cc, S, S, S, generator=g) * 1.5 + 0.3) * up).to(bf16)
        gy = (torch.randn(N, Cc, S, S, S, generator=g) * up).to(bf16)
        xr = x.float().requires_grad_(True)
        yr = reference(xr, train)
        bn.train(train)
        encoder3D._cur_active = mask.to(dev)
        xi = x.to(dev).requires_grad_(True)
        y = bn(xi)
        if train:
            yr.backward(gy.float())
            y.backward(gy.to(dev))
            res[f'dx_{step}'] = _rel(xi.grad.float().cpu() * up, xr.grad * up)
            res[f'dgamma_{step}'] = _rel(bn.weight.grad.cpu(), bn_ref.weight.grad)
            res[f'dbeta_{step}'] = _rel(bn.bias.grad.cpu(), bn_ref.bias.grad)
            bn.zero_grad(); bn_ref.zero_grad()
        torch.cuda.synchronize()
        res[f'fwd_{step}_{"train" if train else "eval"}'] = _rel(y.float().cpu(), yr.detach())
        assert float(y.float().cpu().abs().mul((~up).float()).max()) == 0.0, 'masked voxels must stay exactly zero'
        res[f'rmean_{step}'] = _maxerr(bn.running_mean.cpu(), bn_ref.running_mean)
        res[f'rvar_{step}'] = _maxerr(bn.running_var.cpu(), bn_ref.running_var)
        assert int(bn.num_batches_tracked) == int(bn_ref.num_batches_tracked)
    bad = {k_: v for k_, v in res.items() if not (v < tol)}
    assert not bad, f'SparseBatchNorm3d C={Cc} S={S}: {res}'
    return res


def check_conv_bn_eval(Cin=64, Cout=64, S=16, N=2, act=2, impl=0, seed=0, tol=1.2e-2):
    """Conv3d (no bias) + inference-mode BatchNorm (+ReLU6) fused in the conv epilogue (the EMA teacher's decoder) against
    fp32 torch conv3d → batch_norm(eval) → relu6 on the same bf16-rounded operands."""
    from anatomask_b200 import ops
    dev = _dev()
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, Cin, S, S, S, generator=g).to(bf16)
    w = (torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (Cin * 27) ** 0.5).to(bf16).float()
    gamma, beta = 1 + 0.3 * torch.randn(Cout, generator=g), 0.3 * torch.randn(Cout, generator=g)
    rm, rv = 0.2 * torch.randn(Cout, generator=g), 0.5 + torch.rand(Cout, generator=g)
    yr = F.batch_norm(F.conv3d(x.float(), w, None, 1, 1), rm, rv, gamma, beta, False, 0.1, 1e-5)
    yr = F.relu6(yr) if act == 2 else yr
    xi = x.to(dev).permute(0, 2, 3, 4, 1).contiguous()
    with torch.no_grad():
        y = ops.conv3d_bn_eval(xi, w.to(dev), gamma.to(dev), beta.to(dev), rm.to(dev), rv.to(dev), 1e-5, act, impl=impl)
    torch.cuda.synchronize()
    res = {'fwd': _rel(y.permute(0, 4, 1, 2, 3).cpu(), yr)}
    assert res['fwd'] < tol, f'conv+BN(eval) Cin={Cin} Cout={Cout} S={S} act={act} impl={impl}: {res}'
    return res


CHECKS = {n[6:]: f for n, f in list(globals().items()) if n.startswith('check_')}

if __name__ == '__main__':
    name = sys.argv[1]
    kwargs = json.loads(sys.argv[2]) if len(sys.argv) > 2 else {}
    out = CHECKS[name](**kwargs)
    print('RESULT', name, json.dumps(kwargs), json.dumps(out))


# ----------------------------------------------------------------------------------------------------------------
# remaining sparse-layer API: the CUDA modules against the oracle port on the seeded cases of
# oracle/sparse_layers_port.py (whose reference results are committed in tests/golden/sparse_layers.pt)
# ----------------------------------------------------------------------------------------------------------------
def build_sparse_layer_module(name):
    """Our module for a case, built the way a user of the reference would build it (incl. dense_model_to_sparse)."""
    import torch.nn as nn
    from anatomask_b200 import encoder3D as enc, MedNeXt_head as mh
    from oracle import sparse_layers_port as sl
    kind, opt = sl.CASES[name]
    Cc = sl.CASE_C
    if kind == 'group_norm':
        return enc.SparseGroupNorm(opt['groups'], Cc, eps=1e-5)
    if kind == 'layer_norm':
        return enc.SparseConvNeXtLayerNorm(Cc, eps=1e-6, data_format=opt['fmt'], sparse=opt.get('sparse', True))
    if kind == 'pool':
        if opt['mode'] == 'max':
            return enc.SparseMaxPooling(opt['k'], opt['s'], opt['p'])
        return enc.SparseAvgPooling(opt['k'], opt['s'], opt['p'], count_include_pad=opt.get('include_pad', True))
    if kind == 'adaptive_avg':
        return enc.SparseEncoder.dense_model_to_sparse(nn.AdaptiveAvgPool3d(1))
    if kind == 'dwconv':
        return enc.SparseEncoder.dense_model_to_sparse(nn.Conv3d(Cc, Cc, opt['k'], opt['s'], opt['k'] // 2, groups=Cc))
    if kind == 'convnext':
        m = enc.SparseConvNeXtBlock(Cc, drop_path=0., layer_scale_init_value=0.5, sparse=True, ks=7)
        return enc.SparseEncoder.dense_model_to_sparse(m) if opt['converted'] else m
    if kind == 'mednext':
        blk = mh.MedNeXtDownBlock(Cc, 2 * Cc, exp_r=2, kernel_size=3, do_res=True, norm_type='group') if opt['down'] else \
            mh.MedNeXtBlock(Cc, Cc, exp_r=2, kernel_size=3, do_res=True, norm_type='group')
        return enc.SparseEncoder.dense_model_to_sparse(blk)
    raise KeyError(kind)


def check_sparse_layer_case(name, tol=1.5e-2):
    from anatomask_b200 import encoder3D as enc
    from oracle import sparse_layers_port as sl
    dev = _dev()
    kind, opt = sl.CASES[name]
    m = build_sparse_layer_module(name)
    x, active, g = sl.case_inputs(name)
    params = sl.case_params({k: tuple(v.shape) for k, v in m.named_parameters()}, g)
    # oracle (CPU fp32) on the same bf16-exact inputs and parameters
    xr = x.clone().requires_grad_(True)
    pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    yr = sl.run_case(name, pr, xr, active)
    dy = sl.case_dy(yr.shape, name)
    yr.backward(dy)
    for k, p in m.named_parameters():
        p.data.copy_(params[k])
    m = m.to(dev).train()
    enc._cur_active = active.to(dev)
    xc = x.to(dev).requires_grad_(True)
    y = m(xc)
    y.backward(dy.to(dev).to(y.dtype))
    torch.cuda.synchronize()
    # GroupNorm with one channel per group (MedNeXt's nn.GroupNorm(C, C), and the gn_gC case) normalises every element by
    # itself: x̂ ≡ 0, the output is the bias, and every gradient upstream of x̂ is EXACTLY zero.  torch's group-norm backward
    # returns rstd³-amplified rounding noise there (|x·Σdy − Σdy·x|·rstd³ with rstd = eps^-½ ≈ 316), so those tensors are
    # checked against the true value: zero (ours: exactly; the oracle's noise: small against its own dy scale).
    vanishing = {'gn_gC': ('x', 'weight'), 'mednext': ('conv1.weight', 'conv1.bias', 'norm.weight'),
                 'mednext_down': ('conv1.weight', 'conv1.bias', 'norm.weight')}.get(name, ())
    res = {'y': _rel(y.float().cpu(), yr.detach())}
    got_grads = {'x': xc.grad, **{k: p.grad for k, p in m.named_parameters()}}
    ref_grads = {'x': xr.grad, **{k: v.grad for k, v in pr.items()}}
    for k, gr in ref_grads.items():
        if gr is None:
            continue
        assert got_grads[k] is not None, f'{name}: no gradient for {k}'
        if k in vanishing:
            assert float(got_grads[k].float().abs().max()) <= 1e-6 * float(dy.abs().max()), f'{name}: d{k} must vanish'
            continue
        res['d' + k] = _rel(got_grads[k].float().cpu(), gr)
    if kind in ('group_norm', 'layer_norm', 'dwconv', 'pool') and opt.get('sparse', True):
        fmt_last = kind == 'layer_norm' and opt['fmt'] == 'channels_last'
        yy = y.permute(0, 4, 1, 2, 3) if fmt_last else y
        up = sl.expand_mask(active, yy.shape[2:]).to(dev)
        assert float((yy.float().abs() * (~up)).max()) == 0.0, f'{name}: masked voxels must stay exactly zero'
    bad = {k: v for k, v in res.items() if not (v < tol)}
    assert not bad, f'sparse layer case {name}: {res}'
    return res


# ----------------------------------------------------------------------------------------------------------------
# device-side input pipeline against the numpy / scipy restatement of the reference's CPU pipeline
# ----------------------------------------------------------------------------------------------------------------
def _synthetic_case(shape, seed):
    """CT-like test volume: smooth structures + noise, fp32."""
    g = np.random.RandomState(seed)
    z, y, x = np.meshgrid(*[np.linspace(-1, 1, s) for s in shape], indexing='ij')
    vol = np.sin(3.1 * x + 1.7 * y) * np.cos(2.3 * z - 0.4 * x) + 0.5 * np.exp(-4 * (x ** 2 + y ** 2 + z ** 2))
    return (vol + 0.3 * g.standard_normal(shape)).astype(np.float32)


def check_augment(patch=48, seed=0, tol=3e-4):
    """Every branch of the pipeline on forced parameters (no draw, rotation, zoom in / out, both, all mirror flips, bounding
    boxes reaching outside the case on both sides), then seeded random batches drawn in the reference's order."""
    from anatomask_b200 import augment as A
    from oracle import augment_port as O
    dev = _dev()
    aug = A.DeviceAugmenter(patch_size=(patch,) * 3, p_rot=0.5, p_scale=0.5)
    P = aug.initial_patch_size
    cases_np = [_synthetic_case((70, 96, 88), seed), _synthetic_case((P[0] - 9, 64, P[2] + 30), seed + 1)]
    cases = [torch.from_numpy(c).to(dev) for c in cases_np]
    R = A.rotation_matrix
    forced = [
        (0, (5, 7, 3), None, None, (False, False, False)),
        (0, (-11, 20, -6), None, None, (True, False, True)),                       # bbox reaches outside the case: zero padding
        (0, (2, 9, 4), (0.3, -0.2, 0.45), None, (False, True, False)),
        (0, (-20, -15, 30), (0.5, 0.5, -0.5), None, (True, True, True)),
        (1, (0, -12, 10), None, 0.75, (False, False, True)),                       # zoom in (scale < 1)
        (1, (-4, 3, 40), None, 1.35, (False, False, False)),                       # zoom out: samples beyond the initial patch → 0
        (1, (3, -8, 25), (-0.52, 0.1, 0.33), 1.2, (True, False, False)),
    ]
    worst = 0.0
    for ci, lb, angles, sc, flips in forced:
        mat = None
        if angles is not None or sc is not None:
            mat = (R(*angles) if angles is not None else np.eye(3)) * (sc if sc is not None else 1.0)
        got = aug([cases[ci]], [A.SampleParams(lb, mat, flips)])[0, 0].cpu().numpy()
        want = O.pipeline_sample(cases_np[ci], lb, P, aug.patch_size, angles, sc, flips)
        err = float(np.abs(got - want).max()) / float(np.abs(want).max())
        worst = max(worst, err)
        assert err < tol, f'augment lb={lb} angles={angles} scale={sc} flips={flips}: rel max err {err:.3e}'
        if mat is None:
            assert np.array_equal(got, want), 'crop + mirror must be bit-exact'
    for s in range(4):                                                             # random batches, reference draw order
        r1, r2 = np.random.RandomState(100 + s), np.random.RandomState(100 + s)
        inp, params = aug.sample(cases, r1)
        ref = O.draw_batch(r2, [c.shape for c in cases_np], P, aug.patch_size, aug.angle, aug.scale, aug.p_rot, aug.p_scale)
        for j, (lb, angles, sc, flips) in enumerate(ref):
            want = O.pipeline_sample(cases_np[j], lb, P, aug.patch_size, angles, sc, flips)
            err = float(np.abs(inp[j, 0].cpu().numpy() - want).max()) / max(float(np.abs(want).max()), 1e-6)
            worst = max(worst, err)
            assert err < tol, f'augment batch {s} sample {j}: {err:.3e}'
    return {'worst_rel_max_err': worst}


def check_augment_full_size(reps=3, tol=3e-4):
    """BASELINE size (128³ out of a 205³ initial patch): one rotated + scaled + mirrored sample against the scipy oracle
    (a few seconds of CPU), the mirrored crop flipped back equals the crop bit-for-bit, and throughput of both branches."""
    import time
    from anatomask_b200 import augment as A
    from oracle import augment_port as O
    dev = _dev()
    aug = A.DeviceAugmenter()
    P, Osz = aug.initial_patch_size, aug.patch_size
    case_np = _synthetic_case((230, 260, 240), 3)
    case = torch.from_numpy(case_np).to(dev)
    lb = (10, -25, 57)
    angles, sc, flips = (0.41, -0.3, 0.22), 1.15, (False, True, True)
    got = aug([case], [A.SampleParams(lb, A.rotation_matrix(*angles) * sc, flips)])[0, 0].cpu().numpy()
    want = O.pipeline_sample(case_np, lb, P, Osz, angles, sc, flips)
    err_a = float(np.abs(got - want).max()) / float(np.abs(want).max())
    assert err_a < tol, err_a
    crop = aug([case], [A.SampleParams(lb, None, (False, False, False))])
    assert np.array_equal(crop[0, 0].cpu().numpy(), O.pipeline_sample(case_np, lb, P, Osz, None, None, (False, False, False)))
    res = {'spline_vs_scipy_rel_max_err': err_a}
    for name, mat in (('crop', None), ('spline', A.rotation_matrix(0.3, -0.2, 0.1) * 1.1)):
        ps = [A.SampleParams(lb, mat, (True, False, False))] * 2
        aug([case, case], ps)
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(reps):
            aug([case, case], ps)
        torch.cuda.synchronize()
        res[f'ms_per_volume_{name}'] = (time.time() - t0) / reps / 2 * 1e3
    print('RESULT augment_full_size', res)
    return res


def check_zero_shell(Cc=32, S=32, N=2, f=2, seed=0):
    """amb_zero_shell: exactly the masked voxels within one voxel (3x3x3 neighbourhood) of a visible patch of the same sample
    become zero; every other voxel keeps its bits."""
    from anatomask_b200 import ops
    dev = _dev()
    g = torch.Generator().manual_seed(seed)
    mask = _rand_mask(N, f, seed=seed + 1)
    x = (torch.randn(N, S, S, S, Cc, generator=g) + 3.0).to(bf16).to(dev)          # no zeros in the input
    vis = _up(mask, S).float().to(dev)                                             # (N,1,S,S,S)
    near = F.max_pool3d(vis, 3, 1, 1) > 0
    shell = (near & (vis == 0))[:, 0].unsqueeze(-1)
    want = torch.where(shell, torch.zeros_like(x), x)
    got = x.clone()
    ops.zero_shell(got, ops.MaskCtx(mask.to(dev)))
    torch.cuda.synchronize()
    assert torch.equal(got, want), f'zero_shell C={Cc} S={S} f={f}: {int((got != want).sum())} elements differ'
    return {'shell_voxels': int(shell.sum()), 'of': int(vis.numel())}


def check_add_parity0(Cc=32, S=32, N=2, f=2, masked=True, seed=0):
    """amb_add_parity0: fine[:, 2z, 2y, 2x] += coarse[:, z, y, x] (visible patches only when masked), bf16 round-to-nearest."""
    from anatomask_b200 import ops, _lib as L
    import ctypes as C
    dev = _dev()
    g = torch.Generator().manual_seed(seed)
    fine = torch.randn(N, S, S, S, Cc, generator=g).to(bf16).to(dev)
    coarse = torch.randn(N, S // 2, S // 2, S // 2, Cc, generator=g).to(bf16).to(dev)
    want = fine.clone()
    add = coarse.float()
    m = None
    if masked:
        mask = _rand_mask(N, f, seed=seed + 1)
        m = ops.MaskCtx(mask.to(dev))
        add = add * _up(mask, S // 2)[:, 0].unsqueeze(-1).float().to(dev)
    want[:, ::2, ::2, ::2] = (want[:, ::2, ::2, ::2].float() + add).to(bf16)
    geo = m.geo(coarse, True) if m is not None else ops.dense_geo(coarse)
    L.call('amb_add_parity0', C.byref(geo), ops._p(coarse), ops._p(fine), ops._stream())
    torch.cuda.synchronize()
    assert torch.equal(fine, want), f'add_parity0: {int((fine != want).sum())} elements differ'
    return {}


def check_conv_pair(Cin=32, Cout=64, S=32, N=2, f=2, stride=2, seed=0, tol=1.5e-2):
    """ConvPairFn (conv1 k3 + 1x1 shortcut on the same input, input gradient formed in place) against the two separate ConvFn
    nodes + autograd's add, on visible voxels (the only ones the engine reads)."""
    from anatomask_b200 import ops
    dev = _dev()
    g = torch.Generator().manual_seed(seed)
    mask = _rand_mask(N, f, seed=seed + 1)
    m = ops.MaskCtx(mask.to(dev))
    vis = _up(mask, S)[:, 0].unsqueeze(-1).to(dev)
    x = (torch.randn(N, S, S, S, Cin, generator=g).to(dev) * vis).to(bf16)
    w1 = (torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (27 * Cin) ** 0.5).to(dev)
    w3 = (torch.randn(Cout, Cin, 1, 1, 1, generator=g) / Cin ** 0.5).to(dev)
    b1, b3 = (0.1 * torch.randn(Cout, generator=g)).to(dev), (0.1 * torch.randn(Cout, generator=g)).to(dev)
    So = S // stride
    viso = _up(mask, So)[:, 0].unsqueeze(-1).to(dev)
    dy = (torch.randn(N, So, So, So, Cout, generator=g).to(dev) * viso).to(bf16)
    dsc = (torch.randn(N, So, So, So, Cout, generator=g).to(dev) * viso).to(bf16)
    outs = []
    for pair in (False, True):
        xs = x.clone().requires_grad_(True)
        ps = [t.clone().requires_grad_(True) for t in (w1, b1, w3, b3)]
        if pair:
            y, sc = ops.conv3d_pair(xs, ps[0], ps[1], ps[2], ps[3], 3, stride, m)
        else:
            y = ops.conv3d(xs, ps[0], ps[1], 3, stride, m)
            sc = ops.conv3d(xs, ps[2], ps[3], 1, stride, m)
        torch.autograd.backward([y, sc], [dy, dsc])
        ops.join_side_stream(dev)
        torch.cuda.synchronize()
        keep = lambda t, v: torch.where(v.bool(), t.float(), torch.zeros((), device=dev))      # masked voxels may be unwritten
        outs.append((keep(y, viso), keep(sc, viso), keep(xs.grad, vis), ps[0].grad, ps[2].grad, ps[3].grad))
    names = ('y', 'sc', 'dx', 'dw1', 'dw3', 'db3')
    res = {n: _rel(a, b) for n, a, b in zip(names, outs[1], outs[0])}
    assert res['y'] < 1e-3 and res['sc'] < 1e-3, res           # same kernels; split-K layers commit fp32 atomics in any order
    bad = {k: v for k, v in res.items() if not v < tol}
    assert not bad, f'conv pair Cin={Cin} Cout={Cout} S={S} stride={stride}: {res}'
    return res
