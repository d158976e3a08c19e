"""Oracle vs the LIVE reference modules (build container only: /root/reference is not on the GPU box, so these tests
skip there).  Complements the committed golden fixtures with randomised cases: the hard-mask schedule and shuffles, the
random mask, patchify and both per-patch losses are compared bit-for-bit / to fp32 round-off on fresh inputs every seed."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference/nnunetv2/training/nnUNetTrainer/variants/pretrain'
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present (GPU box)')

sys.path.insert(0, ROOT)
from oracle import reference_port as rp  # noqa: E402


@pytest.fixture(scope='module')
def ref_model():
    """The unmodified reference AnatoMask.SparK for the 'tiny' config (oracle/make_golden.py:build_reference)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('make_golden', os.path.join(ROOT, 'oracle', 'make_golden.py'))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    with contextlib.redirect_stdout(io.StringIO()):
        return mg.build_reference(rp.CONFIGS['tiny'], anatomask=True)


@pytest.mark.parametrize('seed,epoch,epochs', [(0, 0, 20), (1, 3, 20), (2, 9, 20), (3, 18, 20), (4, 500, 1000), (5, 998, 1000)])
def test_generate_mask_bit_exact_against_live_reference(ref_model, seed, epoch, epochs):
    cfg = rp.CONFIGS['tiny']
    B = 3
    g = torch.Generator().manual_seed(seed)
    mask1 = rp.random_mask(cfg, B, g)
    loss_pred = torch.rand(B, cfg.L, generator=g) * mask1.logical_not().view(B, -1)     # zeros on visible patches
    len_loss, _ = rp.hard_mask_lengths(cfg, epoch, epochs - 1)
    np.random.seed(seed)
    torch.manual_seed(seed)                       # the len_loss <= 0 branch draws torch.randn
    want, _ = ref_model.generate_mask(loss_pred.clone(), guide=True, epoch=epoch, total_epoch=epochs - 1)
    state_after_ref = np.random.get_state()[1].copy()
    np.random.seed(seed)
    torch.manual_seed(seed)
    got = rp.generate_mask(cfg, loss_pred.clone(), epoch, epochs - 1)
    assert torch.equal(got, want), (len_loss, int((got != want).sum()))
    assert int(got.sum()) == B * cfg.len_keep
    if len_loss > 0:                              # both consumed the numpy stream identically (two shuffles per sample)
        assert np.array_equal(np.random.get_state()[1], state_after_ref)


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_patchify_and_losses_against_live_reference(ref_model, seed):
    cfg = rp.CONFIGS['tiny']
    B = 2
    g = torch.Generator().manual_seed(100 + seed)
    inp = torch.randn(B, cfg.in_ch, *cfg.input_size, generator=g)
    rec = torch.randn(B, cfg.in_ch, *cfg.input_size, generator=g)
    active = rp.random_mask(cfg, B, g)
    assert torch.equal(rp.patchify(cfg, inp), ref_model.patchify(inp))
    # AnatoMask.forward_loss takes patchified tensors (P/AnatoMask.py:190-202)
    loss_r, pp_r = ref_model.forward_loss(ref_model.patchify(inp), ref_model.patchify(rec), active)
    loss_o, pp_o = rp.patch_loss(cfg, inp, rec, active)
    assert abs(float(loss_o) - float(loss_r)) <= 1e-6 * abs(float(loss_r))
    assert torch.allclose(pp_o, pp_r, rtol=1e-6, atol=1e-7)
    # teacher loss of the script (P/pretrain_AntoMask.py:423-425): raw patches
    inp1, rec1 = ref_model.patchify(inp), ref_model.patchify(rec)
    want = ((rec1 - inp1) ** 2).mean(dim=2) * active.logical_not().int().view(B, -1)
    assert torch.allclose(rp.teacher_patch_loss(cfg, inp, rec, active), want, rtol=1e-6, atol=1e-7)


def test_random_mask_matches_live_reference(ref_model):
    cfg = rp.CONFIGS['tiny']
    for seed in range(4):
        torch.manual_seed(seed)
        want = ref_model.mask(4, 'cpu')           # P/spark3D.py:92-96 draws from the default CPU generator
        torch.manual_seed(seed)
        got = rp.random_mask(cfg, 4, None)
        assert torch.equal(got, want)


@pytest.mark.parametrize('name,seed', [('tiny', 11), ('tiny', 12), ('S64', 13)])
def test_spark_step_against_live_reference_on_fresh_seeds(name, seed):
    """The golden comparison of test_oracle_golden.py repeated on seeds that have no committed fixture: full tensors (not
    digests) of rec, per-patch loss, every gradient and the BN buffers after the step."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('make_golden', os.path.join(ROOT, 'oracle', 'make_golden.py'))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    cfg = rp.CONFIGS[name]
    torch.set_num_threads(8)
    with contextlib.redirect_stdout(io.StringIO()):
        model = mg.build_reference(cfg, anatomask=False)
    state = rp.make_state(cfg, seed)
    model.load_state_dict(state)
    model.train()
    inp = rp.make_input(cfg, 2, seed)
    active = rp.random_mask(cfg, 2, torch.Generator().manual_seed(seed + 1))
    loss = model(inp, active_b1ff=active)
    loss.backward()
    out = rp.spark_loss_and_grads(state, cfg, inp, active)
    assert abs(float(out['loss']) - float(loss)) <= 2e-6 * abs(float(loss))
    ref_grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    assert sorted(ref_grads) == sorted(out['grads'])
    for k, gr in ref_grads.items():
        go = out['grads'][k]
        if float(gr.norm()) < 1e-6:                # analytically-zero bias gradients: noise on both sides
            assert float(go.norm()) < 1e-5, k
            continue
        rel = float((go - gr).norm() / gr.norm())
        assert rel < 5e-3, (k, rel)                # fp32 summation-order noise of 10^5-term reductions
    sd = model.state_dict()
    for k, v in out['new_buffers'].items():
        assert torch.allclose(v.to(sd[k].dtype), sd[k], rtol=1e-5, atol=1e-6), k


def test_public_signatures_match_live_reference():
    """Drop-in contract: the host mirror keeps the reference's class names, constructor and method signatures
    (parameter names, order and defaults)."""
    import importlib
    import inspect
    sys.path[:0] = [os.path.join(ROOT, 'oracle', 'timm_stub'), REF]
    pairs = [('encoder3D', ['SparseEncoder', 'SparseConv3d', 'SparseBatchNorm3d', 'SparseSyncBatchNorm3d',
                            'SparseInstanceNorm']),
             ('decoder3D', ['LightDecoder', 'UNetBlock']),
             ('STUNet_head', ['STUNet', 'BasicResBlock']),
             ('spark3D', ['SparK']), ('AnatoMask', ['SparK'])]
    methods = {'SparK': ['__init__', 'mask', 'forward', 'patchify', 'unpatchify', 'get_config', 'state_dict', 'load_state_dict',
                         'forward_loss', 'generate_mask'],
               'SparseEncoder': ['__init__', 'forward', 'dense_model_to_sparse'],
               'LightDecoder': ['__init__', 'forward'], 'STUNet': ['__init__', 'forward', 'get_downsample_ratio',
                                                                   'get_feature_map_channels'],
               'SparseInstanceNorm': ['__init__', 'forward'], 'BasicResBlock': ['__init__', 'forward'],
               'UNetBlock': ['__init__', 'forward']}

    def sig(fn):
        return [(p.name, p.kind, p.default if p.default is not inspect._empty else '<none>')
                for p in inspect.signature(fn).parameters.values()]

    for modname, classes in pairs:
        ref_mod = importlib.import_module(modname)
        our_mod = importlib.import_module('anatomask_b200.' + modname)
        for cname in classes:
            rc, oc = getattr(ref_mod, cname), getattr(our_mod, cname)
            assert [b.__name__ for b in oc.__mro__ if b.__module__.startswith('torch')][:1] == \
                   [b.__name__ for b in rc.__mro__ if b.__module__.startswith('torch')][:1], cname   # same torch base class
            for m in methods.get(cname, ['__init__']):
                if not hasattr(rc, m):
                    continue
                assert hasattr(oc, m), (modname, cname, m)
                assert sig(getattr(oc, m)) == sig(getattr(rc, m)), (modname, cname, m, sig(getattr(oc, m)), sig(getattr(rc, m)))


def test_mirror_host_side_methods_against_live_reference(ref_model):
    """The parts of the host mirror that are plain torch (no kernel): mask(), patchify / unpatchify, get_config, the
    len_loss <= 0 branch of generate_mask — same results as the reference module on CPU."""
    from anatomask_b200.trainer import build_model
    cfg = rp.CONFIGS['tiny']
    ours = build_model(base=cfg.base, depth=cfg.depth, input_size=cfg.input_size, device='cpu', anatomask=True)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 1, *cfg.input_size, generator=g)
    assert torch.equal(ours.patchify(x), ref_model.patchify(x))
    assert torch.equal(ours.unpatchify(ours.patchify(x)), x)
    assert torch.equal(ours.unpatchify(ours.patchify(x)), ref_model.unpatchify(ref_model.patchify(x)))
    for seed in range(3):
        torch.manual_seed(seed)
        want = ref_model.mask(3, 'cpu')
        torch.manual_seed(seed)
        assert torch.equal(ours.mask(3, 'cpu'), want)
    rc, oc = ref_model.get_config(), ours.get_config()
    assert set(rc) == set(oc) and all(rc[k] == oc[k] for k in rc), (rc, oc)
    # first epochs of a long schedule: len_loss = int(nm * keep_ratio) = 0 → torch.randn branch (P/AnatoMask.py:99-103)
    ours.mask_rng = 'numpy'
    loss_pred = torch.rand(2, cfg.L, generator=g)
    torch.manual_seed(3)
    want, _ = ref_model.generate_mask(loss_pred, guide=True, epoch=0, total_epoch=999)
    torch.manual_seed(3)
    got, _ = ours.generate_mask(loss_pred, guide=True, epoch=0, total_epoch=999)
    assert torch.equal(got, want)


def test_lr_schedule_matches_live_reference_scheduler():
    """trainer.lr_at_epoch (closed form) vs the reference's LinearWarmupCosineAnnealingLR stepped once per epoch as the
    scripts do (P/pretrain_AntoMask.py:359, N/training/lr_scheduler/LinearWarmupCosine.py:63-102), and the EMA decay
    schedule vs the script's inline formula (P/pretrain_AntoMask.py:383-386)."""
    import importlib.util
    import warnings
    from anatomask_b200.trainer import lr_at_epoch, ema_decay_at_epoch
    path = '/root/reference/nnunetv2/training/lr_scheduler/LinearWarmupCosine.py'
    spec = importlib.util.spec_from_file_location('ref_lwc', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for base_lr, warmup, epochs in [(1e-4, 20, 1000), (2e-4, 20, 300), (1e-3, 5, 40)]:
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.AdamW([p], lr=base_lr)
        sch = mod.LinearWarmupCosineAnnealingLR(opt, warmup, epochs, 1e-6)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            for ep in range(epochs):
                want = opt.param_groups[0]['lr']                 # the lr the script trains epoch `ep` with
                got = lr_at_epoch(ep, base_lr, warmup=warmup, max_epochs=epochs, warmup_start_lr=1e-6)
                assert got == pytest.approx(want, rel=1e-6, abs=1e-12), (base_lr, ep, got, want)
                opt.step()
                sch.step()
    src = open(os.path.join(REF, 'pretrain_AntoMask.py')).read()
    assert 'model_ema.decay = 0.999 + i / (epoch//4) * (0.9999 - 0.999)' in src      # the lines restated below
    for epoch in (1000, 300, 40):
        for i in range(epoch):
            want = 0.999 + i / (epoch // 4) * (0.9999 - 0.999) if i < epoch // 4 else 0.9999
            assert ema_decay_at_epoch(i, epoch) == pytest.approx(want, rel=0, abs=1e-15), (epoch, i)
            assert rp.ema_decay(i, epoch) == pytest.approx(want, rel=0, abs=1e-15)
