"""GPU parity tests (pytest -m gpu): every kernel through the C ABI against fp32 torch / the oracle on identical inputs,
then the whole SparK / AnatoMask path through the reference-named modules against the oracle port."""
import pytest
import torch

from tests import kernel_checks as kc
from tests import model_checks as mc

pytestmark = pytest.mark.gpu
D, T, V1 = 1, 2, 3      # AMB_IMPL_DIRECT, AMB_IMPL_TCGEN05, AMB_IMPL_TCGEN05_V1


def test_library_loaded_and_no_cpu_path():
    from anatomask_b200 import _lib
    assert _lib.load().amb_sm_arch() == 100
    assert torch.cuda.is_available()


@pytest.mark.parametrize('name', ['layout', 'loss', 'hard_mask', 'ema_adamw', 'proj', 'stem'])
def test_hbm_bound_kernels(name):
    kc.CHECKS[name]()


@pytest.mark.parametrize('kw', [dict(Cc=64, S=32), dict(Cc=16, S=64, f=4), dict(Cc=8, S=32, N=1), dict(Cc=24, S=32)])
def test_stem_channel_widths(kw):
    """tiled stem wgrad: 8 / 4 / 2 / 1 channel groups per warp row; C = 24 takes the per-run kernel"""
    kc.check_stem(**kw)


def test_weight_repack_layouts():
    """tiled pack / unpack (taps innermost) against torch permutes, incl. ragged tiles, k = 1 and the 64-tap ConvT"""
    kc.check_repack()


@pytest.mark.parametrize('kw', [
    dict(mode='sparse', act=1, residual=True), dict(mode='sparse', act=0, residual=False, Cc=32, S=32),
    dict(mode='dense', act=2, residual=False), dict(mode='dense', act=0, residual=False, Cc=512, S=8),
    dict(mode='fill', act=0, residual=False),
    # row-group walk of the active-patch list: rows of 64 / 32 / 4 chunks, 1 / 2 / 4 rows per group, > 32 channel groups;
    # 96 channels (12 channel groups, not a power of two) stays on the per-chunk walk
    dict(mode='sparse', act=1, residual=True, Cc=512, S=8, f=8), dict(mode='sparse', act=1, residual=True, Cc=16, S=32, f=2),
    dict(mode='sparse', act=0, residual=False, Cc=8, S=16, f=4), dict(mode='sparse', act=1, residual=True, Cc=128, S=16, f=4),
    dict(mode='sparse', act=2, residual=False, Cc=256, S=8, f=4), dict(mode='sparse', act=1, residual=True, Cc=96, S=16, f=2)])
def test_norm_family(kw):
    kc.check_norm(**kw)


@pytest.mark.parametrize('kw', [
    dict(Cin=16, Cout=16, S=8, impl=D), dict(Cin=16, Cout=24, S=8, stride=2, impl=D),
    dict(Cin=8, Cout=16, S=8, k=1, stride=2, impl=D), dict(Cin=16, Cout=16, S=16, impl=D, masked=True),
    # tcgen05 implicit GEMM: 128B / 64B / 32B swizzle, N tiles, box clipping, stride 2, 1x1
    dict(Cin=64, Cout=64, S=16, impl=T), dict(Cin=32, Cout=32, S=16, impl=T), dict(Cin=16, Cout=16, S=16, impl=T),
    dict(Cin=128, Cout=256, S=8, impl=T), dict(Cin=512, Cout=512, S=8, impl=T),
    dict(Cin=64, Cout=32, S=16, impl=T, bias=False), dict(Cin=64, Cout=64, S=12, impl=T),
    dict(Cin=256, Cout=128, S=4, impl=T), dict(Cin=64, Cout=128, S=16, stride=2, impl=T),
    dict(Cin=32, Cout=64, S=16, k=1, stride=2, impl=T), dict(Cin=64, Cout=64, S=16, k=1, impl=T),
    dict(Cin=64, Cout=64, S=32, impl=T), dict(Cin=64, Cout=32, S=24, impl=T),
    dict(Cin=64, Cout=64, S=18, impl=T), dict(Cin=32, Cout=16, S=20, impl=T), dict(Cin=128, Cout=64, S=32, N=1, impl=T),
    # N-stacked halo kernel with wide outputs: Cout = 128 (2 tiles per unit, N = 128 / 256) and 256 (1 tile, N = 256)
    dict(Cin=128, Cout=128, S=16, impl=T), dict(Cin=64, Cout=256, S=16, N=1, impl=T), dict(Cin=256, Cout=128, S=24, N=1, impl=T),
    # masked: active-tile work-list (patch edge >= 8) and dense tiles + epilogue mask (patch edge 4)
    dict(Cin=32, Cout=32, S=32, impl=T, masked=True), dict(Cin=64, Cout=64, S=16, impl=T, masked=True),
    dict(Cin=128, Cout=128, S=8, impl=T, masked=True), dict(Cin=32, Cout=64, S=32, stride=2, impl=T, masked=True),
    dict(Cin=32, Cout=64, S=32, k=1, stride=2, impl=T, masked=True),
    dict(Cin=64, Cout=64, S=32, impl=T, masked=True, f=8),
    # halo-plane kernel walking the active-patch list (patch edge 16 and 32 voxels)
    dict(Cin=64, Cout=64, S=32, impl=T, masked=True, f=2), dict(Cin=64, Cout=32, S=64, N=1, impl=T, masked=True, f=4),
    dict(Cin=32, Cout=32, S=64, N=1, impl=T, masked=True, f=2)])
def test_conv_fwd_dgrad_wgrad(kw):
    kc.check_conv(**kw)


@pytest.mark.parametrize('kw', [dict(Cin=16, Cout=8, S=4, impl=D), dict(Cin=64, Cout=64, S=8, impl=T),
                                dict(Cin=512, Cout=512, S=4, impl=T), dict(Cin=32, Cout=32, S=8, impl=T),
                                dict(Cin=64, Cout=64, S=16, impl=T), dict(Cin=64, Cout=32, S=16, N=1, impl=T),
                                dict(Cin=128, Cout=64, S=16, N=1, impl=T), dict(Cin=32, Cout=16, S=20, N=1, impl=T),
                                # kz-stacked halo-plane forward (conv_igemm4t.cu: Cout <= 64, at least one unit per SM),
                                # incl. ragged extents (partial y / x tiles, odd depth)
                                dict(Cin=64, Cout=64, S=32, impl=T, fwd_kernel='igemm4t_kernel'),
                                dict(Cin=32, Cout=16, S=36, N=1, impl=T, fwd_kernel='igemm4t_kernel'),
                                dict(Cin=64, Cout=32, S=27, N=2, impl=T, fwd_kernel='igemm4t_kernel'),
                                dict(Cin=128, Cout=32, S=26, N=2, impl=T, fwd_kernel='igemm4t_kernel')])
def test_conv_transpose(kw):
    kc.check_convT(**kw)


@pytest.mark.parametrize('keep', [1.0, 0.7, 0.1])
@pytest.mark.parametrize('kw', [dict(Cin=32, Cout=32, S=32, f=2), dict(Cin=64, Cout=64, S=32, f=4),
                                dict(Cin=32, Cout=64, S=32, f=2, stride=2)])
def test_masked_conv_over_mask_ratios(kw, keep):
    """single masked layers from fully visible to 10 % visible (work-list kernels: igemm3 at patch edge 16,
    per-tap kernel at edge 8, stride 2)"""
    kc.check_conv(impl=T, masked=True, keep=keep, **kw)


@pytest.mark.parametrize('kw', [dict(), dict(Cin=64, Cout=32, S=16, act=0), dict(Cin=128, Cout=128, S=8), dict(Cin=512, Cout=256, S=8, act=0),
                                dict(Cin=16, Cout=24, S=8, impl=D), dict(Cin=32, Cout=32, S=20, N=1)])
def test_conv_with_fused_inference_batchnorm(kw):
    """teacher decoder: BN(eval)+ReLU6 folded into the conv epilogue — N-stacked halo kernel, per-tap kernel, CUDA-core kernel"""
    kc.check_conv_bn_eval(**kw)


@pytest.mark.parametrize('kw', [dict(), dict(Cin=64, Cout=64, S=32)])
def test_conv_epilogue_statistics(kw):
    kc.check_conv_stats(**kw)


def test_direct_and_tcgen05_agree_exactly_on_integers():
    """Differential check with small-integer data (every product and partial sum is exact in bf16/fp32): the CUDA-core
    gather kernel and the tcgen05 kernel must agree bit-for-bit, independently of summation order."""
    from anatomask_b200 import ops, _lib as L
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(0)
    x = torch.randint(-3, 4, (2, 16, 16, 16, 64), generator=g).to(torch.bfloat16).to(dev)
    w = torch.randint(-2, 3, (64, 64, 3, 3, 3), generator=g).float().to(dev)
    wp = ops._pack(w, 27, 64, 64, 1, 64 * 27, 27)
    ys = []
    for impl in (L.IMPL_DIRECT, L.IMPL_TCGEN05_V1, L.IMPL_TCGEN05):      # CUDA cores, per-tap tcgen05, N-stacked halo planes
        y = torch.empty(2, 16, 16, 16, 64, dtype=torch.bfloat16, device=dev)
        ops._conv_call(L.OP_CONV, impl, (2, 16, 16, 16), 64, 64, 3, 1, x, y, wp)
        ys.append(y)
    torch.cuda.synchronize()
    assert torch.equal(ys[0], ys[1]) and torch.equal(ys[0], ys[2])


@pytest.mark.parametrize('kw', [dict(), dict(Cc=64, S=32, f=4, seed=3)])
def test_sparse_batchnorm3d_module_train_and_eval(kw):
    """SparseBatchNorm3d under a mask — batch statistics + running-stat update in train, running stats in eval"""
    kc.check_sparse_bn(**kw)


# (config, seed): the triples of oracle/make_yardstick.py.  B128 batch 2 is the BASELINE headline configuration
# (CPU oracle ~40 s on 8 threads), L64 the STUNet-L encoder/decoder widths of BASELINE config 4.
@pytest.mark.parametrize('name,seed', [('tiny', 3), ('S64', 5), ('B64', 5), ('S_aniso', 4), ('L32', 2), ('L64', 2), ('B128', 5)])
def test_spark_step_matches_oracle(name, seed):
    """loss <= max(1e-3, torch-bf16's own error); every gradient tensor <= max(1e-2, 1.25 x torch-bf16-autocast's error on that
    tensor) — tests/golden/autocast_yardstick.json, see tests/model_checks.py"""
    mc.check_spark(name=name, seed=seed, verbose=False)


def test_script_step_bodies_under_the_module_wrapper():
    """The literal step bodies of P/pretrain.py:404-409 and P/pretrain_AntoMask.py:419-441 with the mirror under LocalDDP"""
    mc.check_script_spark()
    mc.check_script_anatomask()


def test_two_gpu_ddp_and_syncbn_checks(tmp_path):
    """tests/dist_checks.py under torchrun on 2 GPUs (NCCL): SyncBN forward/backward all-reduces, bucketed gradient exchange
    captured in the step graph, the literal script step under torch DDP.  Skipped on a 1-GPU box; its retained log from a
    2-GPU gpurun call is profiles/r2_dist_checks_2gpu.log."""
    import os, subprocess, sys
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (run with gpurun --gpus 2)')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29533', os.path.join(root, 'tests', 'dist_checks.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root)
    log = out.stdout + out.stderr
    os.makedirs(os.path.join(root, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(root, 'gpurun_out', 'dist_checks_2gpu.log'), 'w') as f:
        f.write(log)
    assert out.returncode == 0, log[-3000:]


def test_anatomask_steps_match_oracle():
    mc.check_anatomask_steps()


def test_golden_fixture_from_the_unmodified_reference(golden_dir):
    """The CUDA path against the committed fixture produced by the reference itself (oracle/make_golden.py)."""
    import os
    from oracle import reference_port as rp
    g = torch.load(os.path.join(golden_dir, 'spark_S64.pt'), weights_only=False)
    cfg = rp.Cfg(**g['cfg'])
    model = mc.build(cfg, g['seed'], anatomask=True)
    model.train()
    inp = rp.make_input(cfg, g['batch'], g['seed']).cuda()
    rec = model.reconstruct(inp, g['active'].cuda())
    loss, pp = model.forward_loss(inp, rec, g['active'].cuda())
    loss.backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - g['loss']) <= 1e-3 * abs(g['loss'])          # north-star tolerance: loss 1e-3 relative
    assert float((pp.cpu() - g['per_patch']).norm() / g['per_patch'].norm()) < 5e-3
    for k in ('dense_decoder.proj.weight', 'dense_decoder.dec.3.conv.3.weight', 'dense_decoder.dec.3.conv.4.weight'):
        d = g['grads'][k]
        got = dict(model.named_parameters())[k].grad.flatten()[:64].cpu().double()
        err = float((got - d['head'].double()).norm() / d['head'].double().norm())
        assert err < 2e-2, (k, err)                                       # bf16: 1e-2-class on well-conditioned tensors


def test_device_rng_step_runs_without_host_sync():
    mc.check_device_step()


def test_staged_input_copy_overlaps_the_step_and_feeds_the_right_batch():
    mc.check_staged_input()


def test_graph_replay_matches_eager_launches():
    mc.check_graph_matches_eager()


def test_full_size_properties_B128():
    """BASELINE size (STUNet-B, 128^3, batch 2): size-independent properties of one AnatoMask step."""
    from oracle import reference_port as rp
    from anatomask_b200.trainer import PretrainEngine, build_model
    cfg = rp.CONFIGS['B128']
    model = build_model('B', cfg.input_size, anatomask=True)
    eng = PretrainEngine(model, epochs=1000, anatomask=True, mask_rng='device')
    inp = torch.randn(2, 1, 128, 128, 128, device='cuda')
    loss, mask, recon = eng.step(inp, epoch=500)
    torch.cuda.synchronize()
    assert torch.isfinite(loss)
    assert mask.shape == (2, 1, 8, 8, 8) and int(mask.sum()) == 2 * cfg.len_keep          # exactly len_keep visible
    ll, _ = rp.hard_mask_lengths(cfg, 500, 999)
    order = torch.argsort(recon, dim=1)
    hard = order[:, cfg.L - ll:]
    assert not mask.view(2, -1).gather(1, hard).any()                                       # hard patches are masked
    assert (recon.view(2, -1) >= 0).all()
    # dead densify level keeps grad None; every live parameter received a finite gradient
    for n, p in model.named_parameters():
        if n.startswith(('densify_norms.4', 'densify_projs.4', 'mask_tokens.4')):
            assert p.grad is None
        else:
            assert p.grad is not None and torch.isfinite(p.grad).all(), n
    # encoder features are exactly zero outside visible patches (the zero-invariant of SURVEY.md §7.2)
    from anatomask_b200 import encoder3D
    feats = model.sparse_encoder(inp)
    m1 = encoder3D._cur_active
    up = m1.repeat_interleave(16, 2).repeat_interleave(16, 3).repeat_interleave(16, 4)
    assert float(feats[0].float().abs().mul((~up).float()).max()) == 0.0


@pytest.mark.parametrize('name', sorted(__import__('oracle.sparse_layers_port', fromlist=['CASES']).CASES))
def test_remaining_sparse_layers(name):
    """SparseGroupNorm, SparseConvNeXtLayerNorm, SparseMax/AvgPooling, SparseAdaptiveAvgPooling, depthwise
    SparseConv3d (k 3/5/7, stride 1/2), SparseConvNeXtBlock and converted MedNeXt blocks — the CUDA modules against the oracle
    port (itself pinned to the unmodified reference classes by tests/golden/sparse_layers.pt)."""
    kc.check_sparse_layer_case(name)


def test_mednext_encoder_through_sparse_encoder():
    """A small MedNeXt head (P/MedNeXt_head.py) converted by SparseEncoder runs forward/backward on the sm_100a kernels with
    activation checkpointing outside the blocks (checkpoint_style='outside_block'), features zero on masked patches."""
    from anatomask_b200 import encoder3D as enc, MedNeXt_head as mh
    torch.manual_seed(0)
    head = mh.MedNeXt(1, 16, 1, exp_r=2, kernel_size=3, do_res=True, do_res_up_down=True, checkpoint_style='outside_block',
                      block_counts=[1] * 9)
    se = enc.SparseEncoder(head, input_size=(32, 32, 32)).cuda()
    assert se.downsample_ratio == 16 and se.enc_feat_map_chs == [16, 32, 64, 128, 256]
    active = kc._rand_mask(2, 2, keep=0.5, seed=3).cuda()
    enc._cur_active = active
    inp = torch.randn(2, 1, 32, 32, 32, device='cuda') * kc._up(active, 32)
    feats = se(inp)
    assert [tuple(f.shape[1:]) for f in feats] == [(16, 32, 32, 32), (32, 16, 16, 16), (64, 8, 8, 8), (128, 4, 4, 4), (256, 2, 2, 2)]
    sum(f.float().square().mean() for f in feats).backward()
    torch.cuda.synchronize()
    for f, s in zip(feats, (32, 16, 8, 4, 2)):
        assert float(f.float().abs().mul((~kc._up(active, s)).float()).max()) == 0.0
    for n, p in se.named_parameters():
        if 'dummy_tensor' not in n:
            assert p.grad is not None and torch.isfinite(p.grad).all(), n


def test_gradient_checkpointed_variants():
    """P/GC.py: checkpointed STUNet stages / decoder blocks — same loss and gradients as the regular mirror, less memory"""
    mc.check_gradient_checkpointing()


def test_device_side_input_pipeline():
    """crop / pad + order-3 spline rotation / scaling + mirroring on the GPU against the numpy / scipy
    restatement of the loader + batchgenerators transforms, forced branches and seeded random batches"""
    kc.check_augment()


def test_device_side_input_pipeline_full_size():
    kc.check_augment_full_size()


def test_checkpoint_write_resume_roundtrip():
    """`_head_latest.pt` writer + resume on the GPU — same masks, step count, moments and teacher after resume"""
    mc.check_checkpoint_resume()


@pytest.mark.parametrize('name', ['S64', 'B64'])
def test_engine_lean_zero_mode_with_poisoned_allocations(name):
    """ops.LEAN_ZERO (shell zeroing + in-place shortcut gradient) against the full-zero-fill path, masked allocations poisoned with NaN"""
    mc.check_lean_zero(name)


@pytest.mark.parametrize('kw', [dict(Cc=32, S=32, f=2), dict(Cc=64, S=32, f=4), dict(Cc=16, S=16, f=4, N=3), dict(Cc=512, S=8, f=8)])
def test_zero_shell_of_visible_patches(kw):
    kc.check_zero_shell(**kw)


@pytest.mark.parametrize('kw', [dict(Cc=32, S=32, f=2), dict(Cc=64, S=16, f=4), dict(Cc=128, S=16, masked=False), dict(Cc=24, S=16, f=2)])
def test_add_parity0(kw):
    kc.check_add_parity0(**kw)


@pytest.mark.parametrize('kw', [dict(Cin=32, Cout=64, S=32, f=2), dict(Cin=64, Cout=128, S=32, f=4), dict(Cin=128, Cout=256, S=16, f=4),
                                dict(Cin=32, Cout=32, S=32, f=2, stride=1)])
def test_conv_pair_matches_separate_convs(kw):
    kc.check_conv_pair(**kw)
