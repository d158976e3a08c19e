"""CPU-only checks: C-ABI library loads and exports every symbol of include/anatomask_b200.h; the host mirror keeps the
reference's checkpoint-key contract; arena / schedules / loud failure without a GPU."""
import os
import re

import pytest
import torch

from oracle import reference_port as rp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_is_built_from_the_sources_in_the_tree():
    """The build stamp (sha256 of csrc/ + header + flags) equals the hash of the sources: a stale library would make every GPU
    number describe another tree.  Rebuilds first if needed (nvcc cross-compiles sm_100a without a GPU)."""
    from anatomask_b200 import build as b
    b.build(force=False)
    assert os.path.exists(b.LIB) and b.built_hash() == b.source_hash() and not b.needs_build()


def test_library_exports_every_declared_symbol():
    from anatomask_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, 'include', 'anatomask_b200.h')).read()
    declared = set(re.findall(r'\b(amb_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations parsed'
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in the header but not exported'
    assert declared == set(_lib.EXPORTED), declared ^ set(_lib.EXPORTED)
    assert lib.amb_version() == 100 and lib.amb_sm_arch() == 100


@pytest.mark.parametrize('name', ['tiny', 'S64', 'B64'])
def test_state_dict_contract_matches_reference(name):
    from anatomask_b200.trainer import build_model
    cfg = rp.CONFIGS[name]
    model = build_model(base=cfg.base, depth=cfg.depth, input_size=cfg.input_size, device='cpu')
    sd = model.state_dict()
    shapes = rp.param_shapes(cfg)
    assert set(sd.keys()) == set(shapes.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == shapes[k][0], k
    model.load_state_dict(rp.make_state(cfg, 1))          # loads both ways
    assert (model.len_keep, model.fmap_h) == (cfg.len_keep, cfg.fmap[0])


def test_arena_views_and_dead_parameters():
    from anatomask_b200.trainer import build_model, ParamArena, dead_parameter_names
    cfg = rp.CONFIGS['tiny']
    model = build_model(base=cfg.base, input_size=cfg.input_size, device='cpu')
    before = {k: v.clone() for k, v in model.state_dict().items()}
    dead = dead_parameter_names(model)
    assert sorted(dead) == sorted(k for k in rp.param_shapes(cfg)
                                  if k.startswith(('densify_norms.4', 'densify_projs.4', 'mask_tokens.4')))
    arena = ParamArena(model, dead, with_grads=True)
    for k, v in model.state_dict().items():
        assert torch.equal(v, before[k]), k
    live = [n for n in rp.live_param_names(cfg)]
    assert arena.n_live >= sum(before[n].numel() for n in live)
    for n, p in model.named_parameters():
        assert (p.grad is None) == (n in dead), n
        o, k = arena.offsets[n]
        assert p.data_ptr() == arena.flat.data_ptr() + 4 * o
    arena.flat.mul_(2)
    assert torch.allclose(model.mask_tokens[0], 2 * before['mask_tokens.0'])


def test_schedules():
    from anatomask_b200.trainer import lr_at_epoch, ema_decay_at_epoch
    assert lr_at_epoch(0, 1e-4) == pytest.approx(1e-6)
    assert lr_at_epoch(20, 1e-4) == pytest.approx(1e-4)
    assert lr_at_epoch(1000, 1e-4) == pytest.approx(0.0, abs=1e-12)
    assert ema_decay_at_epoch(0, 1000) == 0.999 and ema_decay_at_epoch(500, 1000) == 0.9999
    assert ema_decay_at_epoch(125, 1000) == pytest.approx(rp.ema_decay(125, 1000))


def test_cpu_tensors_fail_loudly():
    from anatomask_b200.trainer import build_model
    cfg = rp.CONFIGS['tiny']
    model = build_model(base=cfg.base, input_size=cfg.input_size, device='cpu')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        model(torch.zeros(1, 1, 32, 32, 32), active_b1ff=torch.ones(1, 1, 2, 2, 2, dtype=torch.bool))


def test_unsupported_layers_raise_like_the_reference():
    from anatomask_b200 import encoder3D
    with pytest.raises(NotImplementedError):
        encoder3D.SparseEncoder.dense_model_to_sparse(torch.nn.Conv1d(1, 1, 1))


def test_checkpoint_keys_feed_the_reference_finetune_loader(tmp_path):
    """`_head_latest.pt` layout (P/pretrain.py:450-463) and the key mapping of load_stunet_ssl_weights
    (nnunetv2/run/load_pretrained_weights.py:76-79): encoder keys come out as conv_blocks_context.{s}.{b}.*"""
    from anatomask_b200.trainer import build_model
    from anatomask_b200 import checkpoint
    cfg = rp.CONFIGS['tiny']
    model = build_model(base=cfg.base, input_size=cfg.input_size, device='cpu')
    path = str(tmp_path / 'STUNet_B_head_latest.pt')
    checkpoint.save_head_checkpoint(path, model, epoch=3)
    ck = torch.load(path, weights_only=False)
    assert set(ck) >= {'network_weights', 'optimizer_state', 'grad_scaler_state', 'train_loss', 'val_loss', 'current_epoch'}
    assert all(k.startswith('module.') for k in ck['network_weights'])
    enc = checkpoint.encoder_state_for_finetune(ck['network_weights'])
    assert 'conv_blocks_context.0.0.conv1.weight' in enc and 'conv_blocks_context.4.0.norm2.bias' in enc
    assert all(k.startswith('conv_blocks_context.') for k in enc) and len(enc) == 50
    # a fine-tuning STUNet (same block naming) accepts them
    from anatomask_b200.STUNet_head import STUNet
    ft = STUNet(1, 1, depth=[1] * 6, dims=[cfg.base * x for x in (1, 2, 4, 8, 16, 16)],
                pool_op_kernel_sizes=[[2, 2, 2]] * 4 + [[1, 1, 1]], conv_kernel_sizes=[[3, 3, 3]] * 6)
    missing, unexpected = ft.load_state_dict(enc, strict=False)
    assert not unexpected and not missing


def test_struct_layouts_match_the_c_header(tmp_path):
    """ctypes mirrors (and the numpy job table of ops.PackPlan) against sizeof / offsetof taken from include/*.h by gcc."""
    import ctypes as C
    import shutil
    import subprocess
    import numpy as np
    from anatomask_b200 import _lib as L
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('no gcc')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    structs = {'amb_conv_args': L.ConvArgs, 'amb_wgrad_args': L.WgradArgs, 'amb_geo': L.Geo}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include <stdint.h>', '#include "anatomask_b200.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf(" %zu", offsetof({cname}, {fname}));')
        lines.append('  printf("\\n");')
    job_fields = ['src', 'dst', 'T', 'A', 'B', 'b_fast', 'tile_begin', 'tiles_b', 'tchunks']
    lines.append('  printf("amb_pack_job %zu", sizeof(amb_pack_job));')
    lines += [f'  printf(" %zu", offsetof(amb_pack_job, {f}));' for f in job_fields]
    lines += ['  printf("\\n");', '  return 0;', '}']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.run([gcc, '-I', os.path.join(root, 'include'), str(src), '-o', str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    got = {ln.split()[0]: [int(v) for v in ln.split()[1:]] for ln in out}
    for cname, cls in structs.items():
        want = [C.sizeof(cls)] + [getattr(cls, f).offset for f, _ in cls._fields_]
        assert got[cname] == want, (cname, got[cname], want)
    dt = np.dtype([('src', np.uint64), ('dst', np.uint64), ('T', np.int32), ('A', np.int32), ('B', np.int32),
                   ('b_fast', np.int32), ('tile_begin', np.int32), ('tiles_b', np.int32), ('tchunks', np.int32),
                   ('pad', np.int32)])
    assert got['amb_pack_job'] == [dt.itemsize] + [dt.fields[f][1] for f in job_fields]


def test_pack_plan_job_table():
    """ops.PackPlan: duplicate requests collapse, layouts outside the tiled kernel are left to the per-call pack, tile
    ranges are contiguous — checked on CPU tensors (the launch itself is GPU-only)."""
    import numpy as np
    from anatomask_b200 import ops
    w1 = torch.zeros(64, 32, 27)          # conv (Cout, Cin, taps)
    w2 = torch.zeros(16, 24, 64)          # ConvTranspose (Cin, Cout, taps)
    w3 = torch.zeros(8, 8, 27)
    rec = []
    def ask(w, T, A, B, sa, sb, st=1):
        rec.append(((w.data_ptr(), T, A, B, sa, sb, st), w))
    ask(w1, 27, 64, 32, 32 * 27, 27)      # forward form  [t][co][ci]: b (= ci) is the fast parameter axis
    ask(w1, 27, 32, 64, 27, 32 * 27)      # dgrad form    [t][ci][co]: a (= ci) is the fast axis
    ask(w1, 27, 64, 32, 32 * 27, 27)      # asked again by the next step: same key
    ask(w2, 64, 24, 16, 64, 24 * 64)      # ConvTranspose forward [t][co][ci] from (ci, co, t)
    ask(w3, 27, 8, 8, 5, 7)               # strides the tiled kernel does not cover
    plan = ops.PackPlan(rec)
    assert plan.n_jobs == 3 and len(plan.cache) == 3
    dt = np.dtype([('src', np.uint64), ('dst', np.uint64), ('T', np.int32), ('A', np.int32), ('B', np.int32),
                   ('b_fast', np.int32), ('tile_begin', np.int32), ('tiles_b', np.int32), ('tchunks', np.int32),
                   ('pad', np.int32)])
    tab = plan.table.numpy().view(dt)
    assert list(tab['b_fast']) == [1, 0, 0]
    # tiles: 4x64 positions (b fast) or 16x16, x ceil(T / 32) tap chunks
    tiles = [16 * 1 * 1, 2 * 4 * 1, 2 * 1 * 2]
    assert list(tab['tile_begin']) == [0, tiles[0], tiles[0] + tiles[1]] and plan.total_tiles == sum(tiles)
    assert list(tab['tiles_b']) == [1, 4, 1] and list(tab['tchunks']) == [1, 1, 2]
    for (key, w), job in zip([rec[0], rec[1], rec[3]], tab):
        assert int(job['src']) == w.data_ptr() and int(job['dst']) == plan.cache[key].data_ptr()
        assert tuple(plan.cache[key].shape) == (key[1], key[2], key[3]) and plan.cache[key].dtype == torch.bfloat16
    # groups (the engine: teacher weights on the main stream, student weights on the side stream): every group is a job table
    # of its own whose tile ranges start at 0, and together they cover the same jobs
    split = ops.PackPlan(rec, group_of=lambda w: 1 if w is w2 else 0)
    assert sorted(split.groups) == [0, 1] and split.n_jobs == 3 and split.total_tiles == sum(tiles) and split.table is None
    t0, n0, tiles0 = split.groups[0]
    t1, n1, tiles1 = split.groups[1]
    assert (n0, tiles0, n1, tiles1) == (2, tiles[0] + tiles[1], 1, tiles[2])
    assert list(t0.numpy().view(dt)['tile_begin']) == [0, tiles[0]] and list(t1.numpy().view(dt)['tile_begin']) == [0]
    assert int(t1.numpy().view(dt)['src'][0]) == w2.data_ptr()
    # taps-major master weights (the engine arena, fp32 [tap][Cout][Cin]): forward form = plain conversion (kind 2, blocks of
    # 2048 elements), input-gradient form = per-tap transpose (kind 3, 32 a x 64 b tiles); a 1x1x1 weight in the stock
    # layout is the same memory order and takes the same two kinds
    w4 = torch.zeros(27, 64, 40)          # [tap][Cout][Cin]
    w5 = torch.zeros(48, 16)              # 1x1x1 conv, stock layout (Cout, Cin)
    rec2 = []
    def ask2(w, T, A, B, sa, sb, st):
        rec2.append(((w.data_ptr(), T, A, B, sa, sb, st), w))
    ask2(w4, 27, 64, 40, 40, 1, 64 * 40)  # forward: A = Cout, B = Cin
    ask2(w4, 27, 40, 64, 1, 40, 64 * 40)  # dgrad:   A = Cin,  B = Cout
    ask2(w5, 1, 48, 16, 16, 1, 1)
    ask2(w5, 1, 16, 48, 1, 16, 1)
    plan2 = ops.PackPlan(rec2)
    tab2 = plan2.table.numpy().view(dt)
    assert list(tab2['b_fast']) == [2, 3, 2, 3]
    tiles2 = [-(-27 * 64 * 40 // 2048), 27 * 1 * 2, 1, 1 * 1 * 1]
    assert list(tab2['tile_begin']) == [0, tiles2[0], tiles2[0] + tiles2[1], sum(tiles2[:3])] and plan2.total_tiles == sum(tiles2)
    assert list(tab2['tiles_b'][[1, 3]]) == [1, 1] and list(tab2['tchunks'][[1, 3]]) == [2, 1]


def test_taps_major_arena_keeps_the_module_contract():
    """ParamArena(taps_major=True): conv weights are stored [tap][Cout][Cin] but the module still sees the reference's shapes
    and values (state_dict round trip), every tensor starts 16-byte aligned, ops.packed_alias finds the kernels' view of a
    weight and of its gradient, and Adam-moment views follow the parameter's element order."""
    from anatomask_b200 import ops
    from anatomask_b200.trainer import build_model, ParamArena, dead_parameter_names, taps_major_weight_names
    cfg = rp.CONFIGS['tiny']
    model = build_model(base=cfg.base, input_size=cfg.input_size, device='cpu')
    ref = build_model(base=cfg.base, input_size=cfg.input_size, device='cpu')
    ref.load_state_dict(model.state_dict())
    before = {k: v.clone() for k, v in model.state_dict().items()}
    names = taps_major_weight_names(model)
    assert names and all(n.endswith('.weight') for n in names)
    assert not any('proj.weight' == n.split('dense_decoder.')[-1] for n in names)          # the 1-channel projection stays stock
    arena = ParamArena(model, dead_parameter_names(model), with_grads=True, taps_major=True)
    mods = dict(model.named_modules())
    for n, p in model.named_parameters():
        assert torch.equal(p.detach(), before[n]), n
        o, k = arena.offsets[n]
        assert o % 4 == 0, n
        if n in names:
            tr = isinstance(mods[n[:-len('.weight')]], torch.nn.ConvTranspose3d)
            al = ops.packed_alias(p.detach(), tr)
            assert al is not None and al.is_contiguous() and al.data_ptr() == arena.flat.data_ptr() + 4 * o, n
            T = p.shape[2] * p.shape[3] * p.shape[4]
            assert tuple(al.shape) == ((T, p.shape[1], p.shape[0]) if tr else (T, p.shape[0], p.shape[1]))
            # element (tap t, out channel co, in channel ci) of the alias is the same number the module sees
            co, ci, t = al.shape[1] - 1, al.shape[2] // 2, T // 2
            w5 = p.detach().reshape(p.shape[0], p.shape[1], T)
            assert float(al[t, co, ci]) == float(w5[ci, co, t] if tr else w5[co, ci, t])
            if p.grad is not None:
                assert ops.packed_alias(p.grad, tr).data_ptr() == arena.grad.data_ptr() + 4 * o
        elif p.grad is not None:
            assert p.grad.is_contiguous() and p.grad.data_ptr() == arena.grad.data_ptr() + 4 * o
    # a state_dict written by the arena'd model loads into a stock model and back, bit for bit
    ref.load_state_dict({k: v.clone() for k, v in model.state_dict().items()})
    for k, v in ref.state_dict().items():
        assert torch.equal(v, before[k]), k
    arena.flat.mul_(2)
    model.load_state_dict(ref.state_dict())
    for k, v in model.state_dict().items():
        assert torch.equal(v, before[k]), k
    # moment views: writing through the view of a parameter touches exactly that parameter's slots of the flat buffer
    mom = torch.zeros_like(arena.grad)
    vs = arena.views(mom)
    n0 = sorted(names)[0]
    vs[n0].fill_(1.0)
    o, k = arena.offsets[n0]
    assert float(mom.sum()) == k and float(mom[o:o + k].sum()) == k


def test_augmentation_draws_follow_the_reference_order():
    """anatomask_b200.augment draws (bbox, rotation, scale, mirror) consume a seeded numpy stream exactly like the oracle's
    restatement of the loader + batchgenerators transforms; the initial patch size is the reference's get_patch_size."""
    import numpy as np
    from anatomask_b200 import augment as A
    from oracle import augment_port as O
    aug = A.DeviceAugmenter()
    assert aug.initial_patch_size == (205, 205, 205)            # 128·(cos30° + sin30°) / 0.85
    assert A.DeviceAugmenter(patch_size=(64, 64, 64)).initial_patch_size == (102, 102, 102)
    shapes = [(220, 310, 290), (150, 200, 180), (100, 512, 512)]
    n_interp = 0
    for seed in range(40):
        r1, r2 = np.random.RandomState(seed), np.random.RandomState(seed)
        mine = aug.draw(shapes, r1)
        ref = O.draw_batch(r2, shapes, aug.initial_patch_size, aug.patch_size, aug.angle)
        for a, (lb, angles, sc, flips) in zip(mine, ref):
            assert a.bbox_lb == lb and a.mirror == flips
            assert (a.matrix is None) == (angles is None and sc is None)
            if a.matrix is not None:
                n_interp += 1
                coords = np.random.RandomState(1).standard_normal((3, 5))
                want = O.rotate_coords_3d(coords.copy(), *angles) if angles is not None else coords.copy()
                want = want * (sc if sc is not None else 1.0)
                assert np.allclose((coords.T @ a.matrix).T, want, atol=1e-12)
        assert np.array_equal(r1.get_state()[1], r2.get_state()[1])      # both consumed the stream identically
    assert n_interp > 10


def test_augmentation_bbox_bounds():
    import numpy as np
    from anatomask_b200 import augment as A
    rng = np.random.RandomState(0)
    for _ in range(200):
        lb = A.draw_bbox((100, 300, 210), (205, 205, 205), (128, 128, 128), rng)
        # need_to_pad = 77 (105 where the case is shorter than the patch): lower corner in [-need//2, shape + need//2 + need%2 - patch]
        assert -53 <= lb[0] <= 100 + 52 + 1 - 205 and -39 <= lb[1] <= 300 + 38 + 1 - 205 and -39 <= lb[2] <= 210 + 38 + 1 - 205
    locs = {1: np.array([[0, 50, 150, 100]]), 2: np.array([])}
    lb = A.draw_bbox((100, 300, 210), (205, 205, 205), (128, 128, 128), rng, force_fg=True, class_locations=locs)
    assert lb == (max(-53, 50 - 102), 150 - 102, max(-39, 100 - 102))
