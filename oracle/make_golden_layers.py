"""ORACLE — TEST INFRASTRUCTURE ONLY.  Generates tests/golden/sparse_layers.pt by executing the UNMODIFIED reference
classes of P/encoder3D.py (SparseGroupNorm, SparseConvNeXtLayerNorm, SparseMaxPooling, SparseAvgPooling,
SparseAdaptiveAvgPooling, SparseConvNeXtBlock, depthwise SparseConv3d) and P/MedNeXt_head.py blocks converted by the
reference's own SparseEncoder.dense_model_to_sparse — forward outputs and gradients on the seeded cases of
oracle/sparse_layers_port.py (inputs, parameters and output gradients are regenerated from the seeds; the fixture holds
digests of the reference's outputs / input gradients and its parameter gradients).

Run in the build container only (needs /root/reference):   python oracle/make_golden_layers.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/nnunetv2/training/nnUNetTrainer/variants/pretrain'
sys.path[:0] = [os.path.join(HERE, 'timm_stub'), REF, os.path.dirname(HERE)]

from oracle import sparse_layers_port as sl  # noqa: E402


def digest(t, n=2048):
    flat = t.detach().double().flatten()
    idx = torch.linspace(0, flat.numel() - 1, min(n, flat.numel())).long()
    return dict(shape=tuple(t.shape), norm=float(flat.norm()), sum=float(flat.sum()), idx=idx, val=flat[idx].float())


def build_reference_module(name):
    """The reference module of a case, constructed (and converted) by the reference's own code."""
    import encoder3D as enc
    import MedNeXt_head as mh
    kind, opt = sl.CASES[name]
    C = sl.CASE_C
    if kind == 'group_norm':
        return enc.SparseGroupNorm(opt['groups'], C, eps=1e-5)
    if kind == 'layer_norm':
        return enc.SparseConvNeXtLayerNorm(C, eps=1e-6, data_format=opt['fmt'], sparse=opt.get('sparse', True))
    if kind == 'pool':
        if opt['mode'] == 'max':
            return enc.SparseMaxPooling(opt['k'], opt['s'], opt['p'])
        return enc.SparseAvgPooling(opt['k'], opt['s'], opt['p'], count_include_pad=opt.get('include_pad', True))
    if kind == 'adaptive_avg':
        return enc.SparseAdaptiveAvgPooling((1, 1, 1))
    if kind == 'dwconv':
        return enc.SparseEncoder.dense_model_to_sparse(torch.nn.Conv3d(C, C, opt['k'], opt['s'], opt['k'] // 2, groups=C))
    if kind == 'convnext':
        m = enc.SparseConvNeXtBlock(C, drop_path=0., layer_scale_init_value=0.5, sparse=True, ks=7)
        return enc.SparseEncoder.dense_model_to_sparse(m) if opt['converted'] else m
    if kind == 'mednext':
        blk = mh.MedNeXtDownBlock(C, 2 * C, exp_r=2, kernel_size=3, do_res=True, norm_type='group') if opt['down'] else \
            mh.MedNeXtBlock(C, C, exp_r=2, kernel_size=3, do_res=True, norm_type='group')
        return enc.SparseEncoder.dense_model_to_sparse(blk)
    raise KeyError(kind)


def run_reference(name):
    import encoder3D as enc
    m = build_reference_module(name)
    x, active, g = sl.case_inputs(name)
    params = sl.case_params({k: tuple(v.shape) for k, v in m.named_parameters()}, g)
    for k, p in m.named_parameters():
        p.data.copy_(params[k])
    enc._cur_active = active
    xr = x.clone().requires_grad_(True)
    y = m(xr)
    dy = sl.case_dy(y.shape, name)
    y.backward(dy)
    return dict(y=y.detach(), dx=xr.grad, grads={k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None},
                param_shapes={k: tuple(v.shape) for k, v in m.named_parameters()})


def main():
    out = {}
    for name in sl.CASES:
        r = run_reference(name)
        out[name] = dict(y=digest(r['y']), dx=digest(r['dx']), grads=r['grads'], param_shapes=r['param_shapes'])
    path = os.path.join(HERE, '..', 'tests', 'golden', 'sparse_layers.pt')
    torch.save(out, path)
    print({k: v['y']['shape'] for k, v in out.items()}, os.path.getsize(path))


if __name__ == '__main__':
    main()
