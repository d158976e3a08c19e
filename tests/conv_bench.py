"""Isolated conv-kernel timings (CUDA events, L2 flushed between iterations) for the layer shapes of STUNet-B @128³.
   python tests/conv_bench.py [v1|v2|all]      → TFLOP/s per layer and pass (fwd / dgrad / wgrad)"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anatomask_b200 import ops, _lib as L  # noqa: E402

bf16 = torch.bfloat16
LAYERS = [  # name, Cin, Cout, S, N
    ('dec3.conv0 64->64 @128', 64, 64, 128, 2), ('dec3.conv3 64->32 @128', 64, 32, 128, 2),
    ('dec2.conv0 128->128 @64', 128, 128, 64, 2), ('dec2.conv3 128->64 @64', 128, 64, 64, 2),
    ('dec1.conv0 256->256 @32', 256, 256, 32, 2), ('dec0.conv0 512->512 @16', 512, 512, 16, 2),
    ('enc0.conv2 32->32 @128', 32, 32, 128, 1),
]


def time_it(fn, flush, iters=5):
    fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return sorted(ms)[len(ms) // 2]


def main(which):
    dev = torch.device('cuda:0')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # T1 = one accumulator per CTA iteration (dependent-MMA chain), T = interleaved accumulators (default)
    impls = {'v1': (L.IMPL_TCGEN05_V1, None), 'v3': (L.IMPL_TCGEN05, None)}
    if which == 'wgrad':
        impls = {}
    elif which != 'all':
        impls = {which: impls[which]}
    out = []
    for name, ci, co, S, N in LAYERS[:int(os.environ.get('AMB_CB_LAYERS', len(LAYERS)))]:
        x = torch.randn(N, S, S, S, ci, device=dev).to(bf16)
        dy = torch.randn(N, S, S, S, co, device=dev).to(bf16)
        w = torch.randn(co, ci, 3, 3, 3, device=dev) / (27 * ci) ** 0.5
        wf = ops._pack(w, 27, co, ci, 1, ci * 27, 27)
        wd = ops._pack(w, 27, ci, co, 1, 27, ci * 27)
        y = torch.empty(N, S, S, S, co, dtype=bf16, device=dev)
        dx = torch.empty_like(x)
        dw = torch.zeros(27, co, ci, device=dev)
        flops = 2.0 * N * S ** 3 * 27 * ci * co
        row = {'layer': name}
        for tag, (impl, tenv) in impls.items():
            if tenv is None:
                os.environ.pop('AMB_IGEMM_T', None)
            else:
                os.environ['AMB_IGEMM_T'] = tenv
            t = time_it(lambda: ops._conv_call(L.OP_CONV, impl, (N, S, S, S), ci, co, 3, 1, x, y, wf), flush)
            row[f'fwd_{tag}'] = round(flops / t / 1e9, 1)
            t = time_it(lambda: ops._conv_call(L.OP_CONV_DGRAD, impl, (N, S, S, S), ci, co, 3, 1, dy, dx, wd), flush)
            row[f'dgrad_{tag}'] = round(flops / t / 1e9, 1)
        a = L.WgradArgs(L.OP_CONV, L.IMPL_TCGEN05, N, S, S, S, ci, co, 3, 1, x.data_ptr(), dy.data_ptr(), dw.data_ptr(),
                        1, 1, 1, 0, 0, torch.cuda.current_stream().cuda_stream)
        t = time_it(lambda: L.call('amb_conv_wgrad', C.byref(a)), flush)
        row['wgrad'] = round(flops / t / 1e9, 1)
        print(json.dumps(row), flush=True)
        out.append(row)
    # ConvTranspose k4 s2 (decoder upsampling): 64 -> 64, 64^3 -> 128^3 and 128 -> 128, 32^3 -> 64^3
    for name, c, S, N in [('dec3.up 64->64 @64->128', 64, 64, 2), ('dec2.up 128->128 @32->64', 128, 32, 2)]:
        x = torch.randn(N, S, S, S, c, device=dev).to(bf16)
        w = torch.randn(c, c, 4, 4, 4, device=dev) / (8 * c) ** 0.5
        wp = ops._pack(w, 64, c, c, 1, 64, c * 64)
        y = torch.empty(N, 2 * S, 2 * S, 2 * S, c, dtype=bf16, device=dev)
        flops = 2.0 * N * S ** 3 * 64 * c * c
        t = time_it(lambda: ops._conv_call(L.OP_CONVT, L.IMPL_TCGEN05, (N, S, S, S), c, c, 4, 2, x, y, wp), flush)
        row = {'layer': name, 'convT_fwd': round(flops / t / 1e9, 1)}
        dy = torch.randn(N, 2 * S, 2 * S, 2 * S, c, device=dev).to(bf16)
        dw = torch.zeros(64, c, c, device=dev)
        a = L.WgradArgs(L.OP_CONVT, L.IMPL_TCGEN05, N, S, S, S, c, c, 4, 2, x.data_ptr(), dy.data_ptr(), dw.data_ptr(),
                        1, 1, 1, 0, 0, torch.cuda.current_stream().cuda_stream)
        t = time_it(lambda: L.call('amb_conv_wgrad', C.byref(a)), flush)
        row['convT_wgrad'] = round(flops / t / 1e9, 1)
        print(json.dumps(row), flush=True)
        out.append(row)
    return out


if __name__ == '__main__' and not (len(sys.argv) > 1 and sys.argv[1] == 'sweep'):
    main(sys.argv[1] if len(sys.argv) > 1 else 'all')


def mask_sweep():
    """BASELINE config 5 / SURVEY §8(d) C5: single masked conv layers at mask ratios 0.0 … 0.9 — the active-tile kernels
    (work-list: masked tiles are never computed) against (a) the same tcgen05 kernel walking every tile densely with the
    mask applied in the epilogue and (b) what the reference does: a dense cuDNN convolution followed by a mask multiply
    (P/encoder3D.py:12-15), here torch F.conv3d in bf16 channels-last on the same GPU.   python tests/conv_bench.py sweep"""
    import torch.nn.functional as F
    dev = torch.device('cuda:0')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    layers = [('32->32 k3 s1 @128^3', 32, 32, 128, 1, 8), ('64->64 k3 s1 @64^3', 64, 64, 64, 1, 8), ('128->128 k3 s1 @32^3', 128, 128, 32, 1, 8),
              ('32->64 k3 s2 128^3->64^3', 32, 64, 128, 2, 8), ('64->128 k3 s2 64^3->32^3', 64, 128, 64, 2, 8)]
    N = 2
    rows = []
    for name, ci, co, S, stride, f in layers:
        So = S // stride
        x = torch.randn(N, S, S, S, ci, device=dev).to(bf16)
        w = torch.randn(co, ci, 3, 3, 3, device=dev) / (27 * ci) ** 0.5
        wf = ops._pack(w, 27, co, ci, 1, ci * 27, 27)
        y = torch.empty(N, So, So, So, co, dtype=bf16, device=dev)
        dense_flops = 2.0 * N * So ** 3 * 27 * ci * co
        xt = x.permute(0, 4, 1, 2, 3)                                    # logical NCDHW view of channels-last memory
        wt = w.to(bf16).contiguous(memory_format=torch.channels_last_3d)
        Lp = f ** 3
        for ratio in [0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9]:
            keep = max(1, round(Lp * (1 - ratio)))
            g = torch.Generator().manual_seed(int(ratio * 10))
            idx = torch.rand(N, Lp, generator=g).argsort(1)[:, :keep]
            mask = torch.zeros(N, Lp, dtype=torch.bool).scatter_(1, idx, True).view(N, 1, f, f, f).to(dev)
            m = ops.MaskCtx(mask)
            up = mask.repeat_interleave(So // f, 2).repeat_interleave(So // f, 3).repeat_interleave(So // f, 4).to(bf16)
            t_list = time_it(lambda: ops._conv_call(L.OP_CONV, L.IMPL_TCGEN05, (N, S, S, S), ci, co, 3, stride, x, y, wf, None, m,
                                                    sparse=True), flush)
            t_dense = time_it(lambda: ops._conv_call(L.OP_CONV, L.IMPL_TCGEN05, (N, S, S, S), ci, co, 3, stride, x, y, wf, None, m,
                                                     sparse=False), flush)
            t_cudnn = time_it(lambda: F.conv3d(xt, wt, None, stride, 1).mul_(up), flush)
            act = keep / Lp
            row = {'layer': name, 'mask_ratio': ratio, 'active_frac': round(act, 3), 'ms_active_tile': round(t_list, 4),
                   'ms_dense_walk': round(t_dense, 4), 'ms_cudnn_dense_plus_mask': round(t_cudnn, 4),
                   'tflops_algorithmic_active_tile': round(dense_flops * act / t_list / 1e9, 1),
                   'speedup_vs_dense_walk': round(t_dense / t_list, 2), 'speedup_vs_cudnn': round(t_cudnn / t_list, 2)}
            print(json.dumps(row), flush=True)
            rows.append(row)
    return rows


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'sweep':
    mask_sweep()
