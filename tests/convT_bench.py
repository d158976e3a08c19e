"""ConvTranspose3d(k4 s2 p1) forward: the kz-stacked halo-plane kernel (conv_igemm4t.cu) against the per-tap kernel — exact
agreement on integer-valued inputs, then isolated timings (CUDA events, L2 flushed) for the decoder's up-sampling layers.
   python tests/convT_bench.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anatomask_b200 import ops, _lib as L  # noqa: E402

bf16 = torch.bfloat16


def run(x, wp, b, co, impl):
    N, D, H, W, ci = x.shape
    y = torch.empty(N, 2 * D, 2 * H, 2 * W, co, dtype=bf16, device=x.device)
    ops._conv_call(L.OP_CONVT, impl, (N, D, H, W), ci, co, 4, 2, x, y, wp, b)
    return y, (L.load().amb_last_conv_kernel() or b'').decode()


def main():
    dev = torch.device('cuda:0')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = []
    for ci, co, dims, N in [(64, 64, (16, 16, 16), 2), (32, 16, (5, 20, 12), 1), (64, 32, (7, 18, 24), 2), (128, 64, (16, 16, 8), 1),
                            (64, 64, (64, 64, 64), 2), (128, 128, (32, 32, 32), 2)]:
        g = torch.Generator().manual_seed(ci + co)
        x = torch.randint(-3, 4, (N, *dims, ci), generator=g).to(bf16).to(dev)
        w = torch.randint(-2, 3, (ci, co, 4, 4, 4), generator=g).float().to(dev)
        b = torch.randint(-4, 5, (co,), generator=g).float().to(dev)
        wp = ops._pack(w, 64, co, ci, 1, 64, co * 64)
        y4, k4 = run(x, wp, b, co, L.IMPL_TCGEN05)
        y1, k1 = run(x, wp, b, co, L.IMPL_TCGEN05_V1)
        torch.cuda.synchronize()
        row = {'shape': f'{ci}->{co} {dims} N={N}', 'kernels': [k4, k1], 'exact': bool(torch.equal(y4, y1))}
        flops = 2.0 * N * dims[0] * dims[1] * dims[2] * 64 * ci * co
        for tag, impl in (('stacked', L.IMPL_TCGEN05), ('per_tap', L.IMPL_TCGEN05_V1)):
            ms = []
            for _ in range(5):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); run(x, wp, b, co, impl); e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            m = sorted(ms)[2]
            row[tag] = {'ms': round(m, 4), 'tflops': round(flops / m / 1e9, 1)}
        print('RESULT convT', json.dumps(row), flush=True)
        out.append(row)
    assert all(r['exact'] for r in out), [r['shape'] for r in out if not r['exact']]


if __name__ == '__main__':
    main()
