"""Per-kernel table of ONE step (torch.profiler / CUPTI kernel records, no replay overhead): every kernel of the step — ours
and torch's — with launch count, total µs and share, for (a) the eager single-stream pass and (b) the CUDA-graph replay
(two streams: the weight-gradient branch overlaps the main chain).  Writes gpurun_out/<tag>_{eager,graph}.txt (aggregate
table followed by the launch sequence: stream, start µs, duration µs, name).

  python tests/step_profile.py [tag] [--model B] [--size 128] [--batch 2]
"""
import argparse
import collections
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anatomask_b200 import ops  # noqa: E402
from anatomask_b200.trainer import PretrainEngine, build_model  # noqa: E402


def short(name: str) -> str:
    name = re.sub(r'^void ', '', name)
    name = re.sub(r'\(.*$', '', name)
    name = re.sub(r'at::native::', '', name)
    return name[:110]


def table(prof, path, header):
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and 'mem' not in e.name.lower()[:3]]
    rows = []
    for e in evs:
        rows.append((e.time_range.start, e.time_range.end - e.time_range.start, short(e.name), getattr(e, 'device_index', 0)))
    rows.sort()
    mem = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.name.lower().startswith('mem')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for _, d, n, _ in rows:
        agg[n][0] += 1
        agg[n][1] += d
    tot = sum(v[1] for v in agg.values())
    span = (max(r[0] + r[1] for r in rows) - rows[0][0]) if rows else 0.0
    with open(path, 'w') as f:
        f.write(f'# {header}\n# kernels {len(rows)}  sum of kernel durations {tot / 1e3:.3f} ms  first-to-last span {span / 1e3:.3f} ms'
                f'  memcpy/memset records {len(mem)} ({sum(e.time_range.end - e.time_range.start for e in mem) / 1e3:.3f} ms)\n')
        for n, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'{v[1] / 1e3:9.3f} ms {v[0]:5d} {100 * v[1] / tot:5.1f}%  {n}\n')
        f.write('# sequence: start_us dur_us name\n')
        t0 = rows[0][0] if rows else 0
        for s, d, n, _ in rows:
            f.write(f'{s - t0:10.1f} {d:8.1f} {n}\n')
    return tot, span


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('tag', nargs='?', default='step_profile')
    ap.add_argument('--model', default='B')
    ap.add_argument('--size', type=int, default=128)
    ap.add_argument('--batch', type=int, default=2)
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    torch.manual_seed(1234)
    S = a.size
    model = build_model(a.model, (S, S, S), anatomask=True)
    eng = PretrainEngine(model, lr=1e-4, epochs=1000, anatomask=True, mask_rng='device')
    inp = torch.randn(a.batch, 1, S, S, S, device=dev)
    os.makedirs('gpurun_out', exist_ok=True)
    from torch.profiler import profile, ProfilerActivity
    for _ in range(3):
        eng.graph_step(inp, 500)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as pg:
        eng.graph_step(inp, 500)
        torch.cuda.synchronize()
    tot, span = table(pg, f'gpurun_out/{a.tag}_graph.txt', f'graph replay, STUNet-{a.model} {a.batch}x{S}^3')
    print(f'graph replay: kernel sum {tot / 1e3:.3f} ms, span {span / 1e3:.3f} ms')
    ops.NO_SIDE = True
    for _ in range(2):
        eng.device_step(inp, 500)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as pe:
        eng.device_step(inp, 500)
        torch.cuda.synchronize()
    ops.NO_SIDE = False
    tot, span = table(pe, f'gpurun_out/{a.tag}_eager.txt', f'eager single-stream pass, STUNet-{a.model} {a.batch}x{S}^3')
    print(f'eager pass: kernel sum {tot / 1e3:.3f} ms, span {span / 1e3:.3f} ms')


if __name__ == '__main__':
    main()
