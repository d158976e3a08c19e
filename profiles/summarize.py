"""Turns the raw ncu outputs a gpurun call brought back (gpurun_out/) into the tracked summaries of this directory.

  python profiles/summarize.py launches gpurun_out/r2f_launches.csv profiles/r2_launches_summary.md
  python profiles/summarize.py hbm      gpurun_out/r2f_hbm.ncu-rep  profiles/r2_hbm_kernels_summary.md

`launches`: the CSV of `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_cycles_active... --clock-control none` over ONE eager single-stream step (tests/ncu_target_step.py) ->
per-kernel launches, total time, share of the step, DRAM bytes, achieved DRAM GB/s, tensor-pipe % (time-weighted).
`hbm`: the section capture (`--section SpeedOfLight --section MemoryWorkloadAnalysis ...`) of the HBM-bound kernels of that
step -> for every kernel its LARGEST launch: duration, DRAM read + write, achieved GB/s against the measured HBM peak.
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    PEAK = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    PEAK = 6537.6


def short(name):
    name = re.sub(r'^void ', '', name)
    name = re.sub(r'\(.*$', '', name)
    name = name.replace('amb::', '').replace('at::native::', '')
    name = re.sub(r'^(at::)?vectorized_elementwise_kernel<(\d+), (at::)?(native::)?(\w+)<([^>]*)>.*', r'torch \5<\6> (vec\2)', name)
    name = re.sub(r'^(at::)?unrolled_elementwise_kernel<(at::)?(native::)?(\w+).*', r'torch \4 (unrolled)', name)
    name = re.sub(r'^(at::)?elementwise_kernel<.*?(\w+Functor).*', r'torch \2 (elementwise)', name)
    return name[:90]


def to_ms(v, unit):
    v = float(v.replace(',', ''))
    return v / 1e6 if unit in ('ns', 'nsecond') else v / 1e3 if unit in ('us', 'usecond') else v if unit in ('ms', 'msecond') else v * 1e3


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    start = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    h = rows[start]
    ki, vi, ui, mi, ii = (h.index(k) for k in ('Kernel Name', 'Metric Value', 'Metric Unit', 'Metric Name', 'ID'))
    per = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        d = per.setdefault(r[ii], {'name': short(r[ki])})
        m = r[mi]
        if m == 'gpu__time_duration.sum':
            d['ms'] = to_ms(r[vi], r[ui])
        elif m.startswith('dram__bytes_read'):
            d['rd'] = to_bytes(r[vi], r[ui])
        elif m.startswith('dram__bytes_write'):
            d['wr'] = to_bytes(r[vi], r[ui])
        elif m.startswith('sm__pipe_tensor'):
            d['tc'] = float(r[vi].replace(',', ''))
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for d in per.values():
        a = agg[d['name']]
        a[0] += 1
        a[1] += d.get('ms', 0.0)
        a[2] += d.get('rd', 0.0) + d.get('wr', 0.0)
        a[3] += d.get('tc', 0.0) * d.get('ms', 0.0)
    tot = sum(a[1] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if not k.startswith('torch'))
    with open(dst, 'w') as f:
        f.write(f'# Launch list of ONE eager single-stream AnatoMask step (STUNet-B, 2x128^3), ncu `--clock-control none`\n\n'
                f'Source: `{os.path.basename(src)}` (`tests/gpu_scripts/r2_evidence2.sh`, `tests/ncu_target_step.py`; one B200 under gpurun).  '
                f'{len(per)} launches, {tot:.2f} ms of kernel time serialised by the profiler (cold caches, so the SHARES are the evidence, '
                f'not the absolute times; the graph replay of the same launches takes the `ms_per_step` of the bench line).  '
                f'Kernels of this repo: {100 * ours / tot:.1f} % of the time; the rest are torch fills / copies / autograd adds.  '
                f'GB/s = (dram read + write) / time; measured HBM peak {PEAK:.0f} GB/s.\n\n'
                f'| kernel | launches | ms | share | DRAM GB | GB/s | tensor pipe % (time-weighted) |\n|---|---|---|---|---|---|---|\n')
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            gbs = a[2] / 1e9 / (a[1] / 1e3) if a[1] > 0 else 0.0
            tc = a[3] / a[1] if a[1] > 0 else 0.0
            f.write(f'| `{k}` | {a[0]} | {a[1]:.3f} | {100 * a[1] / tot:.1f} % | {a[2] / 1e9:.2f} | {gbs:.0f} | {tc:.1f} |\n')
    print(f'{dst}: {len(per)} launches, {tot:.2f} ms')


def hbm(src, dst):
    metrics = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes.sum.per_second',
               'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
               'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
               'launch__block_size']
    out = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    col = {m: h.index(m) for m in metrics if m in h}
    ki = h.index('Kernel Name')
    best = {}
    for r in rows[2:]:
        name = short(r[ki])
        ms = to_ms(r[col['gpu__time_duration.sum']], units[col['gpu__time_duration.sum']])
        if name not in best or ms > best[name][0]:
            best[name] = (ms, r)
    count = collections.Counter(short(r[ki]) for r in rows[2:])
    with open(dst, 'w') as f:
        f.write(f'# HBM-bound kernels of one AnatoMask step (STUNet-B, 2x128^3): ncu section capture, largest launch of each kernel\n\n'
                f'Source: `{os.path.basename(src)}` (`ncu --profile-from-start off --clock-control none --section SpeedOfLight --section '
                f'MemoryWorkloadAnalysis --section Occupancy --section LaunchStats -k regex:...` over `tests/ncu_target_step.py`; one B200 '
                f'under gpurun; read with `ncu -i ... --page raw --csv`).  achieved GB/s = (dram__bytes_read.sum + dram__bytes_write.sum) / '
                f'gpu__time_duration.sum, against the measured HBM peak of {PEAK:.0f} GB/s (MEASURED_PEAKS.json; nominal ~8 000).  '
                f'ncu replays every launch with cold caches, so small launches read lower than in the step.\n\n'
                f'| kernel | launches in the step | largest launch µs | DRAM MB (read [+ write]) | DRAM write MB (0 = included left) | achieved GB/s | % of measured peak | '
                f'dram__throughput % | sm__throughput % | warps active % | regs | grid x block |\n|---|---|---|---|---|---|---|---|---|---|---|---|\n')
        for name, (ms, r) in sorted(best.items(), key=lambda kv: -kv[1][0]):
            g = lambda m: r[col[m]] if m in col else ''
            if 'dram__bytes_read.sum' in col:
                rd = to_bytes(g('dram__bytes_read.sum'), units[col['dram__bytes_read.sum']])
                wr = to_bytes(g('dram__bytes_write.sum'), units[col['dram__bytes_write.sum']])
                gbs = (rd + wr) / 1e9 / (ms / 1e3)
            else:                                   # section capture: ncu's own rate (read + write bytes per second)
                scale = {'Tbyte/s': 1e3, 'Gbyte/s': 1.0, 'Mbyte/s': 1e-3}[units[col['dram__bytes.sum.per_second']]]
                gbs = float(g('dram__bytes.sum.per_second').replace(',', '')) * scale
                rd, wr = gbs * 1e9 * (ms / 1e3), 0.0
            f.write(f'| `{name}` | {count[name]} | {ms * 1e3:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {gbs:.0f} | {100 * gbs / PEAK:.0f} % | '
                    f'{g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")} | {g("sm__throughput.avg.pct_of_peak_sustained_elapsed")} | '
                    f'{g("sm__warps_active.avg.pct_of_peak_sustained_active")} | {g("launch__registers_per_thread")} | '
                    f'{g("launch__grid_size")} x {g("launch__block_size")} |\n')
    print(f'{dst}: {len(best)} kernels')


if __name__ == '__main__':
    {'launches': launches, 'hbm': hbm}[sys.argv[1]](sys.argv[2], sys.argv[3])
