"""Headline benchmark: 128³ volumes/s of one STUNet-B AnatoMask pre-training step (BASELINE.json), per the driver contract.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one process per GPU; torchrun for N > 1)
  python bench.py --impl reference [...]                         the reference's CPU implementation (oracle port) on the
                                                                 host cores, same config/metric, bounded sample per step

A step = teacher forward (eval) → per-patch teacher loss → hard-mask top-k → student forward/backward → global-norm clip
+ AdamW → EMA teacher update, on a batch of 2 synthetic N(0,1) volumes of 1×128³ per GPU (weak scaling), epoch 500 of
1000 (len_loss = 76 hard patches).  `value`: inputs resident in HBM.  `e2e`: the same step through the public API with
the batch coming from pinned host memory every step and the loss read back to the host.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = '128^3 volumes/sec, STUNet-B AnatoMask pretraining step'
UNIT = 'volumes/s'
F_FWD_GF = 1778.0          # algorithmic GFLOP / volume forward, STUNet-B @128³, mask 0.6


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))), 'measured'
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '200', '-i', str(gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        busy = [v for v in sm if v > 0.5 * (max(mx) if mx else 1)] or sm
        return {'sm_mhz': busy[len(busy) // 2] if busy else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def run_reference(args):
    """Reference arm: the oracle port (plain PyTorch fp32 restatement of the reference modules) on the host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import reference_port as rp
    import numpy as np
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = rp.CONFIGS['B128']
    batch = 1                                       # bounded sample: ONE 128³ volume per step (the GPU arm uses 2/GPU)
    tr = rp.RefTrainer(cfg, rp.make_state(cfg, 0), lr=1e-4, epochs=1000, anatomask=True)
    np.random.seed(0)
    inp = rp.make_input(cfg, batch, 0)
    steps, warm = max(1, min(args.steps, 3)), max(0, min(args.warmup, 1))
    times = []
    for i in range(warm + steps):
        mask1 = rp.random_mask(cfg, batch, torch.Generator().manual_seed(i))
        t0 = time.time()
        tr.anatomask_step(inp, mask1, 500)
        dt = time.time() - t0
        if i >= warm:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    val = batch / (ms / 1e3)
    out = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
           'warmup': warm, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
           'dtype': 'f32', 'data': 'synthetic',
           'config': {'workload': 'STUNet-B AnatoMask step, 1x128^3 volumes, mask 0.6, epoch 500/1000 (len_loss 76)',
                      'batch_per_step': batch},
           'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                            'sample': f'{steps} AnatoMask step(s) of {batch} volume (1x128^3), oracle port = plain '
                                      f'PyTorch fp32 restatement of the reference modules, torch.set_num_threads({cores})'},
           'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(out))


def cpu_baseline_sample():
    from oracle import reference_port as rp
    import numpy as np
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = rp.CONFIGS['B128']
    tr = rp.RefTrainer(cfg, rp.make_state(cfg, 0), lr=1e-4, epochs=1000, anatomask=True)
    np.random.seed(0)
    inp = rp.make_input(cfg, 1, 0)
    mask1 = rp.random_mask(cfg, 1, torch.Generator().manual_seed(0))
    t0 = time.time()
    tr.anatomask_step(inp, mask1, 500)
    dt = time.time() - t0
    return {'value': 1.0 / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': f'1 AnatoMask step of 1 volume (1x128^3) with the oracle port (PyTorch fp32, {cores} threads): {dt:.1f} s'}


def _mark(msg):
    if os.environ.get('AMB_BENCH_VERBOSE'):
        print(f'[bench r{os.environ.get("RANK", "0")}] {msg} t={time.time():.1f}', file=sys.stderr, flush=True)


def run_ours(args):
    import torch.distributed as dist
    from anatomask_b200 import ops, _lib
    from anatomask_b200.trainer import PretrainEngine, build_model
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    assert torch.cuda.is_available(), 'bench.py (our arm) needs a B200: there is no CPU fallback'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    group = None
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
        group = dist.group.WORLD
    _mark('pg up')
    lib = _lib.load()
    B, S = args.batch, args.size
    torch.manual_seed(1234 + rank)
    model = build_model(args.model, (S, S, S), anatomask=True)
    if world > 1:                                    # identical initial weights on every rank (DDP broadcast)
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, 0)
    _mark('model built + broadcast')
    eng = PretrainEngine(model, lr=1e-4, epochs=1000, anatomask=True, mask_rng='device', process_group=group)
    inp = torch.randn(B, 1, S, S, S, device=dev)
    host = torch.randn(B, 1, S, S, S).pin_memory()
    epoch = 500

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # The step runs as ONE CUDA-graph launch (no host sync inside: device RNG, device-side scalars); --no-graph or a
    # failed capture falls back to eager launches of the same kernels.
    use_graph = not args.no_graph
    _mark('engine up')
    graph_err = None
    if use_graph:
        try:
            eng.graph_step(inp, epoch)
            torch.cuda.synchronize()
        except Exception as e:                                   # noqa: BLE001
            use_graph, graph_err = False, f'{type(e).__name__}: {e}'[:200]
            torch.cuda.synchronize()
    _mark(f'graph={use_graph} err={graph_err}')
    run = (lambda x: eng.graph_step(x, epoch)) if use_graph else (lambda x: eng.device_step(x, epoch))
    for _ in range(args.warmup):
        run(inp)
    barrier()
    # ---- timed region 1: inputs resident in HBM -----------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        run(inp)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    _mark(f'timed region done {ms_total:.1f} ms')
    clocks = sampler.stop() if sampler else None
    ms_step = ms_total / args.steps
    value = B * world / (ms_step / 1e3)
    # ---- the same K steps launched eagerly with a CUDA-event pair around every conv launch: per-kernel durations for
    #      the roofline and the launch count (a graph replay issues exactly these launches) ---------------------------
    #      (single stream for this pass: with the weight-gradient branch on its side stream two kernels share the SMs and
    #      an event pair would time both)
    ops.PROFILE = []
    ops.NO_SIDE = True
    lib.amb_reset_launch_count()
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(args.steps):
        eng.device_step(inp, epoch)           # the launch sequence the graph captured (batched weight pack, device scalars)
    p1.record()
    barrier()
    ms_eager_total = p0.elapsed_time(p1)
    launches = int(lib.amb_launch_count())
    prof, ops.PROFILE = ops.PROFILE, None
    ops.NO_SIDE = False
    # ---- roofline of the dominant kernel family (live CUDA-event timing of every launch in the timed region) ----
    fam = {}
    for kind, flops, a, b in prof:
        d = fam.setdefault(kind, [0.0, 0.0, 0])
        d[0] += flops; d[1] += a.elapsed_time(b); d[2] += 1
    if os.environ.get('AMB_BENCH_DUMP') and rank == 0:           # per-launch table of the last profiled step
        n_per = len(prof) // args.steps
        rows = [(k, f, a.elapsed_time(b)) for k, f, a, b in prof[-n_per:]]
        for i, (k, f, ms) in enumerate(rows):
            print(f'#LAUNCH {i:3d} {k:12s} {f / 1e9:9.1f} GF {ms:7.3f} ms {f / ms / 1e9:8.1f} TF/s', file=sys.stderr)
    peaks, peak_src = _peaks()
    try:        # DRAM bytes of the family's dominant launch shape from the committed ncu --set full capture
        traffic = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', 'r1_ncu_traffic.json')))
    except (OSError, ValueError):
        traffic = {}
    peak = float(peaks.get('bf16_tflops_sustained', peaks.get('bf16_tflops', 1400.0)))
    igemm = [v for k, v in fam.items() if not k.endswith('wgrad')]
    fl, ms_k, n_k = sum(v[0] for v in igemm), sum(v[1] for v in igemm), sum(v[2] for v in igemm)
    achieved = fl / (ms_k * 1e-3) / 1e12 if ms_k > 0 else 0.0
    conv_ms = sum(v[1] for v in fam.values())
    conv_fl = sum(v[0] for v in fam.values())
    roofline = {'bound': 'tensor', 'kernel': 'igemm_kernel (tcgen05 implicit-GEMM conv fwd/dgrad/convT family)',
                'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
                'peak_source': f'{peak_src} bf16_tflops_sustained (kernel timed inside a long step)',
                'traffic': traffic.get('dram_bytes_per_launch'), 'traffic_note': traffic or None,
                'launches_timed': n_k, 'avg_launch_ms': ms_k / max(1, n_k),
                'all_conv_kernels': {'achieved': conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms else 0.0,
                                     'share_of_step': conv_ms / ms_total, 'eager_pass_ms_per_step': ms_eager_total / args.steps,
                                     'per_family_tflops': {k: v[0] / (v[1] * 1e-3) / 1e12 for k, v in fam.items() if v[1] > 0}},
                'algorithmic_flops_per_step': conv_fl / args.steps}
    # ---- timed region 2: end to end through the public API, host buffers -----------------------------------
    barrier()
    e0.record()
    for _ in range(args.steps):
        x = host.to(dev, non_blocking=True)
        loss, _, _ = run(x)
        lv = loss.item()                       # device → host read of the step's result
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e = {'value': B * world / (ms_e2e / 1e3), 'unit': UNIT, 'h2d_bytes_per_step': host.numel() * 4 * world,
           'd2h_bytes_per_step': 4 * world, 'ms_per_step': ms_e2e, 'last_loss': lv}
    out = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
           'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
           'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
           'config': {'workload': f'STUNet-{args.model} AnatoMask step (teacher fwd + hard mask + student fwd/bwd + clip + '
                                  f'AdamW + EMA), 1x{S}^3 volumes, batch {B}/GPU, mask 0.6, epoch 500/1000 (len_loss 76)',
                      'global_batch': B * world, 'parallelism': f'dp{world}',
                      'cuda_graph': use_graph, 'graph_error': graph_err,
                      'l2': 'per-step working set (multi-GB activations) >> 126 MB L2; no flush needed',
                      'algorithmic_tflop_per_volume': 4 * F_FWD_GF / 1e3 if (args.model == 'B' and S == 128) else None},
           'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches, 'roofline': roofline}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            out['cpu_baseline'] = cpu_baseline_sample()
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--model', default='B')
    ap.add_argument('--size', type=int, default=128)
    ap.add_argument('--batch', type=int, default=2)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true')
    a = ap.parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
