"""Headline benchmark: 128³ volumes/s of one STUNet-B AnatoMask pre-training step (BASELINE.json), per the driver contract.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one process per GPU; torchrun for N > 1)
  python bench.py --impl reference [...]                         the reference's CPU implementation (oracle port) on the
                                                                 host cores, same config/metric, bounded sample per step

A step = teacher forward (eval) → per-patch teacher loss → hard-mask top-k → student forward/backward → global-norm clip
+ AdamW → EMA teacher update, on a batch of 2 synthetic N(0,1) volumes of 1×128³ per GPU (weak scaling), epoch 500 of
1000 (len_loss = 76 hard patches).  `value`: inputs resident in HBM.  `e2e`: the same step through the public API with
the batch coming from pinned host memory every step (a different host batch each step) and the loss read back to the host.

  --model L --sbn      BASELINE config 4: STUNet-L, decoder SyncBN (P/pretrain_DDP.py:224-225), batch-sharded over N GPUs
  At N=1 the line also carries (unless --no-extras):
  gpu_reference        the oracle port on the SAME GPU through torch + cuDNN under bf16 autocast ("the real
                       kernel-to-beat"), separate from `cpu_baseline`
  len_loss_sweep       the step at epochs 0 / 998 (len_loss 0 / 153; the headline epoch 500 has len_loss 76)
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


def gpu_reference_sample(dev, batch: int, steps: int = 3):
    """The oracle port (plain PyTorch restatement of the reference modules) running the same AnatoMask step on THIS GPU
    through torch's own kernels (cuDNN convolutions) under torch.autocast(bfloat16) — P/pretrain.py:394-401 is the
    reference's autocast site.  Context for the hand-written kernels; never part of the product path."""
    from oracle import reference_port as rp
    import numpy as np
    cfg = rp.CONFIGS['B128']
    tr = rp.RefTrainer(cfg, {k: v.to(dev) for k, v in rp.make_state(cfg, 0).items()}, lr=1e-4, epochs=1000, anatomask=True)
    np.random.seed(0)
    inp = rp.make_input(cfg, batch, 0).to(dev)
    times = []
    for i in range(steps + 1):
        mask1 = rp.random_mask(cfg, batch, torch.Generator().manual_seed(i)).to(dev)
        torch.cuda.synchronize()
        t0 = time.time()
        with torch.autocast('cuda', dtype=torch.bfloat16):
            with torch.no_grad():                                   # P/pretrain_AntoMask.py:419-441 over the port
                rec1 = rp.forward(tr.ema, cfg, inp, mask1, training=False)
                recon = rp.teacher_patch_loss(cfg, inp, rec1.float(), mask1)
            mask = rp.generate_mask(cfg, recon.float().cpu(), 500, 999, True, np.random).to(dev)
            tr.spark_step(inp, mask)
        rp.ema_update(tr.ema, tr.state, rp.ema_decay(500, 1000))
        torch.cuda.synchronize()
        if i > 0:
            times.append(time.time() - t0)
    dt = sorted(times)[len(times) // 2]
    return {'value': batch / dt, 'unit': UNIT, 'kind': 'port on cuda (torch + cuDNN, bf16 autocast)', 'ms_per_step': dt * 1e3,
            'sample': f'median of {steps} AnatoMask steps of {batch} volumes (1x128^3) with the oracle port on the same GPU'}


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
    model = build_model(args.model, (S, S, S), anatomask=True, sbn=args.sbn)
    if world > 1:                                    # identical initial weights on every rank (DDP broadcast)
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, 0)
    _mark('model built + broadcast')
    eng = PretrainEngine(model, lr=1e-4, epochs=1000, anatomask=True, mask_rng='device', process_group=group)
    inp = torch.randn(B, 1, S, S, S, device=dev)
    hosts = [torch.randn(B, 1, S, S, S).pin_memory() for _ in range(3)]      # e2e: a different host batch every step
    host = hosts[0]
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
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e0.record()
    for i in range(args.steps):
        run(inp)
        marks[i].record()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    per_step = sorted(a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks))
    ms_median = max_over_ranks(per_step[len(per_step) // 2])
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
    fam, kern = {}, {}
    for kind, flops, a, b, kname in prof:
        ms = a.elapsed_time(b)
        d = fam.setdefault(kind, [0.0, 0.0, 0])
        d[0] += flops; d[1] += ms; d[2] += 1
        d = kern.setdefault(kname, [0.0, 0.0, 0])
        d[0] += flops; d[1] += ms; d[2] += 1
    if os.environ.get('AMB_BENCH_DUMP') and rank == 0:           # per-launch table of the last profiled step
        n_per = len(prof) // args.steps
        rows = [(k, f, a.elapsed_time(b), kn) for k, f, a, b, kn in prof[-n_per:]]
        for i, (k, f, ms, kn) in enumerate(rows):
            print(f'#LAUNCH {i:3d} {k:12s} {f / 1e9:9.1f} GF {ms:7.3f} ms {f / ms / 1e9:8.1f} TF/s  {kn}', file=sys.stderr)
    peaks, peak_src = _peaks()
    # DRAM bytes per launch of the dominant kernel's dominant shape: a CONSTANT read from the committed ncu --set full
    # capture (profiles/*_ncu_traffic.json), not measured in this run — ncu cannot run inside the timed bench
    traffic = {}
    for name in ('r2_ncu_traffic.json', 'r1_ncu_traffic.json'):
        try:
            traffic = json.load(open(os.path.join(ROOT, 'profiles', name)))
            break
        except (OSError, ValueError):
            continue
    peak = float(peaks.get('bf16_tflops_sustained', peaks.get('bf16_tflops', 1400.0)))
    burst = float(peaks.get('bf16_tflops', peak))
    tf = lambda v: v[0] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else 0.0
    # the dominant kernel = the single kernel with the largest share of the step's device time
    dom = max(kern, key=lambda k: kern[k][1]) if kern else ''
    dv = kern.get(dom, [0.0, 0.0, 0])
    igemm = [v for k, v in fam.items() if not k.endswith('wgrad')]
    fam_v = [sum(v[0] for v in igemm), sum(v[1] for v in igemm), sum(v[2] for v in igemm)]
    conv_ms = sum(v[1] for v in fam.values())
    conv_fl = sum(v[0] for v in fam.values())
    roofline = {'bound': 'tensor', 'kernel': dom, 'achieved': tf(dv), 'peak': peak, 'unit': 'TFLOP/s', 'frac': tf(dv) / peak,
                'frac_of_burst_peak': tf(dv) / burst,
                'peak_source': f'{peak_src} bf16_tflops_sustained (kernel timed inside a long step); burst = bf16_tflops',
                'traffic': traffic.get('dram_bytes_per_launch'),
                'traffic_note': dict(traffic, measured_in_this_run=False) if traffic else None,
                'launches_timed': dv[2], 'avg_launch_ms': dv[1] / max(1, dv[2]), 'share_of_step': dv[1] / ms_total,
                'per_kernel': {k: {'tflops': tf(v), 'frac': tf(v) / peak, 'launches': v[2], 'share_of_step': v[1] / ms_total}
                               for k, v in sorted(kern.items(), key=lambda kv: -kv[1][1])},
                'conv_fwd_dgrad_family': {'achieved': tf(fam_v), 'frac': tf(fam_v) / peak, 'frac_of_burst_peak': tf(fam_v) / burst,
                                          'launches_timed': fam_v[2]},
                'all_conv_kernels': {'achieved': conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms else 0.0,
                                     'share_of_step': conv_ms / ms_total, 'eager_pass_ms_per_step': ms_eager_total / args.steps,
                                     'per_family_tflops': {k: v[0] / (v[1] * 1e-3) / 1e12 for k, v in fam.items() if v[1] > 0}},
                'algorithmic_flops_per_step': conv_fl / args.steps}
    # ---- timed region 2: end to end through the public API, host buffers -----------------------------------
    barrier()
    e0.record()
    prefetch = use_graph and os.environ.get('AMB_BENCH_NO_PREFETCH') != '1'
    nxt = eng.stage_input(hosts[0]) if prefetch else None      # inside the timed region: one H2D copy per step (+ this one)
    for i in range(args.steps):
        if prefetch:
            loss, _, _ = run(nxt)                               # waits for the staged copy, takes the batch over
            nxt = eng.stage_input(hosts[(i + 1) % len(hosts)])  # next batch's H2D copy overlaps this step (pinned loader)
        else:
            x = hosts[i % len(hosts)].to(dev, non_blocking=True)
            loss, _, _ = run(x)
        lv = loss.item()                       # device → host read of the step's result, every step
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e = {'value': B * world / (ms_e2e / 1e3), 'unit': UNIT, 'h2d_bytes_per_step': host.numel() * 4 * world,
           'd2h_bytes_per_step': 4 * world, 'ms_per_step': ms_e2e, 'last_loss': lv,
           'input_staging': 'next batch copied from pinned memory on a copy stream during the current step '
                            '(PretrainEngine.stage_input)' if prefetch else 'copy, step, read in series'}
    # ---- easy→hard schedule: the same step at the first and last epochs (len_loss 0 and 153 hard patches) ----------
    sweep = None
    if use_graph and (args.sweep or (world == 1 and not args.no_extras)):
        sweep = {'76': value}
        for ep, ll in ((0, '0'), (998, '153')):
            for _ in range(3):
                eng.graph_step(inp, ep)
            barrier()
            e0.record()
            for _ in range(5):
                eng.graph_step(inp, ep)
            e1.record()
            barrier()
            sweep[ll] = B * world / (max_over_ranks(e0.elapsed_time(e1)) / 5 / 1e3)
    label = 'METRIC' if (args.model == 'B' and not args.sbn) else None
    metric = METRIC if label else f'128^3 volumes/sec, STUNet-{args.model} AnatoMask pretraining step' + (', decoder SyncBN' if args.sbn else '')
    out = {'metric': metric, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
           'warmup': args.warmup, 'ms_per_step': ms_step, 'ms_per_step_median': ms_median, 'higher_is_better': True,
           'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
           'config': {'workload': f'STUNet-{args.model} AnatoMask step (teacher fwd + hard mask + student fwd/bwd + clip + '
                                  f'AdamW + EMA), 1x{S}^3 volumes, batch {B}/GPU, mask 0.6, epoch 500/1000 (len_loss 76)'
                                  + (', decoder SyncBN (sbn=True)' if args.sbn else ''),
                      'gradient_exchange': None if world == 1 else 'bucketed NCCL all-reduce (decoder / densify / encoder groups) '
                                                                   'started from the backward pass, captured in the step graph',
                      'global_batch': B * world, 'parallelism': f'dp{world}',
                      'cuda_graph': use_graph, 'graph_error': graph_err,
                      'l2': 'per-step working set (multi-GB activations) >> 126 MB L2; no flush needed',
                      'algorithmic_tflop_per_volume': 4 * F_FWD_GF / 1e3 if (args.model == 'B' and S == 128) else None},
           'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches, 'roofline': roofline}
    if sweep is not None:
        out['len_loss_sweep'] = {'unit': UNIT, 'by_len_loss': sweep,
                                 'note': 'epochs 0 / 500 / 998 of 1000 → 0 / 76 / 153 hard patches of 307 masked (SURVEY §8d C3)'}
    if rank == 0:
        if world == 1 and args.model == 'B' and S == 128 and (args.gpu_reference or not args.no_extras):
            del eng, model
            torch.cuda.empty_cache()
            try:
                out['gpu_reference'] = gpu_reference_sample(dev, B)
            except Exception as e:                                   # noqa: BLE001 — context only, never fails the bench line
                out['gpu_reference'] = {'unavailable': f'{type(e).__name__}: {e}'[:200]}
        if world == 1 and not args.no_cpu_baseline:
            out['cpu_baseline'] = cpu_baseline_sample()
        print(json.dumps(out))
    if world > 1:
        _hard_exit(dist, dev)


def _hard_exit(dist, dev):
    """Multi-rank teardown.  The step graph holds captured NCCL kernels; `destroy_process_group()` (and the interpreter's
    own teardown of the communicator) was observed to block forever after the JSON line had been printed (round 2, 2 GPUs,
    both bench.py and tests/dist_checks.py sat in it until the outer timeout).  All ranks meet at a barrier with their
    streams drained — nothing is in flight any more — and then leave without running NCCL's destructors."""
    torch.cuda.synchronize()
    t = torch.zeros(1, device=dev)
    dist.all_reduce(t)
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


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
    ap.add_argument('--sbn', action='store_true', help='decoder nn.SyncBatchNorm (BASELINE config 4)')
    ap.add_argument('--sweep', action='store_true', help='also time epochs 0 and 998 (len_loss 0 / 153); default at N=1')
    ap.add_argument('--no-extras', action='store_true', help='N=1: skip the len_loss sweep and the same-GPU torch+cuDNN reference')
    ap.add_argument('--gpu-reference', action='store_true', help='also time the oracle port on this GPU (torch+cuDNN, bf16 autocast)')
    a = ap.parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
