"""Step driver for the pre-training hot path: what P/pretrain.py:388-411 and P/pretrain_AntoMask.py:383-441 do per
iteration, re-designed for one B200 per process.

  * all float parameters and float buffers live in ONE fp32 arena ([live params | dead params | buffers]) and all live
    gradients in another: the EMA teacher update, the global-norm clip and AdamW are each a single vectorised kernel
    over the arena instead of 131 per-tensor chains, and the DDP gradient exchange is one NCCL all-reduce
  * no host synchronisation inside a step (`mask_rng='device'`): masks, work-lists, loss, clip coefficient and the
    hard-mask top-k all stay on the device
  * data parallel: batch sharded across ranks, NCCL only for the gradient all-reduce (and SyncBN statistics when the
    decoder was built with sbn=True) — P/pretrain_DDP.py:196-234
"""
from __future__ import annotations

import copy
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from . import AnatoMask as anatomask_mod
from . import spark3D
from .decoder3D import LightDecoder
from .encoder3D import SparseEncoder
from .STUNet_head import STUNet

STUNET_SIZES = {'S': (16, 1), 'B': (32, 1), 'L': (64, 2), 'H': (96, 3)}     # base width, blocks per stage


def build_model(size: str = 'B', input_size=(128, 128, 128), anatomask: bool = True, mask_ratio: float = 0.6,
                sbn: bool = False, device='cuda', base: Optional[int] = None, depth: Optional[int] = None) -> nn.Module:
    """The construction of P/pretrain.py:180-212 (STUNet head → SparseEncoder → LightDecoder → SparK)."""
    b, d = STUNET_SIZES[size] if base is None else (base, depth or 1)
    head = STUNet(1, 1, depth=[d] * 6, dims=[b * x for x in (1, 2, 4, 8, 16, 16)],
                  pool_op_kernel_sizes=[[2, 2, 2]] * 4 + [[1, 1, 1]], conv_kernel_sizes=[[3, 3, 3]] * 6)
    enc = SparseEncoder(head, input_size=input_size, sbn=sbn)
    dec = LightDecoder(enc.downsample_ratio, sbn=sbn, width=16 * b, out_channel=1)
    cls = anatomask_mod.SparK if anatomask else spark3D.SparK
    return cls(sparse_encoder=enc, dense_decoder=dec, mask_ratio=mask_ratio, densify_norm='in', sbn=sbn).to(device)


def dead_parameter_names(model: spark3D.SparK) -> List[str]:
    """Parameters the step never touches: densify levels the 4-block decoder does not consume."""
    n_live = len(model.dense_decoder.dec)
    pre = tuple(f'{grp}.{i}' for i in range(n_live, model.hierarchy) for grp in ('densify_norms', 'densify_projs', 'mask_tokens'))
    return [n for n, _ in model.named_parameters() if n.startswith(pre)]


def taps_major_weight_names(module: nn.Module) -> set:
    """Names of the conv weights the arena stores taps-major, fp32 [tap][Cout][Cin]: every ungrouped Conv3d / ConvTranspose3d
    whose channel counts the implicit-GEMM kernels take (multiples of 8) — exactly the weights that reach the kernels through
    ops.ConvFn.  The stem (Cin = 1), the 1-channel projection and depthwise convs are read through raw pointers by their own
    kernels and keep the stock layout."""
    out = set()
    for mn, mod in module.named_modules():
        if isinstance(mod, (nn.Conv3d, nn.ConvTranspose3d)) and getattr(mod, 'groups', 1) == 1 and mod.weight is not None \
                and mod.weight.dim() == 5 and mod.in_channels % 8 == 0 and mod.out_channels % 8 == 0:
            out.add((mn + '.' if mn else '') + 'weight')
    return out


def _taps_major_view(flat_slice: torch.Tensor, shape, transposed: bool) -> torch.Tensor:
    """A tensor of the parameter's shape — (Cout, Cin, k, k, k), ConvTranspose3d (Cin, Cout, 4, 4, 4) — over storage laid out
    [tap][Cout][Cin]."""
    c0, c1 = shape[0], shape[1]
    T = shape[2] * shape[3] * shape[4]
    if transposed:
        return flat_slice.view(T, c1, c0).permute(2, 1, 0).unflatten(-1, tuple(shape[2:]))
    return flat_slice.view(T, c0, c1).permute(1, 2, 0).unflatten(-1, tuple(shape[2:]))


class ParamArena:
    """Re-homes a module's float parameters/buffers into one contiguous fp32 buffer (views keep their names/shapes).

    Every tensor starts on a 16-byte boundary.  With `taps_major` (the engine's default) the conv weights the implicit-GEMM
    kernels consume are STORED as [tap][Cout][Cin] — the order of their bf16 operand copies and of the weight-gradient
    kernels' output — and the module's `weight` is a strided view with the reference's shape: packing a step's operands is
    then a dtype conversion (forward form) or a per-tap transpose (input-gradient form) instead of a gather with the taps
    innermost, and the weight gradients are accumulated by the kernels straight into the gradient arena (no staging
    buffer, no fill, no re-layout pass: 0.8 ms of a 20 ms STUNet-B step).  AdamW, EMA, the global-norm clip and the gradient
    all-reduce are element-wise over the flat buffers, so they never see the difference; state_dict() / load_state_dict()
    go through the strided views."""

    def __init__(self, module: nn.Module, dead: List[str], with_grads: bool, taps_major: bool = True):
        dev = next(module.parameters()).device
        named_p = list(module.named_parameters())
        live = [(n, p) for n, p in named_p if n not in dead]
        deadp = [(n, p) for n, p in named_p if n in dead]
        fbuf = [(n, b) for n, b in module.named_buffers() if b.is_floating_point()]
        self.int_buffers = [(n, b) for n, b in module.named_buffers() if not b.is_floating_point()]
        pad = lambda k: (k + 3) // 4 * 4
        self.taps_major = taps_major_weight_names(module) if taps_major else set()
        transposed = {(mn + '.' if mn else '') + 'weight' for mn, mod in module.named_modules()
                      if isinstance(mod, nn.ConvTranspose3d)}
        self.n_live = sum(pad(p.numel()) for _, p in live)
        n_dead, n_buf = sum(pad(p.numel()) for _, p in deadp), sum(pad(b.numel()) for _, b in fbuf)
        self.flat = torch.zeros(self.n_live + n_dead + n_buf, dtype=torch.float32, device=dev)
        self.offsets: Dict[str, Tuple[int, int]] = {}
        self._view_of = {}                       # name -> function(flat buffer) -> view with the parameter's shape
        starts = (0, self.n_live, self.n_live + n_dead)
        for group, off in zip((live, deadp, fbuf), starts):
            for n, t in group:
                k = t.numel()
                if n in self.taps_major:
                    make = (lambda f, o=off, k=k, sh=tuple(t.shape), tr=(n in transposed): _taps_major_view(f[o:o + k], sh, tr))
                else:
                    make = (lambda f, o=off, k=k, sh=tuple(t.shape): f[o:o + k].view(sh))
                view = make(self.flat)
                view.copy_(t.data)
                t.data = view
                self.offsets[n] = (off, k)
                self._view_of[n] = make
                off += pad(k)
        self.grad = None
        if with_grads:
            self.grad = torch.zeros(self.n_live, dtype=torch.float32, device=dev)
            for n, p in live:
                p.grad = self._view_of[n](self.grad)
        if self.int_buffers:
            self.iflat = torch.zeros(len(self.int_buffers), dtype=torch.int64, device=dev)
            for i, (_, b) in enumerate(self.int_buffers):
                self.iflat[i] = b
                b.data = self.iflat[i]
        else:
            self.iflat = None

    def views(self, flat_like: torch.Tensor, names=None) -> Dict[str, torch.Tensor]:
        """name -> view of another flat buffer laid out like this arena (Adam moments, a gradient snapshot) with the
        parameter's shape and element order — what a checkpoint stores, independent of the arena's storage order."""
        names = [n for n, (o, _) in self.offsets.items() if o < flat_like.numel()] if names is None else names
        return {n: self._view_of[n](flat_like) for n in names}

    def zero_grad(self):
        self.grad.zero_()


def lr_at_epoch(epoch: int, base_lr: float, warmup: int = 20, max_epochs: int = 1000, warmup_start_lr: float = 1e-6,
                eta_min: float = 0.0) -> float:
    """Closed form of N/training/lr_scheduler/LinearWarmupCosine.py as the scripts use it (stepped per epoch)."""
    if epoch < warmup:
        return warmup_start_lr + epoch * (base_lr - warmup_start_lr) / max(1, warmup - 1)
    return eta_min + 0.5 * (base_lr - eta_min) * (1 + math.cos(math.pi * (epoch - warmup) / (max_epochs - warmup)))


def ema_decay_at_epoch(epoch: int, epochs: int) -> float:
    """P/pretrain_AntoMask.py:383-386."""
    q = epochs // 4
    return 0.999 + epoch / q * (0.9999 - 0.999) if epoch < q else 0.9999


class PretrainEngine:
    """One process = one GPU.  `step()` is the AnatoMask iteration (teacher fwd → hard mask → student fwd/bwd → clip +
    AdamW → EMA), `spark_step()` the plain SparK iteration."""

    def __init__(self, model: spark3D.SparK, lr: float = 1e-4, weight_decay: float = 1e-5, clip: float = 12.0,
                 betas=(0.9, 0.999), eps: float = 1e-8, epochs: int = 1000, anatomask: bool = True,
                 mask_rng: str = 'device', process_group=None, taps_major: bool = True):
        import os
        if os.environ.get('AMB_NO_TAPS_MAJOR') == '1':       # A/B switch: stock parameter layout in the arena
            taps_major = False
        self.model = model
        self.lr, self.wd, self.clip, self.betas, self.eps, self.epochs = lr, weight_decay, clip, betas, eps, epochs
        self.dead = dead_parameter_names(model)
        self.teacher = None
        if anatomask:
            self.teacher = copy.deepcopy(model)          # timm ModelEma: deepcopy → eval → no grad
            self.teacher.eval()
            for p in self.teacher.parameters():
                p.requires_grad_(False)
            self.teacher.mask_rng = mask_rng
            self.tarena = ParamArena(self.teacher, self.dead, with_grads=False, taps_major=taps_major)
        self.arena = ParamArena(model, self.dead, with_grads=True, taps_major=taps_major)
        self.m = torch.zeros_like(self.arena.grad)
        self.v = torch.zeros_like(self.arena.grad)
        self.t = 0
        self.group = process_group
        self.world = 1
        if process_group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(process_group)
        self.mask_rng = mask_rng
        self._rng_calls = 0
        self.buckets = None
        if self.world > 1:
            from .parallel import GradBuckets
            dev0 = self.arena.flat.device
            self.buckets = GradBuckets(self.arena.grad, self.arena.offsets, self.arena.n_live, process_group,
                                       side_stream_fn=(lambda: ops._side_stream(dev0)) if dev0.type == 'cuda' else None)
        # CUDA-graph mode (no host sync in the step): per-step scalars live in `hyper` on the device, the RNG stream is
        # driven by a device step counter, and the whole step (teacher fwd, hard mask, student fwd/bwd, all-reduce,
        # clip + AdamW, EMA — ~430 launches) is replayed as ONE graph launch
        dev = self.arena.flat.device
        self.hyper = torch.zeros(16, dtype=torch.float32, device=dev)
        # the per-step scalars are staged through a RING of pinned host buffers, each guarded by an event recorded after its
        # H2D copy: the host may run many steps ahead of the device (nothing in graph_step synchronises), and rewriting a
        # single staging buffer would hand an earlier step the lr / bias corrections / EMA decay of a later one
        self._hyper_ring = [[torch.zeros(16, dtype=torch.float32).pin_memory() if dev.type == 'cuda' else torch.zeros(16), None]
                            for _ in range(8)]
        self._hyper_slot = 0
        self.step_counter = torch.zeros(1, dtype=torch.int64, device=dev)
        self._graphs = {}
        self._static_inp = None
        self._stage = None                   # (copy stream, device buffer) of stage_input()
        self._stage_ready = self._stage_consumed = None
        self._stage_pending = False
        self._static_out = None

    # ---- pieces ---------------------------------------------------------------------------------------------
    def random_mask(self, B: int, device) -> torch.Tensor:
        m = self.model
        if self.mask_rng == 'device':                    # len_keep smallest of B×L counter-based uniforms, on device
            self._rng_calls += 1
            L = m.fmap_h * m.fmap_w * m.fmap_d
            zeros = torch.zeros(B, L, dtype=torch.float32, device=device)
            _, mk = ops.hard_mask(zeros, 0, m.len_keep, seed=0xA11CE, offset=self._rng_calls * B * L)
            return mk.bool().view(B, 1, m.fmap_h, m.fmap_w, m.fmap_d)
        return m.mask(B, device)

    def _optimise(self, lr: float):
        a = self.arena
        self._allreduce_grads()
        self.t += 1
        ops.adamw_step_(a.flat[:a.n_live], a.grad, self.m, self.v, lr, self.betas, self.eps, self.wd, self.t,
                        self.clip, 1.0 / self.world)

    def ema_update(self, decay: float):
        ops.ema_update_(self.tarena.flat, self.arena.flat, decay)
        if self.tarena.iflat is not None:                # int64 num_batches_tracked: lerp then truncate (timm copy_)
            ti, si = self.tarena.iflat, self.arena.iflat
            ti.copy_(ti * decay + (1. - decay) * si)

    # ---- CUDA-graph path --------------------------------------------------------------------------------------
    def _set_hyper(self, epoch: int):
        b1, b2 = self.betas
        t = self.t + 1
        d = ema_decay_at_epoch(epoch, self.epochs)
        slot = self._hyper_ring[self._hyper_slot]
        self._hyper_slot = (self._hyper_slot + 1) % len(self._hyper_ring)
        h, ev = slot
        if ev is not None:
            ev.synchronize()                       # blocks only when the host is a full ring ahead of the device
        h[0], h[1] = d, 1.0 - d
        h[2], h[3], h[4], h[5], h[6] = lr_at_epoch(epoch, self.lr, max_epochs=self.epochs), b1, b2, self.eps, self.wd
        h[7], h[8], h[9], h[10] = 1.0 - b1 ** t, math.sqrt(1.0 - b2 ** t), self.clip, 1.0 / self.world
        self.hyper.copy_(h, non_blocking=True)
        if self.hyper.is_cuda:
            if ev is None:
                ev = slot[1] = torch.cuda.Event()
            ev.record()

    def _device_front(self, inp: torch.Tensor, len_loss_epoch: int):
        """Teacher forward → hard mask → student forward/backward, every per-step scalar read from device memory."""
        B = inp.shape[0]
        m = self.model
        Lp = m.fmap_h * m.fmap_w * m.fmap_d
        # bf16 operand copies of every conv weight (student forward + dgrad forms, teacher forward form): one launch per
        # step once the first step has recorded which packs the modules ask for
        plan = getattr(self, '_pack_plan', None)
        tag = (self.arena.flat.data_ptr(), self.tarena.flat.data_ptr())
        if plan is not None and getattr(self, '_pack_plan_tag', None) != tag:
            plan = self._pack_plan = None              # the arenas were rebuilt: recorded weight addresses are stale
        self._pack_plan_tag = tag
        if plan is None:
            ops.PACK_RECORD = []
        else:
            # teacher operands on this stream (the teacher forward is next); the student's — two thirds of the bytes, first
            # read a whole teacher forward later — on the side stream, next to the teacher's small latency-bound layers
            plan.run(0)
            side = ops._side_stream(inp.device) if 1 in plan.groups else None
            if side is not None:
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    plan.run(1)
            elif 1 in plan.groups:
                plan.run(1)
            ops.PACK_CACHE = plan.cache
        if getattr(self, '_zero_pool', None) is None:
            self._zero_pool = torch.empty(4 << 20, dtype=torch.uint8, device=inp.device)
        self._zero_pool.zero_()                    # every small accumulator of the step (ops._zeros_small): one memset
        ops.ZERO_POOL = {'buf': self._zero_pool, 'off': 0}
        try:
            return self._device_front_body(inp, len_loss_epoch, B, m, Lp)
        finally:
            ops.ZERO_POOL = None
            if plan is None and ops.PACK_RECORD:
                lo, hi = self.arena.flat.data_ptr(), self.arena.flat.data_ptr() + self.arena.flat.numel() * 4
                self._pack_plan = ops.PackPlan(ops.PACK_RECORD, group_of=lambda w: 1 if lo <= w.data_ptr() < hi else 0)
            ops.PACK_RECORD = None
            ops.PACK_CACHE = None

    def _device_front_body(self, inp, len_loss_epoch, B, m, Lp):
        self.step_counter.add_(1)
        zeros = torch.zeros(B, Lp, dtype=torch.float32, device=inp.device)
        _, mk = ops.hard_mask(zeros, 0, m.len_keep, seed=0xA11CE, offset=0, offset_dev=self.step_counter)
        mask1 = mk.bool().view(B, 1, m.fmap_h, m.fmap_w, m.fmap_d)
        with torch.no_grad(), ops.lean_zero():
            rec1 = self.teacher.reconstruct(inp, mask1)
            recon = self.teacher.teacher_loss(inp, rec1, mask1)
        mask, _ = self.teacher.generate_mask(recon, guide=True, epoch=len_loss_epoch, total_epoch=self.epochs - 1)
        ops.join_side_stream(inp.device)           # the student's packed weights (side stream) are complete
        loss = self._student_fwd_bwd(inp, mask, defer_wgrad=True)
        ops.join_side_stream(inp.device)
        return loss, mask, recon

    def _device_tail(self):
        """Global-norm clip + AdamW + EMA from the (already all-reduced) gradient arena; scalars from `hyper`."""
        from . import _lib as L
        a, ta = self.arena, self.tarena
        gn = torch.zeros(1, dtype=torch.float64, device=a.flat.device)
        L.call('amb_sumsq', ops._p(a.grad), a.grad.numel(), ops._p(gn), ops._stream())
        L.call('amb_step_dev', ops._p(ta.flat), ops._p(a.flat), ta.flat.numel(), ops._p(a.flat), ops._p(a.grad),
               ops._p(self.m), ops._p(self.v), a.n_live, ops._p(self.hyper), ops._p(gn), 1, 1, ops._stream())
        if ta.iflat is not None:
            ta.iflat.copy_(ta.iflat * self.hyper[0] + self.hyper[1] * a.iflat)

    def _allreduce_grads(self):
        """Joins the bucketed gradient exchange: groups whose backward mark fired are already in flight (overlapped with
        the rest of the backward pass), the remaining ones — the encoder group — start here."""
        if self.buckets is not None:
            ops.join_side_stream(self.arena.flat.device)
            self.buckets.finish()

    def _student_fwd_bwd(self, inp: torch.Tensor, mask: torch.Tensor, defer_wgrad: bool = False) -> torch.Tensor:
        """Student forward + backward.  With more than one rank the gradient all-reduce of each parameter group is started
        from the backward pass as soon as that group is complete (ops.backward_mark: decoder → densify → encoder; what the
        DDP reducer's buckets do in P/pretrain_DDP.py:231-232) and joined by _allreduce_grads()."""
        m = self.model
        if self.buckets is not None:
            from .parallel import MARK_OF_GROUP
            group_of_mark = {mk: g for g, mk in MARK_OF_GROUP.items()}
            self.buckets.begin_step()
            ops.MARK_CALLBACK = lambda tag: self.buckets.start(group_of_mark[tag])
        try:
            with ops.lean_zero():                  # masked voxels nothing reads are not zero-filled (ops.LEAN_ZERO)
                rec = m.reconstruct(inp, mask)
                loss, _ = ops.PatchLossFn.apply(inp, rec, mask[:, 0].to(torch.uint8).contiguous(), True)
                self.arena.zero_grad()
                ops.DEFER_WGRAD = defer_wgrad      # wgrad chain → side stream, written straight into the arena views
                loss.backward()
        finally:
            ops.DEFER_WGRAD = False
            ops.MARK_CALLBACK = None
        return loss.detach()

    def _device_step(self, inp: torch.Tensor, len_loss_epoch: int):
        out = self._device_front(inp, len_loss_epoch)
        self._allreduce_grads()
        self._device_tail()
        return out

    def device_step(self, inp: torch.Tensor, epoch: int = 0):
        """The launch sequence graph_step() captures, issued eagerly (profilers, debugging): device RNG masks, device-side
        scalars, batched weight packing — no host synchronisation."""
        self.model.train()
        self.teacher.mask_rng = 'device'
        self.teacher.rng_counter = self.step_counter
        self._set_hyper(epoch)
        staged = self._staged_begin(inp)
        out = self._device_step(inp, epoch)
        self._staged_end(staged)                     # eager steps read the batch until their last kernel
        self.t += 1
        return out

    def _state_snapshot(self):
        keep = [self.arena.flat, self.m, self.v, self.tarena.flat, self.step_counter]
        if self.arena.iflat is not None:
            keep += [self.arena.iflat, self.tarena.iflat]
        return keep, [t.clone() for t in keep]

    def graph_step(self, inp: torch.Tensor, epoch: int = 0):
        """Same semantics as step() in device-RNG mode, replayed from a CUDA graph.  Graphs are keyed by the number of
        hard patches (a launch parameter of the top-k kernel that changes every few epochs)."""
        m = self.model
        nm = m.fmap_h * m.fmap_w * m.fmap_d - m.len_keep
        len_loss = int(nm * (float((epoch + 1) / (self.epochs - 1)) * 0.5))
        self.model.train()
        if self._static_inp is None or self._static_inp.shape != inp.shape:
            self._static_inp = torch.empty_like(inp)
            self._graphs.clear()
        self._consume_input(inp)
        self._set_hyper(epoch)
        # world > 1: the NCCL collectives (bucketed gradient all-reduce started from the backward pass, SyncBN statistics)
        # are captured INTO the step graph, so a step stays one graph launch and the exchange overlaps the backward pass.
        # AMB_NCCL_OUTSIDE_GRAPH=1 keeps NCCL out of the capture: front graph → eager all-reduce → tail graph (no overlap;
        # not available with SyncBN, whose all-reduces sit inside the forward / backward).
        import os
        split = self.world > 1 and os.environ.get('AMB_NCCL_OUTSIDE_GRAPH') == '1'
        if split and getattr(m, 'sbn', False):
            return self.step(inp, epoch)
        if len_loss not in self._graphs:
            self._graphs.clear()                         # one resident graph (its private pool holds a full step)
            self.teacher.mask_rng = 'device'
            self.teacher.rng_counter = self.step_counter
            tensors, saved = self._state_snapshot()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                # warm-up on a side stream (allocator, function attributes, NCCL)
                for _ in range(2):
                    self._device_step(self._static_inp, epoch)
            torch.cuda.current_stream().wait_stream(side)
            if self.world > 1:
                torch.cuda.synchronize()                 # nothing of the warm-up may still be polled by NCCL's watchdog
            mode = {'capture_error_mode': 'thread_local'} if self.world > 1 else {}
            # The main chain (forward passes, input gradients, norms — every kernel waits for the one before it) is captured
            # on a HIGH-priority stream, the deferred weight-gradient branch stays on the default-priority side stream: kernel
            # nodes keep their stream's priority, so whenever SMs free up the block scheduler hands them to the critical path
            # first and the weight gradients fill in behind (both are persistent one-CTA-per-SM kernels: without priorities
            # whichever was launched first holds the machine, and a weight gradient ahead of its layer's input gradient
            # delays everything downstream).  AMB_NO_PRIORITY=1: capture on a default-priority stream (A/B).
            if os.environ.get('AMB_NO_PRIORITY') != '1':
                mode['stream'] = torch.cuda.Stream(priority=-1)
            if split:
                g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                keep, self.buckets = self.buckets, None  # no collective may be issued while capturing
                try:
                    with torch.cuda.graph(g1, **mode):
                        out = self._device_front(self._static_inp, epoch)
                    with torch.cuda.graph(g2, **mode):
                        self._device_tail()
                finally:
                    self.buckets = keep
                graphs = (g1, g2)
            else:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, **mode):
                    out = self._device_step(self._static_inp, epoch)
                graphs = (g,)
            for t, s0 in zip(tensors, saved):            # warm-up / capture must not count as training steps
                t.copy_(s0)
            self._graphs[len_loss] = (graphs, out)
        graphs, out = self._graphs[len_loss]
        if len(graphs) == 2:
            graphs[0].replay()
            self.buckets.begin_step()
            self._allreduce_grads()
            graphs[1].replay()
        else:
            graphs[0].replay()
        self.t += 1
        return out

    # ---- input staging ------------------------------------------------------------------------------------
    def stage_input(self, host: torch.Tensor) -> torch.Tensor:
        """Starts the host → device copy of the NEXT batch on a copy stream while the current step runs (what the scripts'
        pinned DataLoader + `.to(device, non_blocking=True)` do, P/pretrain_AntoMask.py:312-345, 419).  Returns the device
        buffer; hand it to graph_step(), which waits for the copy and frees the buffer for the following stage_input()
        as soon as it has taken the batch over.  One buffer: stage, step, stage, step, ..."""
        dev = self.arena.flat.device
        if self._stage is None or self._stage[1].shape != host.shape:
            self._stage = (torch.cuda.Stream(dev), torch.empty(host.shape, dtype=torch.float32, device=dev))
            self._stage_consumed = None
        cs, buf = self._stage
        if self._stage_pending:
            raise RuntimeError('stage_input: the batch staged before has not been handed to a step yet (one staging buffer: '
                               'stage, step, stage, step, ...)')
        if self._stage_consumed is not None:
            cs.wait_event(self._stage_consumed)          # the previous batch has been copied out of the buffer
        else:
            cs.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(cs):
            buf.copy_(host, non_blocking=True)
            self._stage_ready = torch.cuda.Event()
            self._stage_ready.record(cs)
        self._stage_pending = True
        return buf

    def _staged_begin(self, inp: torch.Tensor) -> bool:
        """A step is about to read `inp`: if it is the staging buffer, wait for its copy."""
        staged = self._stage is not None and inp is self._stage[1]
        if staged:
            torch.cuda.current_stream().wait_event(self._stage_ready)
        return staged

    def _staged_end(self, staged: bool):
        """The step's last read of the staging buffer has been enqueued: the next stage_input() may overwrite it after this."""
        if staged:
            self._stage_consumed = torch.cuda.Event()
            self._stage_consumed.record()
            self._stage_pending = False

    def _consume_input(self, inp: torch.Tensor):
        staged = self._staged_begin(inp)
        self._static_inp.copy_(inp, non_blocking=True)
        self._staged_end(staged)

    # ---- steps ----------------------------------------------------------------------------------------------
    def spark_step(self, inp: torch.Tensor, active: Optional[torch.Tensor] = None, epoch: int = 0) -> torch.Tensor:
        self.model.train()
        if active is None:
            active = self.random_mask(inp.shape[0], inp.device)
        loss = self._student_fwd_bwd(inp, active)
        self._optimise(lr_at_epoch(epoch, self.lr, max_epochs=self.epochs))
        return loss

    def step(self, inp: torch.Tensor, epoch: int = 0, mask1: Optional[torch.Tensor] = None):
        """P/pretrain_AntoMask.py:419-440.  Returns (loss, hard mask, teacher per-patch loss) — all on device."""
        self.model.train()
        # host-driven RNG offsets here: the device step counter only advances inside device_step / graph_step, and a
        # counter left attached would freeze the random fill of generate_mask at one stream position
        self.teacher.rng_counter = None
        B = inp.shape[0]
        staged = self._staged_begin(inp)
        if mask1 is None:
            mask1 = self.random_mask(B, inp.device)
        with torch.no_grad(), ops.lean_zero():
            rec1 = self.teacher.reconstruct(inp, mask1)
            recon = self.teacher.teacher_loss(inp, rec1, mask1)
        mask, _ = self.teacher.generate_mask(recon, guide=True, epoch=epoch, total_epoch=self.epochs - 1)
        loss = self._student_fwd_bwd(inp, mask)
        self._optimise(lr_at_epoch(epoch, self.lr, max_epochs=self.epochs))
        self.ema_update(ema_decay_at_epoch(epoch, self.epochs))
        self._staged_end(staged)
        return loss, mask, recon
