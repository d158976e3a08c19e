"""Autograd operators over the C ABI.  Internal activation format: bf16, shape (N, D, H, W, C), contiguous.

Each Function enqueues hand-written sm_100a kernels on torch's current stream through ctypes; PyTorch only owns the
memory (caching allocator), the streams and the autograd tape.  Nothing here synchronises the host.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import os

import torch

from . import _lib as L

bf16 = torch.bfloat16


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


# ConvFn.backward runs the weight-gradient branch (wgrad → unpack → bias column sums) on a side stream: dgrad and wgrad
# only share the read-only dy, so inside the captured step graph they become parallel branches — the deep, small layers
# launch too few CTAs to fill 148 SMs on their own.  AMB_NO_SIDE_STREAM=1 disables it.
_SIDE = {}
# Engine mode: weight / bias gradients are written straight into the (pre-allocated) .grad views of the parameter arena
# on the side stream and ConvFn.backward returns None for them, so the whole weight-gradient chain of the backward pass
# runs as one long parallel branch next to the dgrad / norm-backward chain; the engine joins it before the optimiser.
DEFER_WGRAD = False


def join_side_stream(dev=None):
    dev = torch.device('cuda', torch.cuda.current_device()) if dev is None else dev
    s = _SIDE.get(dev)
    if s is not None:
        torch.cuda.current_stream().wait_stream(s)



NO_SIDE = False      # bench.py's per-launch timing pass: one stream, so a launch's CUDA-event pair times that launch alone


def _side_stream(dev):
    import os
    if NO_SIDE or os.environ.get('AMB_NO_SIDE_STREAM') == '1':
        return None
    s = _SIDE.get(dev)
    if s is None:
        s = _SIDE[dev] = torch.cuda.Stream(device=dev)
    return s


# bench.py sets PROFILE to a list: every conv-family launch is then bracketed by CUDA events on the launching stream
# and recorded as (kind, algorithmic FLOPs, start, end) — the live roofline measurement (no effect when None).
PROFILE = None


class _Timed:
    def __init__(self, kind: str, flops: float, tag: str = ''):
        self.kind, self.flops = kind + tag, flops

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e1.record()
            kernel = (L.load().amb_last_conv_kernel() or b'').decode()      # which kernel the dispatcher picked
            PROFILE.append((self.kind, self.flops, self.e0, self.e1, kernel))
        return False


# Backward-progress marks: SparK.reconstruct threads its tensors through identity nodes at the two points where a whole
# parameter group has finished its backward pass (decoder → densify → encoder, the order autograd runs them in).  The
# engine sets MARK_CALLBACK for the duration of a step and starts that group's gradient all-reduce from the callback, so
# the exchange overlaps the rest of the backward pass (the DDP reducer's job, P/pretrain_DDP.py:231-232).  None = no-op.
MARK_CALLBACK = None


class _MarkFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tag, *xs):
        ctx.tag = tag
        ctx.set_materialize_grads(False)
        return tuple(x.view_as(x) for x in xs)

    @staticmethod
    def backward(ctx, *gs):
        cb = MARK_CALLBACK
        if cb is not None:
            cb(ctx.tag)
        return (None,) + gs


def backward_mark(tag: str, *xs):
    """Identity on xs; when the backward pass reaches this point `MARK_CALLBACK(tag)` fires (once, after ALL of xs have
    their gradients).  Skipped entirely outside an engine step so the eager module path keeps its autograd graph."""
    if MARK_CALLBACK is None or not any(x.requires_grad for x in xs):
        return xs
    return _MarkFn.apply(tag, *xs)


def require_cuda(t: torch.Tensor):
    if not t.is_cuda:
        raise RuntimeError('anatomask_b200 runs on a B200 only: got a CPU tensor (there is no CPU fallback)')


class MaskCtx:
    """Per-forward visibility state: replaces the reference's `_cur_active` + `_get_active_ex_or_ii` (P/encoder3D.py:5-10).

    active: (N, fd, fh, fw) uint8 on device; the active-patch work-list is built once on the device."""

    def __init__(self, active_b1fff: torch.Tensor):
        require_cuda(active_b1fff)
        a = active_b1fff
        if a.dim() == 5:
            a = a[:, 0]
        self.active = a.to(torch.uint8).contiguous()
        self.N, self.fd, self.fh, self.fw = self.active.shape
        self.frac_hint = 1.0          # visible fraction, host-side hint for FLOP accounting only (never synchronises)
        n = self.active.numel()
        self.list = torch.empty(n, dtype=torch.int32, device=a.device)
        self.count = torch.empty(1, dtype=torch.int32, device=a.device)
        L.call('amb_build_active_list', _p(self.active), n, _p(self.list), _p(self.count), _stream())

    def geo(self, x: torch.Tensor, sparse: bool) -> L.Geo:
        N, D, H, W, Cc = x.shape
        return L.Geo(N, D, H, W, Cc, self.fd, self.fh, self.fw, self.active.data_ptr(),
                     self.list.data_ptr() if sparse else 0, self.count.data_ptr() if sparse else 0)


def dense_geo(x: torch.Tensor) -> L.Geo:
    N, D, H, W, Cc = x.shape
    # any mask grid that divides the tensor works for dense iteration; use one patch per voxel row block
    return L.Geo(N, D, H, W, Cc, D, H, W, 0, 0, 0)


# ----------------------------------------------------------------------------------------------------------------
# layout conversion at the module boundary
# ----------------------------------------------------------------------------------------------------------------
def to_internal(x: torch.Tensor) -> torch.Tensor:
    """(N,C,D,H,W) logical tensor → (N,D,H,W,C) bf16 contiguous.  Zero-copy for channels-last bf16 inputs."""
    require_cuda(x)
    xi = x.permute(0, 2, 3, 4, 1)
    if x.dtype == bf16 and xi.is_contiguous():
        return xi
    return _ToInternal.apply(x)


def to_external(xi: torch.Tensor) -> torch.Tensor:
    """(N,D,H,W,C) → logical (N,C,D,H,W) view (channels_last_3d strides, bf16)."""
    return xi.permute(0, 4, 1, 2, 3)


class _ToInternal(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.float().contiguous()
        N, Cc, D, H, W = x.shape
        out = torch.empty((N, D, H, W, Cc), dtype=bf16, device=x.device)
        L.call('amb_ncdhw_f32_to_ndhwc_bf16', _p(x), _p(out), N, Cc, D, H, W, _stream())
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        N, D, H, W, Cc = g.shape
        out = torch.empty((N, Cc, D, H, W), dtype=torch.float32, device=g.device)
        L.call('amb_ndhwc_bf16_to_ncdhw_f32', _p(g), _p(out), N, Cc, D, H, W, _stream())
        return out


def to_ncdhw_f32(xi: torch.Tensor) -> torch.Tensor:
    xi = xi.contiguous()
    N, D, H, W, Cc = xi.shape
    out = torch.empty((N, Cc, D, H, W), dtype=torch.float32, device=xi.device)
    L.call('amb_ndhwc_bf16_to_ncdhw_f32', _p(xi), _p(out), N, Cc, D, H, W, _stream())
    return out


# ----------------------------------------------------------------------------------------------------------------
# convolutions
# ----------------------------------------------------------------------------------------------------------------
# Engine mode: the many small zero-initialised accumulators of a step (Σ/Σ² in fp64, bias / γ / β gradient sums) are carved
# out of one pool that the engine clears with a single memset at the start of the step, instead of ~100 fill launches.
ZERO_POOL = None          # {'buf': uint8 tensor, 'off': int} while a step runs

# Engine mode: masked encoder tensors are not zero-filled where nothing reads them.  Every consumer of a sparse tensor at a
# resolution whose patch edge is >= 8 voxels walks the active-patch work-list and reads at most one voxel beyond a visible
# patch, so clearing that 1-voxel shell (amb_zero_shell) — or nothing at all for tensors only ever read at visible voxels —
# replaces the full-tensor fills (1.0 ms of a 22 ms STUNet-B step).  Off outside an engine step: the module API keeps the
# reference's contract (masked voxels of every sparse layer's output are zero).  AMB_NO_LEAN_ZERO=1 disables it (A/B);
# AMB_POISON=1 fills every such allocation with NaN first, so a consumer that does read beyond the shell shows up as a NaN loss.
LEAN_ZERO = False
POISON = bool(os.environ.get('AMB_POISON'))


class lean_zero:
    def __enter__(self):
        global LEAN_ZERO
        self.prev = LEAN_ZERO
        LEAN_ZERO = not os.environ.get('AMB_NO_LEAN_ZERO')

    def __exit__(self, *exc):
        global LEAN_ZERO
        LEAN_ZERO = self.prev


def _sparse_alloc(like: torch.Tensor, mode: str) -> torch.Tensor:
    """mode 'full': zero-filled; 'shell' / 'none': uninitialised (the caller clears the shell for 'shell')."""
    if mode == 'full':
        return torch.zeros_like(like)
    t = torch.empty_like(like)
    if POISON:
        t.fill_(float('nan'))
    return t


def zero_shell(t: torch.Tensor, m: 'MaskCtx') -> None:
    L.call('amb_zero_shell', C.byref(m.geo(t, True)), _p(t), _stream())


def _zeros_small(numel: int, dtype, device) -> torch.Tensor:
    pool = ZERO_POOL
    if pool is not None and pool['buf'].device == device:
        nbytes = (numel * torch.empty((), dtype=dtype).element_size() + 15) & ~15
        off = pool['off']
        if off + nbytes <= pool['buf'].numel():
            pool['off'] = off + nbytes
            return pool['buf'][off:off + nbytes].view(dtype)[:numel]
    return torch.zeros(numel, dtype=dtype, device=device)


# Engine mode: the packed bf16 copies of ALL conv weights are refreshed by one launch at the start of the step
# (PackPlan.run) and ConvFn looks them up here; outside a step (PACK_CACHE is None) every call packs its own weight.
# PACK_RECORD collects the pack calls of one eager step so that the engine can build the plan without knowing the modules.
PACK_CACHE = None
PACK_RECORD = None


def _pack(w: torch.Tensor, T: int, A: int, B: int, st: int, sa: int, sb: int) -> torch.Tensor:
    key = (w.data_ptr(), T, A, B, sa, sb, st)
    if PACK_CACHE is not None:
        hit = PACK_CACHE.get(key)
        if hit is not None:
            return hit
    out = torch.empty((T, A, B), dtype=bf16, device=w.device)
    L.call('amb_pack_weight', _p(w), _p(out), T, A, B, st, sa, sb, _stream())
    if PACK_RECORD is not None:
        PACK_RECORD.append((key, w))
    return out


def _pack_conv(weight: torch.Tensor, transposed: bool, dgrad: bool) -> torch.Tensor:
    """bf16 operand of a conv weight: forward form [tap][Cout][Cin] or input-gradient form [tap][Cin][Cout].  The strides
    come from the tensor itself: a stock parameter is (Cout, Cin, k, k, k)-contiguous (ConvTranspose3d: (Cin, Cout, 4, 4, 4)),
    an engine parameter is a view of the taps-major arena storage [tap][Cout][Cin] (trainer.ParamArena) — then the forward
    form is a plain dtype conversion and the weight gradient needs no re-layout at all (packed_alias)."""
    if weight.dim() == 2:                      # nn.Linear used as a per-voxel (1x1x1) conv: (Cout, Cin)
        s0, s1 = weight.stride()
        if dgrad:
            return _pack(weight, 1, weight.shape[1], weight.shape[0], 1, s1, s0)
        return _pack(weight, 1, weight.shape[0], weight.shape[1], 1, s0, s1)
    s = weight.stride()
    T = weight.shape[2] * weight.shape[3] * weight.shape[4]
    if T > 1 and not (s[2] == weight.shape[3] * s[3] and s[3] == weight.shape[4] * s[4]):
        weight = weight.contiguous()
        s = weight.stride()
    co, ci = (1, 0) if transposed else (0, 1)
    if dgrad:
        return _pack(weight, T, weight.shape[ci], weight.shape[co], s[4], s[ci], s[co])
    return _pack(weight, T, weight.shape[co], weight.shape[ci], s[4], s[co], s[ci])


def packed_alias(t: Optional[torch.Tensor], transposed: bool) -> Optional[torch.Tensor]:
    """The [tap][Cout][Cin]-contiguous tensor that shares memory with conv weight (or weight gradient) `t`, if `t` is stored
    taps-major (engine arena; any 1×1×1 weight qualifies trivially) — the layout the weight-gradient kernels produce."""
    if t is None or t.dim() != 5:
        return None
    v = t.permute(2, 3, 4, 1, 0) if transposed else t.permute(2, 3, 4, 0, 1)
    if not v.is_contiguous():
        return None
    return v.reshape(-1, v.shape[3], v.shape[4])


def _from_packed(dwp: torch.Tensor, like: torch.Tensor, transposed: bool) -> torch.Tensor:
    """[tap][Cout][Cin] gradient → a tensor of `like`'s shape: zero-copy view when `like` is stored taps-major, else the
    re-layout kernel into the parameter's (contiguous) layout."""
    T, Cout, Cin = dwp.shape
    if packed_alias(like, transposed) is not None:
        v = dwp.view(*like.shape[2:], Cout, Cin)
        return v.permute(4, 3, 0, 1, 2) if transposed else v.permute(3, 4, 0, 1, 2)
    out = torch.empty(like.shape, dtype=torch.float32, device=dwp.device)
    if transposed:
        L.call('amb_unpack_wgrad', _p(dwp), _p(out), T, Cout, Cin, 1, T, Cout * T, _stream())
    else:
        L.call('amb_unpack_wgrad', _p(dwp), _p(out), T, Cout, Cin, 1, Cin * T, T, _stream())
    return out


class PackPlan:
    """One-launch packing of every conv weight a step uses (amb_pack_weights_batched).  Built from the (key, weight)
    pairs recorded during one eager step; the weights are views of the parameter arenas, so their addresses are stable.
    `group_of(weight)` splits the jobs into separately launched groups (the engine packs the teacher's weights on the main
    stream and the student's — not needed before the student forward, a teacher forward later — on the side stream)."""

    def __init__(self, record, group_of=None):
        import numpy as np
        seen, self.cache, self._keep = set(), {}, []
        jobs, tiles = {}, {}
        for key, w in record:
            if key in seen:
                continue
            ptr, T, A, B, sa, sb, st = key
            # job kinds (amb_pack_job.b_fast): 2 / 3 = taps-major source (engine arena) → plain conversion / per-tap transpose;
            # 1 / 0 = taps innermost (stock parameter layout), b or a the faster channel axis
            if B % 2 == 0 and (T == 1 or st == A * B) and sa == B and sb == 1:
                b_fast, tiles_b, tchunks = 2, 0, 0
                n_tiles = -(-(T * A * B) // 2048)
            elif B % 2 == 0 and (T == 1 or st == A * B) and sa == 1 and sb == A:
                b_fast, tiles_b, tchunks = 3, -(-B // 64), -(-A // 32)
                n_tiles = T * tiles_b * tchunks
            elif st == 1 and B % 2 == 0 and (sb == T and sa == B * T or sa == T and sb == A * T):
                b_fast = 1 if (sb == T and sa == B * T) else 0
                ta, tb = (4, 64) if b_fast else (16, 16)
                tiles_a, tiles_b, tchunks = -(-A // ta), -(-B // tb), -(-T // 32)
                n_tiles = tiles_a * tiles_b * tchunks
            else:
                continue                                   # layout outside the tiled kernels: stays a per-call pack
            seen.add(key)
            grp = 0 if group_of is None else int(group_of(w))
            out = torch.empty((T, A, B), dtype=bf16, device=w.device)
            tile = tiles.get(grp, 0)
            jobs.setdefault(grp, []).append((ptr, out.data_ptr(), T, A, B, b_fast, tile, tiles_b, tchunks, 0))
            tiles[grp] = tile + n_tiles
            self.cache[key] = out
            self._keep.append(w)
        dt = np.dtype([('src', np.uint64), ('dst', np.uint64), ('T', np.int32), ('A', np.int32), ('B', np.int32),
                       ('b_fast', np.int32), ('tile_begin', np.int32), ('tiles_b', np.int32), ('tchunks', np.int32),
                       ('pad', np.int32)])
        assert dt.itemsize == 48
        dev = record[0][1].device if record else 'cuda'
        self.groups = {}
        for grp, js in jobs.items():
            table = torch.from_numpy(np.array(js, dtype=dt).view(np.uint8).copy()).to(dev)
            self.groups[grp] = (table, len(js), tiles[grp])
        self.n_jobs = sum(v[1] for v in self.groups.values())
        self.total_tiles = sum(v[2] for v in self.groups.values())
        self.table = self.groups[0][0] if (len(self.groups) == 1 and 0 in self.groups) else None

    def run(self, group=None):
        for grp, (table, n_jobs, total_tiles) in self.groups.items():
            if group is None or grp == group:
                L.call('amb_pack_weights_batched', _p(table), n_jobs, total_tiles, _stream())


def _conv_call(op, impl, dims, Cin, Cout, k, stride, x, y, w, bias=None, m: Optional[MaskCtx] = None, sparse=False,
               stats=None, ep_scale=None, ep_act=0):
    N, D, H, W = dims
    a = L.ConvArgs(op, impl, N, D, H, W, Cin, Cout, k, stride, x.data_ptr(), y.data_ptr(), w.data_ptr(),
                   0 if bias is None else bias.data_ptr(), 0 if m is None else m.active.data_ptr(),
                   1 if m is None else m.fd, 1 if m is None else m.fh, 1 if m is None else m.fw,
                   m.list.data_ptr() if (m is not None and sparse) else 0,
                   m.count.data_ptr() if (m is not None and sparse) else 0,
                   0 if stats is None else stats.data_ptr(), 0 if ep_scale is None else ep_scale.data_ptr(), ep_act, 0,
                   0, 0, torch.cuda.current_stream().cuda_stream)
    need = L.load().amb_conv_workspace_bytes(C.byref(a))       # > 0: small spatial extent, the kernel splits the taps over CTAs
    if need > 0:
        ws = torch.zeros(need // 4, dtype=torch.float32, device=y.device)
        a.workspace, a.workspace_bytes = ws.data_ptr(), need
    L.call('amb_conv', C.byref(a))


def column_sums(x: torch.Tensor, m: Optional[MaskCtx]) -> torch.Tensor:
    """Σ over voxels per channel (fp32) — bias gradients."""
    Cc = x.shape[-1]
    sums = _zeros_small(2 * Cc, torch.float64, x.device)
    g = m.geo(x, True) if m is not None else dense_geo(x)
    L.call('amb_norm_stats', C.byref(g), _p(x), _p(sums), _stream())
    return sums[:Cc].float()


class GradSumTap:
    """Carries Σ_voxels dx (per channel) from the input-gradient kernel of one conv to the bias gradient of the layer that
    produced its input.  ∂loss/∂bias of the decoder's ConvTranspose3d (P/decoder3D.py:17) is the per-channel sum of the
    gradient w.r.t. its output — which is exactly what the following conv's dgrad kernel writes, so its Σ epilogue delivers
    the sum without another 537 MB pass over that tensor."""

    def __init__(self):
        self.sums = None


class ConvFn(torch.autograd.Function):
    """nn.Conv3d (k∈{1,3}, stride∈{1,2}, pad k//2) and nn.ConvTranspose3d (k4 s2 p1) on channels-last bf16.

    With a MaskCtx the output is zero outside visible patches (SparseConv3d, P/encoder3D.py:12-15) and — where a
    2×8×8 tile fits a patch — masked tiles are never computed."""

    @staticmethod
    def forward(ctx, x, weight, bias, k, stride, m, transposed, impl, stats=None, zero_inactive=True, zero_bias_grad=False,
                zero_dx=True, tap=None):
        require_cuda(x)
        x = x.contiguous()
        N, D, H, W, Cin = x.shape
        k3 = k * k * k
        if transposed:
            Cout = weight.shape[1]
            wp = _pack_conv(weight, True, False)
            y = torch.empty((N, 2 * D, 2 * H, 2 * W, Cout), dtype=bf16, device=x.device)
            flops = 2.0 * N * D * H * W * 64 * Cin * Cout
            with _Timed('convT_fwd', flops):
                _conv_call(L.OP_CONVT, impl, (N, D, H, W), Cin, Cout, 4, 2, x, y, wp, bias, stats=stats)
        else:
            Cout = weight.shape[0]
            wp = _pack_conv(weight, False, False)
            shape = (N, D // stride, H // stride, W // stride, Cout)
            # masked voxels are never written where the work-list skips whole tiles: zero-fill unless the only consumer
            # (a pooled masked norm) visits visible voxels exclusively
            y = torch.zeros(shape, dtype=bf16, device=x.device) if (m is not None and zero_inactive) else \
                torch.empty(shape, dtype=bf16, device=x.device)
            flops = 2.0 * N * (D // stride) * (H // stride) * (W // stride) * k3 * Cin * Cout * \
                (m.frac_hint if m is not None else 1.0)
            with _Timed('conv_fwd', flops):
                _conv_call(L.OP_CONV, impl, (N, D, H, W), Cin, Cout, k, stride, x, y, wp, bias, m, sparse=m is not None,
                           stats=stats)
        ctx.save_for_backward(x, weight)
        ctx.bias_ref = bias
        ctx.weight_ref = weight
        ctx.flops = flops
        ctx.cfg = (k, stride, m, transposed, impl, bias is not None)
        ctx.zero_bias_grad = zero_bias_grad
        ctx.zero_dx = zero_dx
        ctx.tap = tap           # conv: publishes Σdx here; ConvTranspose: takes its bias gradient from it (GradSumTap)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        k, stride, m, transposed, impl, has_bias = ctx.cfg
        bias_ref = ctx.bias_ref
        dy = dy.contiguous()
        N, D, H, W, Cin = x.shape
        k3 = k * k * k
        dx = dw = db = None
        need_w = ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2])
        main = torch.cuda.current_stream()
        side = _side_stream(x.device) if (need_w and (ctx.needs_input_grad[0] or getattr(ctx, 'force_side', False))) else None
        if side is not None:
            side.wait_stream(main)

        def weight_branch(dst_w=None):
            # dst_w: the parameter's .grad view in the engine's (pre-zeroed) gradient arena.  Stored taps-major it IS the
            # [tap][Cout][Cin] buffer the kernels accumulate into (no staging buffer, no fill, no re-layout pass); a
            # contiguous destination goes through a staging buffer and amb_unpack_wgrad.
            dw_ = db_ = None
            Cout = weight.shape[1] if transposed else weight.shape[0]
            T = 64 if transposed else k3
            if ctx.needs_input_grad[1]:
                direct = packed_alias(dst_w, transposed)
                dwp = direct if direct is not None else torch.zeros((T, Cout, Cin), dtype=torch.float32, device=x.device)
                if transposed:
                    a = L.WgradArgs(L.OP_CONVT, impl, N, D, H, W, Cin, Cout, 4, 2, x.data_ptr(), dy.data_ptr(),
                                    dwp.data_ptr(), 1, 1, 1, 0, 0, torch.cuda.current_stream().cuda_stream)
                else:
                    a = L.WgradArgs(L.OP_CONV, impl, N, D, H, W, Cin, Cout, k, stride, x.data_ptr(), dy.data_ptr(),
                                    dwp.data_ptr(), 1 if m is None else m.fd, 1 if m is None else m.fh,
                                    1 if m is None else m.fw, 0 if m is None else m.list.data_ptr(),
                                    0 if m is None else m.count.data_ptr(), torch.cuda.current_stream().cuda_stream)
                with _Timed('convT_wgrad' if transposed else 'conv_wgrad', ctx.flops):
                    L.call('amb_conv_wgrad', C.byref(a))
                if direct is not None:
                    dw_ = dst_w
                elif dst_w is not None:
                    L.call('amb_unpack_wgrad', _p(dwp), _p(dst_w), T, Cout, Cin, 1, *((T, Cout * T) if transposed else (Cin * T, T)),
                           _stream())
                    dw_ = dst_w
                else:
                    dw_ = _from_packed(dwp, weight, transposed)
            if has_bias and ctx.needs_input_grad[2]:
                tap = getattr(ctx, 'tap', None)
                if transposed and tap is not None and tap.sums is not None:
                    db_ = tap.sums[:dy.shape[-1]].float()          # Σ over voxels of dy, from the consumer's dgrad epilogue
                    tap.sums = None
                elif ctx.zero_bias_grad:
                    # the conv feeds a batch-statistics norm: dy is that norm's input gradient, whose per-channel sum over the
                    # pooled voxels is identically zero (the mean subtraction) — the reference's own value here is fp32
                    # rounding noise (~1e-9 relative to the weights' gradients).  No pass over dy.
                    db_ = _zeros_small(dy.shape[-1], torch.float32, dy.device)
                else:
                    db_ = column_sums(dy, m if not transposed else None)
            return dw_, db_

        deferred = False
        if side is not None:
            wref = ctx.weight_ref
            can_defer = DEFER_WGRAD and wref.grad is not None and \
                (wref.grad.is_contiguous() or packed_alias(wref.grad, transposed) is not None) and \
                (bias_ref is None or not has_bias or bias_ref.grad is not None)
            with torch.cuda.stream(side):
                dw, db = weight_branch(wref.grad if can_defer else None)
                if can_defer:
                    if db is not None and not ctx.zero_bias_grad:     # (an identically-zero bias gradient: the arena is already zero)
                        bias_ref.grad.copy_(db)
                    deferred = True
            for t in (dw, db, dy, x):
                if t is not None:
                    t.record_stream(side)
            if deferred:
                dw = db = None
        if ctx.needs_input_grad[0]:
            if transposed:
                Cout = weight.shape[1]
                wp = _pack_conv(weight, True, True)
                dx = torch.empty_like(x)
                with _Timed('convT_dgrad', ctx.flops):
                    _conv_call(L.OP_CONVT_DGRAD, impl, (N, D, H, W), Cin, Cout, 4, 2, dy, dx, wp)
            else:
                Cout = weight.shape[0]
                wp = _pack_conv(weight, False, True)
                # a masked input gradient is zero-filled unless the caller vouches that it is only read at visible voxels
                # (zero_dx=False: list-walking kernels skip masked tiles, dense-walking ones write every voxel themselves);
                # a 1x1 stride-2 conv only ever produces the even voxels
                need_zero = (m is not None and ctx.zero_dx) or (k == 1 and stride == 2)
                dx = torch.zeros_like(x) if need_zero else _sparse_alloc(x, 'none')
                tap = getattr(ctx, 'tap', None)
                dstats = None
                if tap is not None and m is None and fused_stats_ok(Cin, Cout) and impl != L.IMPL_DIRECT:
                    dstats = tap.sums = _zeros_small(2 * Cin + 1, torch.float64, x.device)
                with _Timed('conv_dgrad', ctx.flops):
                    _conv_call(L.OP_CONV_DGRAD, impl, (N, D, H, W), Cin, Cout, k, stride, dy, dx, wp, None, m,
                               sparse=m is not None, stats=dstats)
        if side is not None and not deferred:
            main.wait_stream(side)
        elif side is None and need_w:
            dw, db = weight_branch()
        return dx, dw, db, None, None, None, None, None, None, None, None, None, None


class _SubCtx:
    """Stand-in for an autograd ctx so that ConvFn.forward / backward can serve as building blocks of ConvPairFn."""

    def __init__(self):
        self.saved_tensors = ()
        self.needs_input_grad = (False,) * 13
        self.tap = None

    def save_for_backward(self, *ts):
        self.saved_tensors = ts


class ConvPairFn(torch.autograd.Function):
    """conv1 (k3) and the 1x1 shortcut conv3 of a residual block, both reading the block input with the same stride
    (P/STUNet_head.py:79-80,96-101).  One autograd node, so that the input gradient is formed in place: dx = dgrad(conv1),
    then the shortcut's contribution — which lands on the even voxels only at stride 2 — is added by amb_add_parity0 instead
    of a zero-filled full-resolution tensor plus autograd's full-resolution add.  Engine mode only (ops.LEAN_ZERO)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w3, b3, k, stride, m, impl, stats):
        c1, c3 = _SubCtx(), _SubCtx()
        y = ConvFn.forward(c1, x, w1, b1, k, stride, m, False, impl, stats, False, True, False)
        sc = ConvFn.forward(c3, x, w3, b3, 1, stride, m, False, impl, None, False, False, False)
        x = c1.saved_tensors[0]
        ctx.save_for_backward(x, w1, w3)
        ctx.subs = (c1, c3)
        return y, sc

    @staticmethod
    def backward(ctx, dy, dsc):
        x, w1, w3 = ctx.saved_tensors
        c1, c3 = ctx.subs
        nig = ctx.needs_input_grad
        c1.saved_tensors, c3.saved_tensors = (x, w1), (x, w3)
        c1.needs_input_grad = (nig[0], nig[1], nig[2]) + (False,) * 9
        c3.needs_input_grad = (False, nig[3], nig[4]) + (False,) * 9
        c3.force_side = True
        k, stride, m, _, impl, _ = c1.cfg
        dx, dw1, db1 = ConvFn.backward(c1, dy)[:3]
        dw3, db3 = ConvFn.backward(c3, dsc)[1:3]
        if nig[0]:
            dsc = dsc.contiguous()
            N, D, H, W, Cin = x.shape
            Cout = w3.shape[0]
            wp = _pack_conv(w3, False, True)
            if stride == 2:
                # the shortcut is a per-voxel (Cout -> Cin) product on the coarse grid; its result belongs to the even voxels
                coarse = torch.zeros((N, D // 2, H // 2, W // 2, Cin), dtype=bf16, device=x.device) if m is not None else \
                    torch.empty((N, D // 2, H // 2, W // 2, Cin), dtype=bf16, device=x.device)
                with _Timed('conv_dgrad', c3.flops):
                    _conv_call(L.OP_CONV_DGRAD, impl, (N, D // 2, H // 2, W // 2), Cin, Cout, 1, 1, dsc, coarse, wp, None, m,
                               sparse=m is not None)
                g = m.geo(coarse, True) if m is not None else dense_geo(coarse)
                L.call('amb_add_parity0', C.byref(g), _p(coarse), _p(dx), _stream())
            else:
                dx3 = torch.zeros_like(x) if m is not None else torch.empty_like(x)
                with _Timed('conv_dgrad', c3.flops):
                    _conv_call(L.OP_CONV_DGRAD, impl, (N, D, H, W), Cin, Cout, 1, 1, dsc, dx3, wp, None, m, sparse=m is not None)
                L.call('amb_add', _p(dx), _p(dx3), _p(dx), dx.numel(), _stream())
        return dx, dw1, db1, dw3, db3, None, None, None, None, None


def conv3d_pair(x, w1, b1, w3, b3, k, stride, m, impl=L.IMPL_AUTO, stats=None):
    return ConvPairFn.apply(x, w1, b1, w3, b3, k, stride, m, impl, stats)


def fused_stats_ok(cin: int, cout: int) -> bool:
    """The Σy/Σy² epilogue exists in the tcgen05 kernel only (channel counts it takes: multiples of 16)."""
    return cin % 16 == 0 and cout % 16 == 0


def new_stats(channels: int, device) -> torch.Tensor:
    """(Σy[C], Σy²[C], n) accumulator handed to a conv (fused epilogue) and then to the norm that follows it."""
    return _zeros_small(2 * channels + 1, torch.float64, torch.device(device))


def conv3d(x, weight, bias=None, k=3, stride=1, m: Optional[MaskCtx] = None, impl=L.IMPL_AUTO, stats=None,
           zero_inactive=True, zero_bias_grad=False, zero_dx=True, dx_sum: Optional[GradSumTap] = None):
    """zero_bias_grad: the caller guarantees the output goes ONLY into a norm that uses batch statistics (then ∂loss/∂bias ≡ 0).
    zero_dx=False: the caller guarantees the input gradient is only read at visible voxels (see LEAN_ZERO)."""
    return ConvFn.apply(x, weight, bias, k, stride, m, False, impl, stats, zero_inactive, zero_bias_grad, zero_dx, dx_sum)


def conv_transpose3d(x, weight, bias=None, impl=L.IMPL_AUTO, bias_grad_from: Optional[GradSumTap] = None):
    return ConvFn.apply(x, weight, bias, 4, 2, None, True, impl, None, True, False, True, bias_grad_from)


class StemFn(torch.autograd.Function):
    """Cin = 1 stage-0 pair: conv1 (k3) and the 1×1 shortcut conv3 on the masked fp32 input, one pass
    (P/STUNet_head.py:96-101 with P/spark3D.py:104-107 folded in)."""

    @staticmethod
    def forward(ctx, inp, w1, b1, w3, b3, m: MaskCtx, zero_inactive=True):
        require_cuda(inp)
        inp = inp.float().contiguous()
        N, _, D, H, W = inp.shape
        Cc = w1.shape[0]
        alloc = torch.zeros if zero_inactive else torch.empty
        out1 = alloc((N, D, H, W, Cc), dtype=bf16, device=inp.device)
        out3 = alloc((N, D, H, W, Cc), dtype=bf16, device=inp.device)
        L.call('amb_stem_fwd', _p(inp), _p(m.active), _p(m.list), _p(m.count), N, D, H, W, m.fd, m.fh, m.fw, Cc,
               _p(w1), _p(b1), _p(w3), _p(b3), _p(out1), _p(out3), _stream())
        ctx.save_for_backward(inp)
        ctx.m, ctx.C = m, Cc
        ctx.param_refs = (w1, b1, w3, b3)
        return out1, out3

    @staticmethod
    def backward(ctx, d1, d3):
        (inp,) = ctx.saved_tensors
        m, Cc = ctx.m, ctx.C
        N, _, D, H, W = inp.shape
        d1, d3 = d1.contiguous(), d3.contiguous()
        refs = ctx.param_refs
        if DEFER_WGRAD and all(isinstance(r, torch.nn.Parameter) and r.grad is not None and r.grad.is_contiguous() for r in refs):
            # engine mode: accumulate straight into the (pre-zeroed) arena .grad views
            L.call('amb_stem_wgrad', _p(inp), _p(m.active), _p(m.list), _p(m.count), N, D, H, W, m.fd, m.fh, m.fw, Cc,
                   _p(d1), _p(d3), _p(refs[0].grad), _p(refs[1].grad), _p(refs[2].grad), _p(refs[3].grad), _stream())
            return None, None, None, None, None, None, None
        g = torch.zeros(Cc * 30, dtype=torch.float32, device=inp.device)
        dw1, db1, dw3, db3 = g[:Cc * 27], g[Cc * 27:Cc * 28], g[Cc * 28:Cc * 29], g[Cc * 29:]
        L.call('amb_stem_wgrad', _p(inp), _p(m.active), _p(m.list), _p(m.count), N, D, H, W, m.fd, m.fh, m.fw, Cc,
               _p(d1), _p(d3), _p(dw1), _p(db1), _p(dw3), _p(db3), _stream())
        return None, dw1.view(Cc, 1, 3, 3, 3), db1, dw3.view(Cc, 1, 1, 1, 1), db3, None, None


class ProjFn(torch.autograd.Function):
    """Conv3d(C→1, k1, bias): (N,D,H,W,C) bf16 → rec fp32 (N,1,D,H,W)  (P/decoder3D.py:51,61)."""

    @staticmethod
    def forward(ctx, x, w, b):
        x = x.contiguous()
        N, D, H, W, Cc = x.shape
        rec = torch.empty((N, 1, D, H, W), dtype=torch.float32, device=x.device)
        L.call('amb_proj_fwd', _p(x), _p(w), _p(b), _p(rec), N * D * H * W, Cc, _stream())
        ctx.save_for_backward(x, w)
        ctx.param_refs = (w, b)
        return rec

    @staticmethod
    def backward(ctx, drec):
        x, w = ctx.saved_tensors
        drec = drec.contiguous()
        N, D, H, W, Cc = x.shape
        dx = torch.empty_like(x)
        w_ref, b_ref = ctx.param_refs
        if DEFER_WGRAD and w_ref.grad is not None and b_ref.grad is not None and w_ref.grad.is_contiguous():
            # engine mode: the kernel accumulates into the (pre-zeroed) arena .grad views directly
            L.call('amb_proj_bwd', _p(x), _p(w), _p(drec), _p(dx), _p(w_ref.grad), _p(b_ref.grad), N * D * H * W, Cc, _stream())
            return dx, None, None
        g = torch.zeros(Cc + 1, dtype=torch.float32, device=x.device)
        L.call('amb_proj_bwd', _p(x), _p(w), _p(drec), _p(dx), _p(g[:Cc]), _p(g[Cc:]), N * D * H * W, Cc, _stream())
        return dx, g[:Cc].view_as(w), g[Cc:]


# ----------------------------------------------------------------------------------------------------------------
# pooled masked norm / BatchNorm / densify fill
# ----------------------------------------------------------------------------------------------------------------
class NormFn(torch.autograd.Function):
    """y = act(γ·(x−μ)/√(σ²+eps) + β [+ residual]) with μ, σ² (biased) pooled over the visited voxels of the local batch.

    m given, token None  : SparseInstanceNorm / SparseBatchNorm3d (P/encoder3D.py:17-25,149-158) — visible voxels only
    m given, token given : densify — normalise visible voxels, fill masked ones with the mask token (P/spark3D.py:117-122)
    m None               : nn.BatchNorm3d in training mode (decoder)
    running = (rm, rv, nbt) updates the running statistics in place (momentum 0.1, unbiased variance).
    group (torch.distributed process group) pools the statistics over all ranks: one all-reduce of (Σx, Σx², n) forward
    and one of (Σg, Σg·x̂) backward — SparseSyncBatchNorm3d / nn.SyncBatchNorm (P/encoder3D.py:43, P/decoder3D.py:42)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, residual, token, eps, act, m, running, momentum, group=None, sums=None,
                zero=('full', 'full', 'full')):
        x = x.contiguous()
        Cc = x.shape[-1]
        dev = x.device
        sparse = m is not None
        g = m.geo(x, True) if sparse else dense_geo(x)
        if sums is None:            # not produced by the conv epilogue: one read pass over the visited voxels
            sums = _zeros_small(2 * Cc + 1, torch.float64, dev)
            L.call('amb_norm_stats', C.byref(g), _p(x), _p(sums), _stream())
        ntot = None
        if group is not None:
            import torch.distributed as dist
            L.call('amb_count_voxels', C.byref(g), _p(sums[2 * Cc:]), _stream())
            dist.all_reduce(sums, group=group)
            ntot = sums[2 * Cc:]
        ss = torch.empty(4 * Cc, dtype=torch.float32, device=dev)
        scale, shift, saved = ss[:Cc], ss[Cc:2 * Cc], ss[2 * Cc:]
        rm, rv, nbt = running if running is not None else (None, None, None)
        L.call('amb_norm_finalize', C.byref(g), _p(sums), _p(gamma), _p(beta), eps, _p(scale), _p(shift), _p(saved),
               _p(rm), _p(rv), _p(nbt), momentum, _p(ntot), _stream())
        fill = token is not None
        # zero = (output, dx, dresidual) treatment of the masked voxels of a sparse tensor: 'full' zero-fill (the reference's
        # contract), 'shell' (1-voxel shell of the visible patches only) or 'none' — see LEAN_ZERO
        out = _sparse_alloc(x, zero[0]) if (sparse and not fill) else torch.empty_like(x)
        tok = token.reshape(-1).contiguous() if fill else None
        if residual is not None:
            residual = residual.contiguous()
        L.call('amb_norm_apply', C.byref(g), _p(x), _p(scale), _p(shift), _p(residual), _p(tok), act, _p(out),
               _stream())
        if sparse and not fill and zero[0] == 'shell':
            zero_shell(out, m)
        ctx.save_for_backward(x, residual, ss, ntot)
        ctx.zero = zero
        ctx.cfg = (act, m, fill, token.shape if fill else None, group)
        ctx.param_refs = (gamma, beta, token)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, residual, ss, ntot = ctx.saved_tensors
        act, m, fill, tshape, group = ctx.cfg
        dout = dout.contiguous()
        Cc = x.shape[-1]
        dev = x.device
        sparse = m is not None
        g = m.geo(x, True) if sparse else dense_geo(x)
        scale, shift, saved = ss[:Cc], ss[Cc:2 * Cc], ss[2 * Cc:]
        sums = _zeros_small(3 * Cc, torch.float64, dev)
        L.call('amb_norm_bwd_reduce', C.byref(g), _p(dout), _p(x), _p(residual), _p(scale), _p(shift), _p(saved), act,
               int(fill), _p(sums), _p(sums[2 * Cc:]) if fill else C.c_void_p(0), _stream())
        # engine mode: γ / β gradients are written straight into the parameters' (pre-zeroed, arena-resident) .grad views and
        # autograd gets None for them — no temporaries, no accumulate-add launches (each norm parameter is used exactly once)
        g_ref, b_ref, t_ref = ctx.param_refs
        direct = DEFER_WGRAD and group is None and g_ref.grad is not None and b_ref.grad is not None and \
            g_ref.grad.is_contiguous() and b_ref.grad.is_contiguous() and g_ref.grad.dtype == torch.float32
        gb = None if direct else torch.empty(2 * Cc, dtype=torch.float32, device=dev)
        p_dgamma, p_dbeta = (_p(g_ref.grad), _p(b_ref.grad)) if direct else (_p(gb[:Cc]), _p(gb[Cc:]))
        local_gb = None
        if group is not None:       # parameter grads stay rank-local (DDP averages them); dx needs the pooled sums
            import torch.distributed as dist
            local_gb = sums[:2 * Cc].float()
            dist.all_reduce(sums[:2 * Cc], group=group)
        zero = ctx.zero if not fill else ('full', 'full', 'full')
        dx = _sparse_alloc(x, zero[1]) if sparse else torch.empty_like(x)
        dres = None
        if residual is not None:
            dres = _sparse_alloc(x, zero[2]) if sparse else torch.empty_like(x)
        L.call('amb_norm_bwd_apply', C.byref(g), _p(dout), _p(x), _p(residual), _p(scale), _p(shift), _p(saved),
               _p(sums), act, int(fill), _p(dx), _p(dres), p_dgamma, p_dbeta, _p(ntot), _stream())
        if sparse and zero[1] == 'shell':
            zero_shell(dx, m)
        if local_gb is not None:
            dbeta, dgamma = local_gb[:Cc], local_gb[Cc:]
        elif direct:
            dgamma = dbeta = None
        else:
            dgamma, dbeta = gb[:Cc], gb[Cc:]
        dtoken = None
        if fill:
            if direct and t_ref.grad is not None and t_ref.grad.is_contiguous():
                t_ref.grad.view(-1).copy_(sums[2 * Cc:])           # fp64 sums → fp32 .grad, one launch
            else:
                dtoken = sums[2 * Cc:].float().view(tshape)
        return dx, dgamma, dbeta, dres, dtoken, None, None, None, None, None, None, None, None


def masked_norm(x, gamma, beta, eps, m: MaskCtx, act=L.ACT_NONE, residual=None, running=None, momentum=0.0,
                group=None, sums=None, zero=('full', 'full', 'full')):
    return NormFn.apply(x, gamma, beta, residual, None, eps, act, m, running, momentum, group, sums, zero)


def densify_norm_fill(x, gamma, beta, token, eps, m: MaskCtx, running=None, momentum=0.0, group=None):
    return NormFn.apply(x, gamma, beta, None, token, eps, L.ACT_NONE, m, running, momentum, group)


def batch_norm_train(x, gamma, beta, eps, act, running, momentum, group=None, sums=None):
    return NormFn.apply(x, gamma, beta, None, None, eps, act, None, running, momentum, group, sums)


def norm_eval(x, gamma, beta, rm, rv, eps, act, m: Optional[MaskCtx] = None, token=None):
    """Inference-mode BatchNorm (teacher forward, P/pretrain_AntoMask.py:422): running statistics, no grad.
    With m: visible voxels only (SparseBatchNorm3d in eval) and optionally the densify fill."""
    x = x.contiguous()
    Cc = x.shape[-1]
    ss = torch.empty(2 * Cc, dtype=torch.float32, device=x.device)
    L.call('amb_norm_eval', _p(gamma), _p(beta), _p(rm), _p(rv), eps, _p(ss[:Cc]), _p(ss[Cc:]), Cc, _stream())
    fill = token is not None
    out = torch.zeros_like(x) if (m is not None and not fill) else torch.empty_like(x)
    g = m.geo(x, True) if m is not None else dense_geo(x)
    tok = token.reshape(-1).contiguous() if fill else None
    L.call('amb_norm_apply', C.byref(g), _p(x), _p(ss[:Cc]), _p(ss[Cc:]), C.c_void_p(0), _p(tok), act, _p(out),
           _stream())
    return out


def conv3d_bn_eval(x, weight, gamma, beta, rm, rv, eps, act, k=3, impl=L.IMPL_AUTO):
    """Conv3d (no bias, stride 1) followed by an inference-mode BatchNorm (+activation) as ONE kernel: the BN's
    scale = γ/√(σ²+eps) and shift = β − μ·scale are applied per output channel in the conv epilogue, so the normalised
    tensor is written once and never re-read (the teacher's decoder: P/decoder3D.py:19-22 under model_ema.ema.eval(),
    P/pretrain_AntoMask.py:422).  No autograd: inference only."""
    require_cuda(x)
    x = x.contiguous()
    N, D, H, W, Cin = x.shape
    Cout = weight.shape[0]
    ss = torch.empty(2 * Cout, dtype=torch.float32, device=x.device)
    L.call('amb_norm_eval', _p(gamma), _p(beta), _p(rm), _p(rv), eps, _p(ss[:Cout]), _p(ss[Cout:]), Cout, _stream())
    wp = _pack_conv(weight, False, False)
    y = torch.empty((N, D, H, W, Cout), dtype=bf16, device=x.device)
    with _Timed('conv_fwd', 2.0 * N * D * H * W * k ** 3 * Cin * Cout):
        _conv_call(L.OP_CONV, impl, (N, D, H, W), Cin, Cout, k, 1, x, y, wp, ss[Cout:], ep_scale=ss[:Cout], ep_act=act)
    return y


def batch_norm_eval(x, gamma, beta, rm, rv, eps, act):
    return norm_eval(x, gamma, beta, rm, rv, eps, act)


class AddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = a.contiguous(), b.contiguous()
        out = torch.empty_like(a)
        L.call('amb_add', _p(a), _p(b), _p(out), a.numel(), _stream())
        return out

    @staticmethod
    def backward(ctx, g):
        return g, g


# ----------------------------------------------------------------------------------------------------------------
# loss / hard mask / arena optimiser pieces
# ----------------------------------------------------------------------------------------------------------------
class PatchLossFn(torch.autograd.Function):
    """patchify + per-patch normalised masked MSE (P/spark3D.py:130-138).  Returns (loss, per_patch (N,L))."""

    @staticmethod
    def forward(ctx, inp, rec, active_u8, normalize):
        inp, rec = inp.float().contiguous(), rec.float().contiguous()
        N, _, D, H, W = inp.shape
        Lp = (D // 16) * (H // 16) * (W // 16)
        dev = inp.device
        per_patch = torch.empty((N, Lp), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        pstats = torch.empty(2 * N * Lp + 1, dtype=torch.float32, device=dev)
        ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        L.call('amb_patch_loss_fwd', _p(inp), _p(rec), _p(active_u8), N, D, H, W, int(normalize), _p(per_patch),
               _p(loss), _p(pstats), _p(ticket), _stream())
        ctx.save_for_backward(inp, rec, active_u8, pstats)
        ctx.mark_non_differentiable(per_patch)
        return loss, per_patch

    @staticmethod
    def backward(ctx, dloss, _dpp):
        inp, rec, active_u8, pstats = ctx.saved_tensors
        N, _, D, H, W = inp.shape
        drec = torch.empty_like(rec)
        dl = dloss.float().contiguous()
        L.call('amb_patch_loss_bwd', _p(inp), _p(rec), _p(active_u8), _p(pstats), _p(dl), N, D, H, W, _p(drec),
               _stream())
        return None, drec, None, None


def hard_mask(loss_pred: torch.Tensor, len_loss: int, len_keep: int, seed: int = 0, offset: int = 0,
              want_mask: bool = True, want_order: bool = False, offset_dev: Optional[torch.Tensor] = None):
    """Returns (hard (B,len_loss) int32 in ascending-loss order, mask (B,L) uint8 or None[, order (B,L) int32])."""
    loss_pred = loss_pred.float().contiguous()
    B, Lp = loss_pred.shape
    hard = torch.empty((B, max(len_loss, 1)), dtype=torch.int32, device=loss_pred.device)
    mask = torch.empty((B, Lp), dtype=torch.uint8, device=loss_pred.device) if want_mask else None
    order = torch.empty((B, Lp), dtype=torch.int32, device=loss_pred.device) if want_order else None
    L.call('amb_hard_mask', _p(loss_pred), B, Lp, len_loss, len_keep, seed, offset, _p(offset_dev), _p(hard), _p(order),
           _p(mask), _stream())
    if want_order:
        return hard[:, :len_loss], mask, order
    return hard[:, :len_loss], mask


def ema_update_(ema_flat: torch.Tensor, model_flat: torch.Tensor, decay: float):
    L.call('amb_ema_update', _p(ema_flat), _p(model_flat), ema_flat.numel(), float(decay), _stream())


def adamw_step_(p, g, m, v, lr, betas, eps, wd, step, max_norm: Optional[float], gscale: float = 1.0):
    gn = None
    if max_norm is not None:
        gn = torch.zeros(1, dtype=torch.float64, device=p.device)
        L.call('amb_sumsq', _p(g), g.numel(), _p(gn), _stream())
    L.call('amb_adamw_step', _p(p), _p(g), _p(m), _p(v), p.numel(), lr, betas[0], betas[1], eps, wd, step, _p(gn),
           0.0 if max_norm is None else float(max_norm), float(gscale), _stream())
    return gn


# ----------------------------------------------------------------------------------------------------------------
# remaining sparse-layer API: per-voxel group / layer norm, masked pooling, depthwise conv, GELU,
# layer scale — the pieces of P/encoder3D.py:30-37,47-78,181-279 the MedNeXt / ConvNeXt heads use
# ----------------------------------------------------------------------------------------------------------------
class VoxelNormFn(torch.autograd.Function):
    """SparseGroupNorm / SparseConvNeXtLayerNorm: every visible voxel normalised over each channel group
    (P/encoder3D.py:61-68,205-225: nn.GroupNorm / nn.LayerNorm applied to the (N_active, C) matrix of visible voxels)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, groups, eps, m: MaskCtx):
        require_cuda(x)
        x = x.contiguous()
        out = torch.zeros_like(x) if m is not None else torch.empty_like(x)
        g = m.geo(x, True) if m is not None else dense_geo(x)
        L.call('amb_voxel_norm_fwd', C.byref(g), _p(x), _p(gamma), _p(beta), groups, eps, _p(out), _stream())
        ctx.save_for_backward(x, gamma)
        ctx.cfg = (groups, eps, m)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, gamma = ctx.saved_tensors
        groups, eps, m = ctx.cfg
        dout = dout.contiguous()
        dx = torch.zeros_like(x) if m is not None else torch.empty_like(x)
        gb = torch.zeros(2 * x.shape[-1], dtype=torch.float32, device=x.device)
        Cc = x.shape[-1]
        g = m.geo(x, True) if m is not None else dense_geo(x)
        L.call('amb_voxel_norm_bwd', C.byref(g), _p(dout), _p(x), _p(gamma), groups, eps, _p(dx), _p(gb[:Cc]), _p(gb[Cc:]),
               _stream())
        return dx, gb[:Cc], gb[Cc:], None, None, None


class PoolFn(torch.autograd.Function):
    """nn.MaxPool3d / nn.AvgPool3d followed by the mask multiply at the output resolution (P/encoder3D.py:12-15,31-36)."""

    @staticmethod
    def forward(ctx, x, k, s, p, mode, include_pad, divisor, m: Optional[MaskCtx]):
        require_cuda(x)
        x = x.contiguous()
        N, D, H, W, Cc = x.shape
        od, oh, ow = ((v + 2 * p - k) // s + 1 for v in (D, H, W))
        y = torch.empty((N, od, oh, ow, Cc), dtype=bf16, device=x.device)
        mk = (_p(m.active), m.fd, m.fh, m.fw) if m is not None else (C.c_void_p(0), 1, 1, 1)
        L.call('amb_pool3d_fwd', _p(x), _p(y), N, D, H, W, Cc, k, s, p, mode, int(include_pad), int(divisor or 0), *mk, _stream())
        ctx.save_for_backward(x)
        ctx.cfg = (k, s, p, mode, include_pad, divisor, m)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        k, s, p, mode, include_pad, divisor, m = ctx.cfg
        N, D, H, W, Cc = x.shape
        dx = torch.empty_like(x)
        mk = (_p(m.active), m.fd, m.fh, m.fw) if m is not None else (C.c_void_p(0), 1, 1, 1)
        L.call('amb_pool3d_bwd', _p(x), _p(dy.contiguous()), _p(dx), N, D, H, W, Cc, k, s, p, mode, int(include_pad),
               int(divisor or 0), *mk, _stream())
        return dx, None, None, None, None, None, None, None


class MaskedMeanFn(torch.autograd.Function):
    """SparseAdaptiveAvgPooling(1): Σ x·mask / (Σ mask + 1e-6) → (N, C) fp32 (P/encoder3D.py:186-190)."""

    @staticmethod
    def forward(ctx, x, m: MaskCtx):
        require_cuda(x)
        x = x.contiguous()
        mean = torch.empty((x.shape[0], x.shape[-1]), dtype=torch.float32, device=x.device)
        L.call('amb_masked_mean_fwd', C.byref(m.geo(x, False)), _p(x), _p(mean), _stream())
        ctx.m, ctx.shape = m, x.shape
        return mean

    @staticmethod
    def backward(ctx, dmean):
        dx = torch.empty(ctx.shape, dtype=bf16, device=dmean.device)
        g = L.Geo(*ctx.shape, ctx.m.fd, ctx.m.fh, ctx.m.fw, ctx.m.active.data_ptr(), 0, 0)
        L.call('amb_masked_mean_bwd', C.byref(g), _p(dmean.float().contiguous()), _p(dx), _stream())
        return dx, None


class DepthwiseConvFn(torch.autograd.Function):
    """nn.Conv3d(C, C, k, stride, padding=k//2, groups=C) · mask — SparseConv3d on a depthwise layer
    (SparseConvNeXtBlock.dwconv P/encoder3D.py:259; MedNeXtBlock.conv1 P/MedNeXt_head.py:255-262,339-346)."""

    @staticmethod
    def forward(ctx, x, weight, bias, k, stride, m: Optional[MaskCtx]):
        require_cuda(x)
        x = x.contiguous()
        N, D, H, W, Cc = x.shape
        y = torch.empty((N, D // stride, H // stride, W // stride, Cc), dtype=bf16, device=x.device)
        mk = (_p(m.active), m.fd, m.fh, m.fw) if m is not None else (C.c_void_p(0), 1, 1, 1)
        w = weight.detach().float().contiguous()
        L.call('amb_dwconv3d', L.OP_CONV, _p(x), _p(w), _p(None if bias is None else bias.detach().float().contiguous()), _p(y),
               N, D, H, W, Cc, k, stride, *mk, _stream())
        ctx.save_for_backward(x, w)
        ctx.cfg = (k, stride, m, bias is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        k, stride, m, has_bias = ctx.cfg
        dy = dy.contiguous()
        N, D, H, W, Cc = x.shape
        mk = (_p(m.active), m.fd, m.fh, m.fw) if m is not None else (C.c_void_p(0), 1, 1, 1)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            L.call('amb_dwconv3d', L.OP_CONV_DGRAD, _p(dy), _p(w), C.c_void_p(0), _p(dx), N, D, H, W, Cc, k, stride, *mk, _stream())
        if ctx.needs_input_grad[1]:
            dw = torch.zeros_like(w)
            L.call('amb_dwconv3d_wgrad', _p(x), _p(dy), _p(dw), N, D, H, W, Cc, k, stride, *mk, _stream())
        if has_bias and ctx.needs_input_grad[2]:
            db = column_sums(dy, m)          # Σ over visible outputs (the mask multiply's backward)
        return dx, dw, db, None, None, None


class GeluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        require_cuda(x)
        x = x.contiguous()
        out = torch.empty_like(x)
        L.call('amb_gelu', _p(x), C.c_void_p(0), _p(out), x.numel(), _stream())
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        dx = torch.empty_like(x)
        L.call('amb_gelu', _p(x), _p(dout.contiguous()), _p(dx), x.numel(), _stream())
        return dx


class LayerScaleFn(torch.autograd.Function):
    """out = inp + mask·γ_c·x — the tail of SparseConvNeXtBlock (P/encoder3D.py:270-279); m None = no mask, gamma None = 1."""

    @staticmethod
    def forward(ctx, inp, x, gamma, m: Optional[MaskCtx]):
        require_cuda(x)
        inp, x = inp.contiguous(), x.contiguous()
        out = torch.empty_like(x)
        g = m.geo(x, False) if m is not None else dense_geo(x)
        L.call('amb_layer_scale', C.byref(g), _p(inp), _p(x), _p(gamma), C.c_void_p(0), _p(out), C.c_void_p(0), _stream())
        ctx.save_for_backward(x, gamma)
        ctx.m = m
        return out

    @staticmethod
    def backward(ctx, dout):
        x, gamma = ctx.saved_tensors
        m = ctx.m
        dout = dout.contiguous()
        dxb = torch.empty_like(x)
        dg = torch.zeros(x.shape[-1], dtype=torch.float32, device=x.device) if gamma is not None else None
        g = m.geo(x, False) if m is not None else dense_geo(x)
        L.call('amb_layer_scale', C.byref(g), C.c_void_p(0), _p(x), _p(gamma), _p(dout), _p(dxb), _p(dg), _stream())
        return dout, dxb, dg, None


def voxel_norm(x, gamma, beta, groups, eps, m):
    return VoxelNormFn.apply(x, gamma, beta, groups, eps, m)


def pool3d(x, k, s, p, mode, m=None, include_pad=True, divisor=None):
    return PoolFn.apply(x, k, s, p, mode, include_pad, divisor, m)


def depthwise_conv3d(x, weight, bias, k, stride, m=None):
    return DepthwiseConvFn.apply(x, weight, bias, k, stride, m)


def gelu(x):
    return GeluFn.apply(x)
