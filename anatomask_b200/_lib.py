"""ctypes binding of libanatomask_b200.so (the C ABI declared in include/anatomask_b200.h).

The product path fails loudly when the CUDA library is missing: there is no CPU fallback and no other backend.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('AMB_LIB_PATH') or os.path.join(_HERE, 'libanatomask_b200.so')   # override: A/B kernel builds

vp, i32, i64, f32, f64, u64 = C.c_void_p, C.c_int, C.c_long, C.c_float, C.c_double, C.c_ulonglong

OP_CONV, OP_CONV_DGRAD, OP_CONVT, OP_CONVT_DGRAD = 0, 1, 2, 3
IMPL_AUTO, IMPL_DIRECT, IMPL_TCGEN05, IMPL_TCGEN05_V1 = 0, 1, 2, 3
ACT_NONE, ACT_LRELU, ACT_RELU6 = 0, 1, 2


class ConvArgs(C.Structure):
    _fields_ = [('op', i32), ('impl', i32), ('N', i32), ('D', i32), ('H', i32), ('W', i32), ('Cin', i32),
                ('Cout', i32), ('k', i32), ('stride', i32), ('x', vp), ('y', vp), ('w', vp), ('bias', vp),
                ('active', vp), ('fd', i32), ('fh', i32), ('fw', i32), ('active_list', vp), ('active_count', vp),
                ('stats', vp), ('ep_scale', vp), ('ep_act', i32), ('pad_', i32), ('workspace', vp), ('workspace_bytes', i64),
                ('stream', vp)]


class WgradArgs(C.Structure):
    _fields_ = [('op', i32), ('impl', i32), ('N', i32), ('D', i32), ('H', i32), ('W', i32), ('Cin', i32),
                ('Cout', i32), ('k', i32), ('stride', i32), ('x', vp), ('dy', vp), ('dw', vp), ('fd', i32),
                ('fh', i32), ('fw', i32), ('active_list', vp), ('active_count', vp), ('stream', vp)]


class Geo(C.Structure):
    _fields_ = [('N', i32), ('D', i32), ('H', i32), ('W', i32), ('C', i32), ('fd', i32), ('fh', i32), ('fw', i32),
                ('active', vp), ('active_list', vp), ('active_count', vp)]


_SIGNATURES = {
    'amb_last_error': (C.c_char_p, []),
    'amb_version': (i32, []),
    'amb_sm_arch': (i32, []),
    'amb_launch_count': (i64, []),
    'amb_reset_launch_count': (None, []),
    'amb_last_conv_kernel': (C.c_char_p, []),
    'amb_build_active_list': (i32, [vp, i32, vp, vp, vp]),
    'amb_ncdhw_f32_to_ndhwc_bf16': (i32, [vp, vp, i32, i32, i32, i32, i32, vp]),
    'amb_ndhwc_bf16_to_ncdhw_f32': (i32, [vp, vp, i32, i32, i32, i32, i32, vp]),
    'amb_pack_weight': (i32, [vp, vp, i32, i32, i32, i64, i64, i64, vp]),
    'amb_unpack_wgrad': (i32, [vp, vp, i32, i32, i32, i64, i64, i64, vp]),
    'amb_pack_weights_batched': (i32, [vp, i32, i32, vp]),
    'amb_conv': (i32, [C.POINTER(ConvArgs)]),
    'amb_conv_workspace_bytes': (i64, [C.POINTER(ConvArgs)]),
    'amb_conv_wgrad': (i32, [C.POINTER(WgradArgs)]),
    'amb_stem_fwd': (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]),
    'amb_stem_wgrad': (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]),
    'amb_proj_fwd': (i32, [vp, vp, vp, vp, i64, i32, vp]),
    'amb_proj_bwd': (i32, [vp, vp, vp, vp, vp, vp, i64, i32, vp]),
    'amb_norm_stats': (i32, [C.POINTER(Geo), vp, vp, vp]),
    'amb_norm_finalize': (i32, [C.POINTER(Geo), vp, vp, vp, f32, vp, vp, vp, vp, vp, vp, f32, vp, vp]),
    'amb_count_voxels': (i32, [C.POINTER(Geo), vp, vp]),
    'amb_norm_eval': (i32, [vp, vp, vp, vp, f32, vp, vp, i32, vp]),
    'amb_norm_apply': (i32, [C.POINTER(Geo), vp, vp, vp, vp, vp, i32, vp, vp]),
    'amb_norm_bwd_reduce': (i32, [C.POINTER(Geo), vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp]),
    'amb_norm_bwd_apply': (i32, [C.POINTER(Geo), vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp]),
    'amb_add': (i32, [vp, vp, vp, i64, vp]),
    'amb_zero_shell': (i32, [C.POINTER(Geo), vp, vp]),
    'amb_add_parity0': (i32, [C.POINTER(Geo), vp, vp, vp]),
    'amb_patch_loss_fwd': (i32, [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp]),
    'amb_patch_loss_bwd': (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp]),
    'amb_hard_mask': (i32, [vp, i32, i32, i32, i32, u64, u64, vp, vp, vp, vp, vp]),
    'amb_step_dev': (i32, [vp, vp, i64, vp, vp, vp, vp, i64, vp, vp, i32, i32, vp]),
    'amb_ema_update': (i32, [vp, vp, i64, f64, vp]),
    'amb_sumsq': (i32, [vp, i64, vp, vp]),
    'amb_adamw_step': (i32, [vp, vp, vp, vp, i64, f64, f64, f64, f64, f64, i32, vp, f64, f64, vp]),
    'amb_voxel_norm_fwd': (i32, [C.POINTER(Geo), vp, vp, vp, i32, f32, vp, vp]),
    'amb_voxel_norm_bwd': (i32, [C.POINTER(Geo), vp, vp, vp, i32, f32, vp, vp, vp, vp]),
    'amb_pool3d_fwd': (i32, [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, i32, i32, i32, vp]),
    'amb_pool3d_bwd': (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, i32, i32, i32, vp]),
    'amb_masked_mean_fwd': (i32, [C.POINTER(Geo), vp, vp, vp]),
    'amb_masked_mean_bwd': (i32, [C.POINTER(Geo), vp, vp, vp]),
    'amb_dwconv3d': (i32, [i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, i32, i32, i32, vp]),
    'amb_dwconv3d_wgrad': (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, i32, i32, i32, vp]),
    'amb_gelu': (i32, [vp, vp, vp, i64, vp]),
    'amb_layer_scale': (i32, [C.POINTER(Geo), vp, vp, vp, vp, vp, vp, vp]),
    'amb_aug_spline_prefilter': (i32, [vp, i32, i32, i32, i32, i32, i32, vp, vp, i32, i32, i32, vp]),
    'amb_aug_resample': (i32, [vp, i32, i32, i32, C.POINTER(f64), C.POINTER(i32), f32, vp, i32, i32, i32, vp]),
    'amb_aug_crop_mirror': (i32, [vp, i32, i32, i32, i32, i32, i32, C.POINTER(i32), vp, i32, i32, i32, vp]),
}

EXPORTED = tuple(_SIGNATURES.keys())
_lib = None


def load() -> C.CDLL:
    """Loads the shared library (no compute).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} is missing: build it with `python -m anatomask_b200.build` (nvcc, sm_100a). '
                'anatomask_b200 has no CPU or non-B200 fallback.')
        if os.environ.get('AMB_LIB_PATH') is None:
            # a library older than its sources measures (and tests) something else than the tree says: say so loudly;
            # AMB_STRICT_LIB=1 turns the warning into an error
            try:
                from . import build as _b
                if os.path.isdir(_b.CSRC) and _b.built_hash() != _b.source_hash():
                    msg = (f'{LIB_PATH} was not built from the sources in {_b.CSRC} (source hash differs from the build stamp): '
                           'run `python -m anatomask_b200.build`')
                    if os.environ.get('AMB_STRICT_LIB') == '1':
                        raise RuntimeError(msg)
                    import sys
                    print('[anatomask_b200] WARNING: ' + msg, file=sys.stderr)
            except OSError:
                pass
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class AmbError(RuntimeError):
    pass


def check(code: int) -> None:
    if code != 0:
        raise AmbError(f'libanatomask_b200 error {code}: {load().amb_last_error().decode()}')


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args))
