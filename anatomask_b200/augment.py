"""Device-side input pipeline: crop / pad + spatial augmentation + mirroring producing `inp` in HBM.

What the reference does on 12 CPU worker processes per GPU (P/pretrain_AntoMask.py:312-345):
  nnUNetDataLoader3D.generate_train_batch  (N/training/dataloading/data_loader_3d.py:7-51, bbox from
      N/training/dataloading/base_data_loader.py:64-135)            crop the INITIAL patch, zero-pad outside the case
  SpatialTransform(patch_size, do_elastic_deform=False, do_rotation=True, angle ±30° per axis, p_rot_per_axis=1,
      do_scale=True, scale=(0.7, 1.4), border_mode_data='constant', border_cval_data=0, order_data=3, random_crop=False,
      p_scale_per_sample=0.2, p_rot_per_sample=0.2, independent_scale_for_each_axis=False)        P/pretrain_AntoMask.py:79-91
  MirrorTransform((0, 1, 2))                                                                      P/pretrain_AntoMask.py:112-113
  NumpyToTensor(['data', 'target'], 'float')                                                      P/pretrain_AntoMask.py:148
(the colour / noise / low-resolution transforms are commented out in the scripts; segmentation-only transforms do not
touch `data`).  batchgenerators (>= 0.25, pyproject.toml:39) is a third-party dependency absent from the reference tree:
its published algorithm (batchgenerators/augmentations/spatial_transformations.py::augment_spatial, utils.py::
create_zero_centered_coordinate_mesh / rotate_coords_3d / scale_coords / interpolate_img, and
transforms/spatial_transforms.py::MirrorTransform) is restated here; the random draws follow its order so that a seeded
numpy stream gives the parameters the CPU pipeline would have drawn.

Here the case volumes stay resident in HBM, the host only draws a handful of scalars per sample, and three kernels do the
voxel work (csrc/augment.cu): spline prefilter (fused with crop + pad), spline evaluation (+ mirror), or plain crop + mirror
when no rotation / scaling was drawn (64 % of the samples).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L

DEG30 = 30. / 360 * 2. * np.pi


def rotation_matrix(angle_x: float, angle_y: float, angle_z: float) -> np.ndarray:
    """batchgenerators rotate_coords_3d: R = Rx·Ry·Rz, applied to ROW vectors (coords·R)."""
    cx, sx, cy, sy, cz, sz = (math.cos(angle_x), math.sin(angle_x), math.cos(angle_y), math.sin(angle_y),
                              math.cos(angle_z), math.sin(angle_z))
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], dtype=np.float64)
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=np.float64)
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]], dtype=np.float64)
    return rx @ ry @ rz


def get_patch_size(final_patch_size, rot_x, rot_y, rot_z, scale_range) -> np.ndarray:
    """N/training/data_augmentation/compute_initial_patch_size.py:4-27: the patch the loader must crop so that any drawn
    rotation / zoom-out still finds data under the final patch."""
    lim = 90 / 360 * 2. * np.pi
    rx, ry, rz = (min(lim, max(np.abs(r)) if isinstance(r, (tuple, list)) else r) for r in (rot_x, rot_y, rot_z))
    coords = np.array(final_patch_size, dtype=np.float64)
    shape = coords.copy()
    for a in ((rx, 0, 0), (0, ry, 0), (0, 0, rz)):
        shape = np.max(np.vstack((np.abs(coords @ rotation_matrix(*a)), shape)), 0)
    shape /= min(scale_range)
    return shape.astype(int)


@dataclass
class SampleParams:
    """Everything random about one sample."""
    bbox_lb: Tuple[int, int, int]                     # lower corner of the initial patch in case coordinates
    matrix: Optional[np.ndarray] = None               # 3×3 (rotation · scale); None = nothing drawn → centre crop
    mirror: Tuple[bool, bool, bool] = (False, False, False)
    angles: Tuple[float, float, float] = (0., 0., 0.)
    scale: float = 1.0


def draw_bbox(case_shape: Sequence[int], patch_size: Sequence[int], final_patch_size: Sequence[int], rng,
              force_fg: bool = False, class_locations: Optional[dict] = None) -> Tuple[int, int, int]:
    """nnUNetDataLoaderBase.get_bbox (N/training/dataloading/base_data_loader.py:64-135) without an ignore label:
    uniform lower corner in [-need_to_pad // 2, shape + need_to_pad // 2 + need_to_pad % 2 - patch]; with force_fg a random
    voxel of a random foreground class becomes the patch centre (clamped to the lower bound only, like the reference)."""
    dim = len(case_shape)
    need = [int(patch_size[d]) - int(final_patch_size[d]) for d in range(dim)]
    for d in range(dim):
        if need[d] + case_shape[d] < patch_size[d]:
            need[d] = patch_size[d] - case_shape[d]
    lbs = [-need[d] // 2 for d in range(dim)]
    ubs = [case_shape[d] + need[d] // 2 + need[d] % 2 - patch_size[d] for d in range(dim)]
    if force_fg:
        assert class_locations is not None, 'if force_fg is set class_locations cannot be None'
        eligible = [k for k in class_locations.keys() if len(class_locations[k]) > 0]
        if eligible:
            cls = eligible[rng.choice(len(eligible))]
            vox = class_locations[cls][rng.choice(len(class_locations[cls]))]
            return tuple(max(lbs[d], int(vox[d + 1]) - patch_size[d] // 2) for d in range(dim))
    return tuple(int(rng.randint(lbs[d], ubs[d] + 1)) for d in range(dim))


def draw_spatial(rng, angle=(-DEG30, DEG30), scale=(0.7, 1.4), p_rot=0.2, p_scale=0.2, p_rot_per_axis=1.0):
    """One sample's draws of batchgenerators augment_spatial, in its order (elastic deformation is off and draws nothing)."""
    mat, angles, sc = None, (0., 0., 0.), 1.0
    if rng.uniform() < p_rot:
        a = []
        for _ in range(3):
            a.append(rng.uniform(angle[0], angle[1]) if rng.uniform() <= p_rot_per_axis else 0.)
        angles = tuple(a)
        mat = rotation_matrix(*angles)
    if rng.uniform() < p_scale:
        if rng.random() < 0.5 and scale[0] < 1:
            sc = rng.uniform(scale[0], 1)
        else:
            sc = rng.uniform(max(scale[0], 1), scale[1])
        mat = (np.eye(3) if mat is None else mat) * sc
    return mat, angles, sc


def draw_mirror(rng, axes=(0, 1, 2), p_per_sample=1.0):
    """MirrorTransform.__call__ + augment_mirroring for one sample: one draw for the sample, one per axis."""
    flips = [False, False, False]
    if rng.uniform() < p_per_sample:
        for ax in (0, 1, 2):
            if ax in axes and rng.uniform() < 0.5:
                flips[ax] = True
    return tuple(flips)


class DeviceAugmenter:
    """`__call__(cases, params)` → inp (B, 1, *patch_size) fp32 on the cases' device.

    cases: one single-channel fp32 CUDA tensor (D, H, W) per sample (the preprocessed case, resident in HBM).
    params: one SampleParams per sample — from `draw(...)` or from the oracle, so both sides can share the draws."""

    def __init__(self, patch_size=(128, 128, 128), angle=(-DEG30, DEG30), scale=(0.7, 1.4), p_rot=0.2, p_scale=0.2,
                 mirror_axes=(0, 1, 2), patch_scale_range=(0.85, 1.25), oversample_foreground_percent=0.33):
        self.patch_size = tuple(int(p) for p in patch_size)
        self.angle, self.scale, self.p_rot, self.p_scale, self.mirror_axes = angle, scale, p_rot, p_scale, tuple(mirror_axes)
        a = (angle[0], angle[1])
        self.initial_patch_size = tuple(int(v) for v in get_patch_size(self.patch_size, a, a, a, patch_scale_range))
        self.oversample = oversample_foreground_percent
        self._buf = {}

    def draw(self, case_shapes: Sequence[Sequence[int]], rng=np.random, class_locations: Optional[List[dict]] = None) -> List[SampleParams]:
        """The batch's parameters in the order the reference pipeline consumes its numpy stream: loader (one bbox per sample),
        SpatialTransform (all samples), MirrorTransform (all samples)."""
        B = len(case_shapes)
        boxes = []
        for j, shp in enumerate(case_shapes):
            force_fg = class_locations is not None and not j < round(B * (1 - self.oversample))
            boxes.append(draw_bbox(shp, self.initial_patch_size, self.patch_size, rng, force_fg,
                                   class_locations[j] if class_locations is not None else None))
        spatial = [draw_spatial(rng, self.angle, self.scale, self.p_rot, self.p_scale) for _ in range(B)]
        mirrors = [draw_mirror(rng, self.mirror_axes) for _ in range(B)]
        return [SampleParams(boxes[j], spatial[j][0], mirrors[j], spatial[j][1], spatial[j][2]) for j in range(B)]

    def _buffers(self, dev):
        if dev not in self._buf:
            n = self.initial_patch_size[0] * self.initial_patch_size[1] * self.initial_patch_size[2]
            self._buf[dev] = (torch.empty(n, dtype=torch.float32, device=dev), torch.empty(n, dtype=torch.float32, device=dev))
        return self._buf[dev]

    def __call__(self, cases: Sequence[torch.Tensor], params: Sequence[SampleParams], out: Optional[torch.Tensor] = None) -> torch.Tensor:
        assert len(cases) == len(params)
        dev = cases[0].device
        if dev.type != 'cuda':
            raise RuntimeError('DeviceAugmenter runs on a B200 only (there is no CPU fallback)')
        B = len(cases)
        O, P = self.patch_size, self.initial_patch_size
        if out is None:
            out = torch.empty((B, 1) + O, dtype=torch.float32, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for j, (case, p) in enumerate(zip(cases, params)):
            if case.dim() == 4 and case.shape[0] == 1:
                case = case[0]
            if case.dim() != 3 or case.dtype != torch.float32 or not case.is_contiguous():
                raise RuntimeError('DeviceAugmenter: a case is one contiguous single-channel fp32 volume (D, H, W)')
            sD, sH, sW = case.shape
            mir = (C.c_int * 3)(*[int(bool(f)) for f in p.mirror])
            dst = C.c_void_p(out[j].data_ptr())
            if p.matrix is None:
                lb = [p.bbox_lb[d] + (P[d] - O[d]) // 2 for d in range(3)]          # center_crop_aug
                L.call('amb_aug_crop_mirror', C.c_void_p(case.data_ptr()), sD, sH, sW, lb[0], lb[1], lb[2], mir, dst,
                       O[0], O[1], O[2], stream)
                continue
            coef, scratch = self._buffers(dev)
            L.call('amb_aug_spline_prefilter', C.c_void_p(case.data_ptr()), sD, sH, sW, p.bbox_lb[0], p.bbox_lb[1], p.bbox_lb[2],
                   C.c_void_p(coef.data_ptr()), C.c_void_p(scratch.data_ptr()), P[0], P[1], P[2], stream)
            m9 = (C.c_double * 9)(*[float(v) for v in np.asarray(p.matrix, dtype=np.float64).reshape(-1)])
            L.call('amb_aug_resample', C.c_void_p(coef.data_ptr()), P[0], P[1], P[2], m9, mir, 0.0, dst, O[0], O[1], O[2], stream)
        return out

    def sample(self, cases: Sequence[torch.Tensor], rng=np.random, class_locations=None):
        params = self.draw([tuple(c.shape[-3:]) for c in cases], rng, class_locations)
        return self(cases, params), params
