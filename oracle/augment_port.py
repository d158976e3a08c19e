"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product package `anatomask_b200`.

CPU restatement (numpy + scipy) of the reference's per-sample input pipeline for the pre-training scripts
(SURVEY.md §8f row 2):
  N/training/dataloading/data_loader_3d.py:22-49        crop the initial patch at the bounding box, zero-pad outside the case
  P/pretrain_AntoMask.py:79-91                          SpatialTransform as configured there
  P/pretrain_AntoMask.py:112-113                        MirrorTransform((0, 1, 2))
`batchgenerators` (>= 0.25, /root/reference/pyproject.toml:39) is a third-party dependency that is NOT in the reference tree
and not installed here, so the transform bodies follow its published source (batchgenerators/augmentations/
spatial_transformations.py::augment_spatial, augmentations/utils.py::create_zero_centered_coordinate_mesh,
rotate_coords_3d, scale_coords, interpolate_img; transforms/spatial_transforms.py::MirrorTransform): PARITY UNPINNED by any
reference artefact for the transform glue and the order of random draws.  The numerical core is pinned: interpolation is
`scipy.ndimage.map_coordinates(img.astype(float), coords, order=3, mode='constant', cval=0)` — the very call
batchgenerators' interpolate_img makes — executed by the installed scipy.
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import map_coordinates


def crop_and_pad(case: np.ndarray, bbox_lb, patch_size) -> np.ndarray:
    """data_loader_3d.py:29-49 for one single-channel case (D,H,W): valid part of the bbox, then constant-0 padding."""
    shape = case.shape
    ubs = [bbox_lb[i] + patch_size[i] for i in range(3)]
    vlb = [max(0, bbox_lb[i]) for i in range(3)]
    vub = [min(shape[i], ubs[i]) for i in range(3)]
    data = case[tuple(slice(a, b) for a, b in zip(vlb, vub))]
    padding = [(-min(0, bbox_lb[i]), max(ubs[i] - shape[i], 0)) for i in range(3)]
    return np.pad(data, padding, 'constant', constant_values=0)


def zero_centered_mesh(shape) -> np.ndarray:
    tmp = tuple(np.arange(i) for i in shape)
    coords = np.array(np.meshgrid(*tmp, indexing='ij')).astype(float)
    for d in range(len(shape)):
        coords[d] -= ((np.array(shape).astype(float) - 1) / 2.)[d]
    return coords


def rot_x(a): return np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
def rot_y(a): return np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
def rot_z(a): return np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])


def rotate_coords_3d(coords, ax, ay, az):
    m = np.identity(3)
    m = np.dot(m, rot_x(ax))
    m = np.dot(m, rot_y(ay))
    m = np.dot(m, rot_z(az))
    return np.dot(coords.reshape(3, -1).transpose(), m).transpose().reshape(coords.shape)


def augment_spatial_sample(patch: np.ndarray, out_size, angles=None, scale=None, order=3) -> np.ndarray:
    """augment_spatial for one sample / one channel with do_elastic_deform=False, random_crop=False, constant border 0.
    angles / scale None = not drawn."""
    if angles is None and scale is None:
        lb = [(patch.shape[d] - out_size[d]) // 2 for d in range(3)]                     # center_crop_aug
        return patch[tuple(slice(lb[d], lb[d] + out_size[d]) for d in range(3))].astype(np.float32)
    coords = zero_centered_mesh(out_size)
    if angles is not None:
        coords = rotate_coords_3d(coords, *angles)
    if scale is not None:
        coords = coords * scale
    for d in range(3):
        coords[d] += patch.shape[d] / 2. - 0.5
    return map_coordinates(patch.astype(float), coords, order=order, mode='constant', cval=0).astype(np.float32)


def mirror(sample: np.ndarray, flips) -> np.ndarray:
    """augment_mirroring on one (D,H,W) sample."""
    if flips[0]:
        sample = sample[::-1]
    if flips[1]:
        sample = sample[:, ::-1]
    if flips[2]:
        sample = sample[:, :, ::-1]
    return np.ascontiguousarray(sample)


def pipeline_sample(case, bbox_lb, initial_patch_size, out_size, angles, scale, flips):
    patch = crop_and_pad(case, bbox_lb, initial_patch_size)
    return mirror(augment_spatial_sample(patch, out_size, angles, scale), flips)


def draw_batch(rng, case_shapes, initial_patch_size, out_size, angle, scale=(0.7, 1.4), p_rot=0.2, p_scale=0.2):
    """The numpy draws of one batch in the reference's order: loader bboxes (base_data_loader.py:86-88, random branch), then
    augment_spatial per sample, then MirrorTransform per sample.  Returns [(bbox_lb, angles|None, scale|None, flips)]."""
    B = len(case_shapes)
    boxes = []
    for shp in case_shapes:
        need = [initial_patch_size[d] - out_size[d] for d in range(3)]
        for d in range(3):
            if need[d] + shp[d] < initial_patch_size[d]:
                need[d] = initial_patch_size[d] - shp[d]
        lbs = [-need[d] // 2 for d in range(3)]
        ubs = [shp[d] + need[d] // 2 + need[d] % 2 - initial_patch_size[d] for d in range(3)]
        boxes.append(tuple(int(rng.randint(lbs[d], ubs[d] + 1)) for d in range(3)))
    spatial = []
    for _ in range(B):
        angles = sc = None
        if rng.uniform() < p_rot:
            a = []
            for _ax in range(3):
                a.append(rng.uniform(angle[0], angle[1]) if rng.uniform() <= 1 else 0)
            angles = tuple(a)
        if rng.uniform() < p_scale:
            if rng.random() < 0.5 and scale[0] < 1:
                sc = rng.uniform(scale[0], 1)
            else:
                sc = rng.uniform(max(scale[0], 1), scale[1])
        spatial.append((angles, sc))
    flips = []
    for _ in range(B):
        f = [False, False, False]
        if rng.uniform() < 1:
            for ax in range(3):
                if rng.uniform() < 0.5:
                    f[ax] = True
        flips.append(tuple(f))
    return [(boxes[j], spatial[j][0], spatial[j][1], flips[j]) for j in range(B)]
