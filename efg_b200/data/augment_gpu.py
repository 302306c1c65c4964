"""The reference's per-sample training processors for 3-D detection, with the point cloud staying on the GPU
(efg/data/augmentations/extend_3d.py:109-316, config VD/config.yaml:22-46):

    RandomFlip3D(p) -> GlobalRotation(rotation) -> GlobalScaling(min, max) -> FilterByRange(pc_range) -> PointShuffle(p)
    -> Voxelization

The reference runs them in 6 DataLoader worker processes per rank, each pass a numpy / torch round trip over the cloud,
followed by the numba voxelizer with its 362 MB scratch array per call.  Here the random draws are made on the host with
the SAME numpy calls in the SAME order (so a seeded run draws the reference's parameters), the ground-truth boxes (a few
hundred) are transformed on the host with the reference's formulas, and the points take ONE fused transform + filter +
order-preserving compaction on the device (csrc/augment.cu), an optional device shuffle, and the device voxelizer — no
host synchronisation between them: the kept-point count stays on the device and becomes the voxelizer's scene offset.
"""
import math

import numpy as np
import torch

from .. import _lib, ops


def limit_boxes_z(boxes, limit_range):
    """mask_boxes_outside_range_bev_z_bound (efg/geometry/box_ops.py:459-477): centre inside the x-y range and the box
    not entirely below / above the z range (rotation about z leaves the corner heights at z +- dz / 2)."""
    m1 = (boxes[:, 0] >= limit_range[0]) & (boxes[:, 0] <= limit_range[3]) & (boxes[:, 1] >= limit_range[1]) & \
         (boxes[:, 1] <= limit_range[4])
    half = boxes[:, 5] * np.float32(0.5)
    zmax, zmin = boxes[:, 2] + half, boxes[:, 2] - half
    m2 = (zmax < limit_range[2]) ^ (zmin > limit_range[5])
    return m1 & ~m2


class GpuPointAugmentation:
    def __init__(self, pc_range, flip_p=0.5, rotation=0.78539816, min_scale=0.8, max_scale=1.2, translation_std=None,
                 shuffle_p=1.0, filter_gt=True):
        self.pc_range = [float(v) for v in pc_range]
        self.flip_p = flip_p
        self.rotation = list(rotation) if isinstance(rotation, (list, tuple)) else [-rotation, rotation]
        self.min_scale, self.max_scale = min_scale, max_scale
        self.translation_std = translation_std
        self.shuffle_p = shuffle_p
        self.filter_gt = filter_gt

    # ---- random draws: the reference's numpy calls, in its order ---------------------------------------------------
    def draw(self):
        p = self.flip_p
        flip_x = bool(np.random.choice([False, True], replace=False, p=[1 - p, p]))   # extend_3d.py:129
        flip_y = bool(np.random.choice([False, True], replace=False, p=[1 - p, p]))   # :148
        angle = float(np.random.uniform(self.rotation[0], self.rotation[1]))          # :194
        scale = float(np.random.uniform(self.min_scale, self.max_scale))              # :212
        trans = None
        if self.translation_std is not None:
            trans = np.random.normal(scale=np.array(self.translation_std, dtype=np.float32), size=3).T   # :231
        return {"flip_x": flip_x, "flip_y": flip_y, "angle": angle, "scale": scale, "translation": trans}

    # ---- ground truth (host, a few hundred boxes; formulas of extend_3d.py:130-220) --------------------------------
    def transform_annotations(self, ann, d):
        ann = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in ann.items()}
        b = ann["gt_boxes"]
        if d["flip_x"]:
            b[:, 1] = -b[:, 1]
            b[:, -1] = -b[:, -1]
            if b.shape[1] > 7:
                b[:, 7] = -b[:, 7]
        if d["flip_y"]:
            b[:, 0] = -b[:, 0]
            b[:, -1] = -(b[:, -1] + np.pi)
            if b.shape[1] > 7:
                b[:, 6] = -b[:, 6]
        a = np.float32(d["angle"])
        c, s = np.float32(math.cos(a)), np.float32(math.sin(a))
        c, s = np.cos(a, dtype=np.float32), np.sin(a, dtype=np.float32)
        x, y = b[:, 0].copy(), b[:, 1].copy()
        b[:, 0], b[:, 1] = x * c + y * (-s), x * s + y * c
        b[:, -1] += a
        if b.shape[1] > 7:
            vx, vy = b[:, 6].copy(), b[:, 7].copy()
            b[:, 6], b[:, 7] = vx * c + vy * (-s), vx * s + vy * c
        b[:, :-1] *= np.float32(d["scale"])
        if d["translation"] is not None:
            b[:, :3] += d["translation"].astype(np.float32)
        if self.filter_gt:
            keep = limit_boxes_z(b, self.pc_range)
            ann = {k: (v[keep] if isinstance(v, np.ndarray) and v.shape[:1] == keep.shape else v) for k, v in ann.items()}
        return ann

    # ---- points (device) -------------------------------------------------------------------------------------------
    def transform_points(self, points, d, generator=None):
        """points [N, F] f32 CUDA -> (out [N, F] with the kept points first, count int32 [1] on the device)."""
        ops._check(points, "points", torch.float32)
        n, f = points.shape
        out = torch.empty_like(points)
        count = torch.zeros(1, dtype=torch.int32, device=points.device)
        L = _lib.lib()
        ws = ops.workspace(L.efgb_augment_workspace_bytes(n), points.device)
        a = np.float32(d["angle"])
        trans = _lib.f32array(d["translation"]) if d["translation"] is not None else None
        rc = L.efgb_augment_points(ops._p(points), n, f, int(d["flip_x"]), int(d["flip_y"]), float(np.cos(a, dtype=np.float32)),
                                   float(np.sin(a, dtype=np.float32)), float(np.float32(d["scale"])), trans,
                                   _lib.f32array(self.pc_range), ops._p(out), ops._p(count), ops._p(ws), ws.numel(), ops._stream())
        _lib.check(rc, "augment_points")
        if self.shuffle_p > 0 and np.random.uniform(0, 1) <= self.shuffle_p:   # PointShuffle (extend_3d.py:115-118)
            # a uniformly random order of the kept points, dropped slots last: random keys + sort, no host sync
            keys = torch.rand(n, device=points.device, generator=generator)
            keys = torch.where(torch.arange(n, device=points.device) < count, keys, torch.full_like(keys, 2.0))
            out = out[torch.argsort(keys)]
        return out, count

    def __call__(self, points, info):
        """points: CUDA [N, F]; info: {"annotations": {...numpy...}} -> (dict(points, num_points), info) like a processor."""
        d = self.draw()
        out, count = self.transform_points(points, d)
        if "annotations" in info:
            info = dict(info, annotations=self.transform_annotations(info["annotations"], d))
        return {"points": out, "num_points": count}, info


def voxelize_augmented(samples, spec_voxel_size, pc_range, max_points, max_voxels):
    """Batch of GpuPointAugmentation outputs -> the device voxelizer, the kept-point counts feeding its scene offsets
    without leaving the device.  Returns the dict of ops.hard_voxelize_batched."""
    pts = torch.cat([s["points"] for s in samples], 0)
    sizes = torch.tensor([s["points"].shape[0] for s in samples], dtype=torch.int64, device=pts.device)
    counts = torch.cat([s["num_points"] for s in samples]).to(torch.int64)
    # compact the per-scene kept prefixes into one contiguous array: position of every point inside its scene
    starts = torch.cumsum(sizes, 0) - sizes
    idx = torch.arange(pts.shape[0], device=pts.device)
    scene = torch.repeat_interleave(torch.arange(len(samples), device=pts.device), sizes, output_size=pts.shape[0])
    valid = (idx - starts[scene]) < counts[scene]
    order = torch.argsort((~valid).to(torch.int8), stable=True)      # kept points first, scene order preserved
    packed = pts[order].contiguous()
    offsets = torch.cat([counts.new_zeros(1), torch.cumsum(counts, 0)]).to(torch.int32)
    return ops.hard_voxelize_batched(packed, offsets, spec_voxel_size, pc_range, max_points, max_voxels, coors_dim=4,
                                     want_voxels=False, capacity=min(pts.shape[0], len(samples) * max_voxels))
