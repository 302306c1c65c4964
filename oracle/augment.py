"""ORACLE (test infrastructure) — the reference's point-cloud processors restated in numpy with the random draws as
inputs: RandomFlip3D, GlobalRotation, GlobalScaling, FilterByRange (efg/data/augmentations/extend_3d.py:121-316,
efg/geometry/box_ops.py:459-548).  Pinned: tests/golden/augment_*.npz hold the output of the reference's own classes
(imported from /root/reference) for seeded draws."""
import numpy as np


def transform_points(points, flip_x, flip_y, angle, scale, pc_range=None):
    p = points.astype(np.float32).copy()
    if flip_x:
        p[:, 1] = -p[:, 1]
    if flip_y:
        p[:, 0] = -p[:, 0]
    a = np.float32(angle)
    c, s = np.cos(a, dtype=np.float32), np.sin(a, dtype=np.float32)
    x, y = p[:, 0].copy(), p[:, 1].copy()
    p[:, 0] = x * c + y * (-s)      # points[:, :3] @ [[c, s, 0], [-s, c, 0], [0, 0, 1]]
    p[:, 1] = x * s + y * c
    p[:, :3] *= np.float32(scale)
    if pc_range is None:
        return p, np.ones(p.shape[0], dtype=bool)
    keep = ((p[:, 0] >= pc_range[0]) & (p[:, 0] <= pc_range[3]) & (p[:, 1] >= pc_range[1]) & (p[:, 1] <= pc_range[4]) &
            (p[:, 2] >= pc_range[2]) & (p[:, 2] <= pc_range[5]))
    return p[keep], keep


def transform_boxes(boxes, flip_x, flip_y, angle, scale, pc_range=None):
    b = boxes.astype(np.float32).copy()
    if flip_x:
        b[:, 1] = -b[:, 1]
        b[:, -1] = -b[:, -1]
        if b.shape[1] > 7:
            b[:, 7] = -b[:, 7]
    if flip_y:
        b[:, 0] = -b[:, 0]
        b[:, -1] = -(b[:, -1] + np.pi)
        if b.shape[1] > 7:
            b[:, 6] = -b[:, 6]
    a = np.float32(angle)
    c, s = np.cos(a, dtype=np.float32), np.sin(a, dtype=np.float32)
    x, y = b[:, 0].copy(), b[:, 1].copy()
    b[:, 0], b[:, 1] = x * c + y * (-s), x * s + y * c
    b[:, -1] += a
    if b.shape[1] > 7:
        vx, vy = b[:, 6].copy(), b[:, 7].copy()
        b[:, 6], b[:, 7] = vx * c + vy * (-s), vx * s + vy * c
    b[:, :-1] *= np.float32(scale)
    if pc_range is None:
        return b, np.ones(b.shape[0], dtype=bool)
    m1 = (b[:, 0] >= pc_range[0]) & (b[:, 0] <= pc_range[3]) & (b[:, 1] >= pc_range[1]) & (b[:, 1] <= pc_range[4])
    half = b[:, 5] * np.float32(0.5)
    m2 = ((b[:, 2] + half) < pc_range[2]) ^ ((b[:, 2] - half) > pc_range[5])
    keep = m1 & ~m2
    return b[keep], keep
