"""Goldens for BEV IoU / rotated NMS from the REFERENCE's own iou3d_cpu.cpp (compiled from /root/reference into
oracle/_ref by oracle/build_ref.py:build_iou3d).  The reference's NMS itself is CUDA-only (iou3d_nms.cpp:60-121); its
result is determined by its IoU values: kept = greedy scan over score-sorted boxes with `iou > thresh` (the host loop of
nms_gpu), evaluated here on the reference's IoU matrix.
Usage: python tests/golden/make_golden_iou3d.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))


def cases():
    rng = np.random.default_rng(20261017)
    out = {}

    def boxes(n, spread, size=(4.5, 2.0, 1.6), jitter=0.3):
        c = rng.uniform(-spread, spread, (n, 2))
        z = rng.uniform(-1, 1, (n, 1))
        d = np.asarray(size)[None] * rng.uniform(1 - jitter, 1 + jitter, (n, 3))
        yaw = rng.uniform(-np.pi, np.pi, (n, 1))
        return np.concatenate([c, z, d, yaw], 1).astype(np.float32)

    out["random_dense"] = (boxes(96, 12.0), boxes(80, 12.0))                       # many partial overlaps
    a = boxes(64, 30.0)
    dup = a[rng.integers(0, 64, 64)] + rng.normal(0, [0.15, 0.15, 0.05, 0.1, 0.05, 0.05, 0.05], (64, 7)).astype(np.float32)
    out["near_duplicates"] = (np.concatenate([a, dup]).astype(np.float32), np.concatenate([a, dup]).astype(np.float32))  # NMS-like
    ax = boxes(40, 8.0)
    ax[:, 6] = rng.choice([0.0, np.pi / 2, np.pi, -np.pi / 2], 40)                 # axis-aligned: degenerate parallel edges
    out["axis_aligned"] = (ax, ax.copy())
    t = np.array([[0, 0, 0, 4, 2, 1.5, 0.0], [4.0, 0, 0, 4, 2, 1.5, 0.0], [4.005, 0, 0, 4, 2, 1.5, 0.0], [0, 2.0, 0, 4, 2, 1.5, 0.0],
                  [0, 0, 0, 4, 2, 1.5, np.pi / 4], [0, 0, 0, 0.5, 0.5, 1.5, 0.3], [100, 100, 0, 4, 2, 1.5, 1.0]], np.float32)
    out["touching_contained_far"] = (t, t.copy())                                   # the 1 cm margin, containment, disjoint
    out["pedestrians"] = (boxes(128, 6.0, size=(0.9, 0.9, 1.7)), boxes(128, 6.0, size=(0.9, 0.9, 1.7)))
    return out


def main():
    from oracle import build_ref

    build_ref.build_iou3d()
    ref = build_ref.load_iou3d()
    assert ref is not None
    for name, (a, b) in cases().items():
        iou = torch.zeros(a.shape[0], b.shape[0])
        ref.boxes_iou_bev_cpu(torch.from_numpy(a).contiguous(), torch.from_numpy(b).contiguous(), iou)
        iou = iou.numpy()
        # NMS golden on set a: descending pseudo-scores = reverse index order (already "sorted"), three thresholds
        self_iou = torch.zeros(a.shape[0], a.shape[0])
        ref.boxes_iou_bev_cpu(torch.from_numpy(a).contiguous(), torch.from_numpy(a).contiguous(), self_iou)
        self_iou = self_iou.numpy()
        keeps, margins = {}, {}
        for thr in (0.1, 0.5, 0.7):
            keep = []
            for i in range(a.shape[0]):
                if not any(self_iou[j, i] > thr for j in keep):
                    keep.append(i)
            keeps["keep_%g" % thr] = np.asarray(keep, dtype=np.int64)
            # smallest distance of a decisive IoU value to the threshold (a test may not rely on ties within rounding)
            margins["margin_%g" % thr] = float(np.min(np.abs(self_iou[np.triu_indices(a.shape[0], 1)] - thr)))
        np.savez_compressed(os.path.join(HERE, "iou3d_%s.npz" % name), boxes_a=a, boxes_b=b, iou=iou, self_iou=self_iou,
                            **keeps, **{k: np.float32(v) for k, v in margins.items()})
        print(name, a.shape, b.shape, "iou>0: %d" % int((iou > 0).sum()), {k: len(v) for k, v in keeps.items()},
              {k: "%.1e" % v for k, v in margins.items()})


if __name__ == "__main__":
    main()
