"""Goldens for the point-cloud augmentation chain from the REFERENCE's own processor classes
(efg/data/augmentations/extend_3d.py: RandomFlip3D, GlobalRotation, GlobalScaling, FilterByRange), imported from
/root/reference through tests/golden/ref_env.py and run under a seeded np.random.  The random draws each case made are
recovered by replaying the same numpy calls under the same seed and stored with the outputs.
Usage: python tests/golden/make_golden_augment.py"""
import copy
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, HERE)
import ref_env  # noqa: E402

RANGE = [-75.2, -75.2, -2.0, 75.2, 75.2, 4.0]


def main():
    from efg_b200.data import WAYMO, make_scene

    ref_env.install()
    try:
        from efg.data.augmentations import extend_3d as E

        for seed in (1, 2, 3, 6):
            pts, ann = make_scene(5000, WAYMO, seed=seed, num_objects=40)
            rng = np.random.default_rng(seed)
            ann["gt_boxes"][:, 6:8] = rng.normal(0, 3, (ann["gt_boxes"].shape[0], 2)).astype(np.float32)
            ann["gt_boxes"][::3, :2] *= np.float32(1.6)     # some boxes leave the range
            ann["gt_boxes"][5::7, 2] += np.float32(9.0)     # ... or float above it
            pts[::4, :2] *= np.float32(1.25)                # a share of the points leaves the range after scaling
            info = {"annotations": copy.deepcopy(ann)}
            procs = [E.RandomFlip3D(p=0.5), E.GlobalRotation(rotation=0.78539816), E.GlobalScaling(min_scale=0.8, max_scale=1.2),
                     E.FilterByRange(pc_range=RANGE)]
            np.random.seed(100 + seed)
            p = pts.copy()
            for proc in procs:
                p, info = proc(p, info)
            # the draws, replayed: two flip choices, one angle, one scale (extend_3d.py:129,148,194,212)
            np.random.seed(100 + seed)
            flip_x = bool(np.random.choice([False, True], replace=False, p=[0.5, 0.5]))
            flip_y = bool(np.random.choice([False, True], replace=False, p=[0.5, 0.5]))
            angle = float(np.random.uniform(-0.78539816, 0.78539816))
            scale = float(np.random.uniform(0.8, 1.2))
            np.savez_compressed(os.path.join(HERE, "augment_seed%d.npz" % seed), points=pts, gt_boxes=ann["gt_boxes"],
                                flip_x=flip_x, flip_y=flip_y, angle=angle, scale=scale, pc_range=np.asarray(RANGE, np.float32),
                                out_points=p, out_boxes=info["annotations"]["gt_boxes"],
                                out_names=info["annotations"]["gt_names"].astype("U16"), np_seed=100 + seed)
            print(seed, flip_x, flip_y, round(angle, 4), round(scale, 4), pts.shape, "->", p.shape, ann["gt_boxes"].shape, "->",
                  info["annotations"]["gt_boxes"].shape)
    finally:
        ref_env.uninstall()


if __name__ == "__main__":
    main()
