"""Goldens for the GT-database paste from the REFERENCE's own code, imported from /root/reference through
tests/golden/ref_env.py: `box_collision_test` and `center_to_corner_box2d` (efg/geometry/box_ops.py:27-95,561-577),
`DataBaseSampler.sample_class` / `sample_all` (efg/data/samplers/gt_database_sampler.py:111-212) over an in-memory
database, and `points_in_rbbox` (box_ops.py:98-112) used by `DatabaseSampling(rm_points_after_sample=True)`
(efg/data/augmentations/extend_3d.py:84-90).
Usage: python tests/golden/make_golden_gt_paste.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, HERE)
import ref_env  # noqa: E402


def random_boxes(rng, n, spread):
    b = np.zeros((n, 7), dtype=np.float32)
    b[:, 0:2] = rng.uniform(-spread, spread, (n, 2))
    b[:, 2] = rng.uniform(-1.0, 1.0, n)
    b[:, 3] = rng.uniform(1.5, 5.0, n)
    b[:, 4] = rng.uniform(0.8, 2.2, n)
    b[:, 5] = rng.uniform(1.2, 2.0, n)
    b[:, 6] = rng.uniform(-3.2, 3.2, n)
    return b


def main():
    ref_env.install()
    try:
        from efg.data.samplers.gt_database_sampler import DataBaseSampler
        from efg.geometry import box_ops

        for seed in (1, 2, 3):
            rng = np.random.default_rng(seed)
            n_gt = int(rng.integers(6, 14))
            spread = 14.0 if seed != 3 else 7.0           # seed 3: crowded scene, most candidates collide
            gt_boxes = random_boxes(rng, n_gt, spread)
            classes = ["VEHICLE", "PEDESTRIAN", "CYCLIST"]
            gt_names = np.array([classes[int(i)] for i in rng.integers(0, 3, n_gt)])
            # in-memory database: per class a list of infos in the reference's format
            db, db_points = {}, {}
            for ci, name in enumerate(classes):
                infos = []
                for k in range(12):
                    box = random_boxes(rng, 1, spread)[0]
                    npts = int(rng.integers(5, 40))
                    pts = np.concatenate([rng.normal(0, 0.5, (npts, 3)), rng.uniform(0, 1, (npts, 2))], 1).astype(np.float32)
                    key = "%s_%d" % (name, k)
                    db_points[key] = pts
                    infos.append({"name": name, "path": key, "box3d_lidar": box, "num_points_in_gt": npts, "difficulty": 0})
                db[name] = infos
            groups = [{"VEHICLE": 8}, {"PEDESTRIAN": 6}, {"CYCLIST": 5}]

            # collision matrix of the reference on gt + all database boxes of one class
            allb = np.concatenate([gt_boxes] + [np.stack([i["box3d_lidar"] for i in db["VEHICLE"]])], 0)
            bv = box_ops.center_to_corner_box2d(allb[:, 0:2], allb[:, 3:5], allb[:, -1])
            coll = box_ops.box_collision_test(bv, bv)

            # the sampler, built without its pickle loader: sample_func hands out the first `num` infos of the class
            sampler = DataBaseSampler.__new__(DataBaseSampler)
            sampler._sample_classes = [list(g.keys())[0] for g in groups]
            sampler._sample_max_nums = [list(g.values())[0] for g in groups]
            sampler.sample_func = lambda name, num: db[name][:num]
            accepted = {}
            avoid = gt_boxes
            order = []
            for name, max_num in zip(sampler._sample_classes, sampler._sample_max_nums):
                num = int(max_num - np.sum([n == name for n in gt_names]))
                if num > 0:
                    got = sampler.sample_class(name, num, avoid)
                    accepted[name] = [g["path"] for g in got]
                    order += [g["path"] for g in got]
                    if got:
                        avoid = np.concatenate([avoid, np.stack([g["box3d_lidar"] for g in got])], 0)
            # points in rotated boxes (rm_points_after_sample)
            scene = np.concatenate([rng.uniform(-spread - 2, spread + 2, (4000, 2)), rng.uniform(-2.5, 2.5, (4000, 1)),
                                    rng.uniform(0, 1, (4000, 2))], 1).astype(np.float32)
            # a share of the points exactly on box centres / near faces
            scene[:n_gt, :3] = gt_boxes[:, :3]
            pasted_boxes = avoid[n_gt:]
            masks = box_ops.points_in_rbbox(scene, pasted_boxes) if pasted_boxes.shape[0] else np.zeros((4000, 0), bool)
            flat = {}
            for name, infos in db.items():
                flat["db_boxes_" + name] = np.stack([i["box3d_lidar"] for i in infos])
                flat["db_counts_" + name] = np.array([i["num_points_in_gt"] for i in infos], np.int32)
                flat["db_points_" + name] = np.concatenate([db_points[i["path"]] for i in infos], 0)
            np.savez_compressed(os.path.join(HERE, "gt_paste_seed%d.npz" % seed), gt_boxes=gt_boxes, gt_names=gt_names.astype("U16"),
                                max_nums=np.array(sampler._sample_max_nums, np.int32), classes=np.array(sampler._sample_classes).astype("U16"),
                                coll_boxes=allb, coll_corners=bv.astype(np.float32), coll=coll, accepted=np.array(order).astype("U24"),
                                scene=scene, pasted_boxes=pasted_boxes, in_box=masks.any(-1), **flat)
            print(seed, "gt", n_gt, "accepted", {k: len(v) for k, v in accepted.items()}, "coll pairs", int(coll.sum()),
                  "points removed", int(masks.any(-1).sum()))
    finally:
        ref_env.uninstall()


if __name__ == "__main__":
    main()
