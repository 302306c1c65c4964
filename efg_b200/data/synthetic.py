"""Synthetic "LiDAR-ring" scenes shaped like the reference's Waymo / nuScenes inputs.

There is no dataset on the box, so every measurement and parity test runs on this generator
(SURVEY.md §8d): a 64-beam spinning sensor at the origin ray-cast against a ground plane, ~60
object boxes (also returned as ground truth) and a few wall segments, with range noise, then
resampled to an exact point count and shuffled (the reference shuffles too, PointShuffle,
efg/data/augmentations/extend_3d.py:109).  Uniform noise would give ~1 point per voxel and
almost no active neighbours, i.e. unrepresentative rulebooks.

The per-sample dict mirrors what the reference's Voxelization processor emits
(extend_3d.py:273-282) and ``info["annotations"]`` mirrors waymo.py:135-140.
"""
from dataclasses import dataclass, field

import numpy as np


@dataclass
class SceneSpec:
    pc_range: list
    voxel_size: list
    num_point_features: int = 5
    max_points_in_voxel: int = 5
    max_voxel_num: int = 150000
    classes: list = field(default_factory=lambda: ["VEHICLE", "PEDESTRIAN", "CYCLIST"])
    sensor_height: float = 1.7
    num_beams: int = 64
    elevation_deg: tuple = (-17.6, 2.4)
    nsweeps: int = 1

    @property
    def grid_size(self):
        r = np.asarray(self.pc_range, dtype=np.float32)
        v = np.asarray(self.voxel_size, dtype=np.float32)
        return np.round((r[3:] - r[:3]) / v).astype(np.int64)  # (x, y, z)


# VoxelDETR / CenterPoint Waymo configs: pc_range / voxel_size from
# playground/detection.3d/waymo/conquer/VoxelDETR.../config.yaml:19-20
WAYMO = SceneSpec(pc_range=[-75.2, -75.2, -2.0, 75.2, 75.2, 4.0], voxel_size=[0.1, 0.1, 0.15])
# playground/detection.3d/nuscenes/centerpoint/.../config.yaml:14-15
NUSCENES = SceneSpec(pc_range=[-54.0, -54.0, -5.0, 54.0, 54.0, 3.0], voxel_size=[0.075, 0.075, 0.2],
                     max_points_in_voxel=10, max_voxel_num=160000, nsweeps=11, elevation_deg=(-30.0, 10.0),
                     num_beams=32,
                     classes=["car", "truck", "construction_vehicle", "bus", "trailer", "barrier", "motorcycle",
                              "bicycle", "pedestrian", "traffic_cone"])

_CLASS_DIMS = {0: (4.7, 2.1, 1.7), 1: (0.9, 0.9, 1.7), 2: (1.8, 0.8, 1.7)}  # l, w, h


def _ray_boxes(origin, dirs, centers, dims, yaws):
    """Nearest hit distance of rays (origin + t*dirs) against rotated boxes; inf when none. Slab test."""
    t_best = np.full(dirs.shape[0], np.inf, dtype=np.float32)
    for c, d, yaw in zip(centers, dims, yaws):
        cs, sn = np.cos(-yaw), np.sin(-yaw)
        o = origin - c
        ox = cs * o[0] - sn * o[1]
        oy = sn * o[0] + cs * o[1]
        oz = o[2]
        dx = cs * dirs[:, 0] - sn * dirs[:, 1]
        dy = sn * dirs[:, 0] + cs * dirs[:, 1]
        dz = dirs[:, 2]
        tmin = np.full(dirs.shape[0], -np.inf, dtype=np.float32)
        tmax = np.full(dirs.shape[0], np.inf, dtype=np.float32)
        for oo, dd, half in ((ox, dx, d[0] / 2), (oy, dy, d[1] / 2), (oz, dz, d[2] / 2)):
            with np.errstate(divide="ignore", invalid="ignore"):
                inv = 1.0 / dd
                t1 = (-half - oo) * inv
                t2 = (half - oo) * inv
            lo, hi = np.minimum(t1, t2), np.maximum(t1, t2)
            tmin = np.maximum(tmin, lo)
            tmax = np.minimum(tmax, hi)
        hit = (tmax >= tmin) & (tmax > 0)
        t = np.where(tmin > 0, tmin, tmax)
        t_best = np.where(hit & (t < t_best), t, t_best)
    return t_best


def make_scene(num_points, spec=WAYMO, seed=0, num_objects=60):
    """-> (points [N, F] f32, annotations dict).  Deterministic in (num_points, spec, seed)."""
    rng = np.random.default_rng(1234 + seed)
    r = np.asarray(spec.pc_range, dtype=np.float64)
    ground_z = -spec.sensor_height
    ncls = min(len(spec.classes), 3)

    # objects: 2/3 vehicles, 1/4 pedestrians, rest cyclists, on the ground
    n_obj = num_objects
    cls = rng.choice(ncls, size=n_obj, p=np.array([0.67, 0.25, 0.08])[:ncls] / np.sum([0.67, 0.25, 0.08][:ncls]))
    dims = np.array([_CLASS_DIMS[int(c)] for c in cls]) * rng.uniform(0.85, 1.15, (n_obj, 3))
    rad = np.sqrt(rng.uniform(0.02, 1.0, n_obj)) * min(r[3], r[4]) * 0.92
    ang = rng.uniform(0, 2 * np.pi, n_obj)
    centers = np.stack([rad * np.cos(ang), rad * np.sin(ang), ground_z + dims[:, 2] / 2], 1)
    yaws = rng.uniform(-np.pi, np.pi, n_obj)
    # walls: long thin boxes, not part of the ground truth
    wall_c = np.array([[30.0, 0, ground_z + 1.5], [-35.0, 5, ground_z + 1.5], [0, 40.0, ground_z + 1.5],
                       [5, -45.0, ground_z + 1.5]])
    wall_d = np.array([[0.4, 50.0, 3.0], [0.4, 60.0, 3.0], [60.0, 0.4, 3.0], [50.0, 0.4, 3.0]])
    wall_y = rng.uniform(-0.3, 0.3, 4)

    pts_all = []
    per_sweep = int(np.ceil(num_points / spec.nsweeps))
    for sweep in range(spec.nsweeps):
        origin = np.array([1.0 * sweep, 0.0, 0.0])  # ego shift 1 m per past sweep
        n_az = int(np.ceil(per_sweep * 1.35 / spec.num_beams))
        az = (np.arange(n_az) + rng.uniform(0, 1)) * (2 * np.pi / n_az)
        el = np.deg2rad(np.linspace(spec.elevation_deg[0], spec.elevation_deg[1], spec.num_beams))
        azg, elg = np.meshgrid(az, el, indexing="ij")
        dirs = np.stack([np.cos(elg) * np.cos(azg), np.cos(elg) * np.sin(azg), np.sin(elg)], -1).reshape(-1, 3)
        with np.errstate(divide="ignore", invalid="ignore"):
            t_ground = np.where(dirs[:, 2] < 0, (ground_z - origin[2]) / dirs[:, 2], np.inf)
        t = np.minimum(t_ground, _ray_boxes(origin, dirs, np.concatenate([centers, wall_c]),
                                            np.concatenate([dims, wall_d]), np.concatenate([yaws, wall_y])))
        ok = np.isfinite(t)
        t = t[ok] + rng.normal(0, 0.02, ok.sum())
        p = origin[None] + dirs[ok] * t[:, None]
        inside = ((p[:, 0] > r[0]) & (p[:, 0] < r[3]) & (p[:, 1] > r[1]) & (p[:, 1] < r[4]) & (p[:, 2] > r[2]) &
                  (p[:, 2] < r[5]))
        p = p[inside]
        feats = [p, rng.uniform(0, 1, (p.shape[0], 1))]
        if spec.nsweeps == 1:
            feats.append(rng.uniform(0, 1, (p.shape[0], 1)))  # elongation
        else:
            feats.append(np.full((p.shape[0], 1), 0.05 * sweep))  # delta t
        extra = spec.num_point_features - 5
        if extra > 0:
            feats.append(rng.uniform(0, 1, (p.shape[0], extra)))
        pts_all.append(np.concatenate(feats, 1))
    pts = np.concatenate(pts_all, 0)
    if pts.shape[0] >= num_points:
        pts = pts[rng.permutation(pts.shape[0])[:num_points]]
    else:  # pad by re-observing random points with a little jitter
        extra = pts[rng.integers(0, pts.shape[0], num_points - pts.shape[0])].copy()
        extra[:, :3] += rng.normal(0, 0.03, (extra.shape[0], 3))
        pts = np.concatenate([pts, extra], 0)
        pts = pts[rng.permutation(pts.shape[0])]
    pts = np.ascontiguousarray(pts.astype(np.float32))

    gt_boxes = np.zeros((n_obj, 9), dtype=np.float32)  # x,y,z,l,w,h,vx,vy,heading (waymo_decoder.py:190-203)
    gt_boxes[:, :3] = centers
    gt_boxes[:, 3:6] = dims
    gt_boxes[:, 8] = yaws
    names = np.array([spec.classes[int(c)] for c in cls])
    annotations = {
        "gt_boxes": gt_boxes,
        "gt_names": names,
        "difficulty": np.zeros((n_obj,), dtype=np.int64),
        "num_points_in_gt": np.full((n_obj,), 50, dtype=np.int64),
        "labels": (cls + 1).astype(np.int64),  # 1-based (waymo.py:135-140)
    }
    return pts, annotations


def make_batch(batch_size, num_points, spec=WAYMO, seed=0, num_objects=60):
    """List of (points, annotations) for `batch_size` scenes with seeds seed*1000 + i."""
    return [make_scene(num_points, spec, seed * 1000 + i, num_objects) for i in range(batch_size)]
