"""Generate the golden fixtures under tests/golden/ from the REFERENCE itself.

Runs only in the build container (needs /root/reference).  What it pins:
  voxelize_*.npz   outputs of the reference's numba voxelizer (efg/geometry/point_cloud_ops.py,
                   imported by path, unmodified) and of its C++ twin (voxelization_cpu.cpp compiled
                   by oracle/build_ref.py); the two must agree before anything is written.
  box_attn_*.pt    outputs and autograd gradients of the reference's
                   ``ms_deform_attn_core_pytorch`` (efg/operators/ms_deform_attn.py:55-76), the
                   torch twin of the box-attention CUDA kernel, imported by path with ``efg._C`` stubbed.
Inputs are seeded and stored next to the outputs so the tests never need /root/reference.
Usage: python tests/golden/make_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
REF = "/root/reference"

WAYMO_RANGE = [-75.2, -75.2, -2.0, 75.2, 75.2, 4.0]
WAYMO_VOXEL = [0.1, 0.1, 0.15]


def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def voxel_cases():
    rng = np.random.default_rng(20261017)
    cases = {}
    # (a) config-1-like: 20k uniform points on the Waymo grid, some outside the range
    pts = np.concatenate([rng.uniform(-80, 80, (20000, 2)), rng.uniform(-3, 5, (20000, 1)),
                          rng.uniform(0, 1, (20000, 2))], 1).astype(np.float32)
    cases["uniform20k"] = dict(points=pts, voxel_size=WAYMO_VOXEL, coors_range=WAYMO_RANGE, max_points=5,
                               max_voxels=150000)
    # (b) dense clusters: many points per voxel (max_points saturates) and the max_voxels cut-off fires
    centers = rng.uniform(-20, 20, (40, 3)) * np.array([1, 1, 0.05])
    pts = (centers[rng.integers(0, 40, 6000)] + rng.normal(0, 0.25, (6000, 3))).astype(np.float32)
    pts = np.concatenate([pts, rng.uniform(0, 1, (6000, 2)).astype(np.float32)], 1)
    cases["clusters_cutoff"] = dict(points=pts, voxel_size=WAYMO_VOXEL, coors_range=WAYMO_RANGE, max_points=5,
                                    max_voxels=700)
    # (c) coarse grid, 10 points per voxel, 4 features (nuScenes-like max_points)
    pts = np.concatenate([rng.uniform(-10, 10, (5000, 3)), rng.uniform(0, 1, (5000, 1))], 1).astype(np.float32)
    cases["coarse_mp10"] = dict(points=pts, voxel_size=[1.0, 1.0, 2.0], coors_range=[-8, -8, -4, 8, 8, 4],
                                max_points=10, max_voxels=20000)
    # (d) boundary values: points exactly on range borders and voxel edges
    g = np.arange(-8, 8.5, 0.5, dtype=np.float32)
    xx, yy, zz = np.meshgrid(g, g, np.array([-4, -2, 0, 2, 3.999, 4], dtype=np.float32), indexing="ij")
    pts = np.stack([xx.ravel(), yy.ravel(), zz.ravel(), np.zeros(xx.size, np.float32)], 1).astype(np.float32)
    pts = pts[rng.permutation(pts.shape[0])]
    cases["borders"] = dict(points=pts, voxel_size=[1.0, 1.0, 2.0], coors_range=[-8, -8, -4, 8, 8, 4], max_points=3,
                            max_voxels=600)
    # (e) a real LiDAR frame shipped with the reference for visualisation, subsampled to 20k points
    real = os.path.join(REF, "datasets/visualization/waymo_vis/example_data/example_point_cloud.bin.npy")
    if os.path.exists(real):
        cloud = np.load(real).astype(np.float32)
        sel = rng.permutation(cloud.shape[0])[:20000]
        cases["real_frame20k"] = dict(points=np.ascontiguousarray(cloud[sel]), voxel_size=WAYMO_VOXEL,
                                      coors_range=WAYMO_RANGE, max_points=5, max_voxels=150000)
    # (f) empty cloud
    cases["empty"] = dict(points=np.zeros((0, 5), np.float32), voxel_size=WAYMO_VOXEL, coors_range=WAYMO_RANGE,
                          max_points=5, max_voxels=100)
    return cases


def gen_voxelize():
    pco = load_by_path("ref_point_cloud_ops", os.path.join(REF, "efg/geometry/point_cloud_ops.py"))
    from oracle import build_ref

    build_ref.build()
    ref_cpp = build_ref.load()
    for name, case in voxel_cases().items():
        pts = case["points"]
        vs = np.array(case["voxel_size"], dtype=np.float32)
        rg = np.array(case["coors_range"], dtype=np.float32)
        if pts.shape[0]:
            v, c, n = pco.points_to_voxel(pts, vs, rg, case["max_points"], True, case["max_voxels"])
        else:
            v = np.zeros((0, case["max_points"], pts.shape[1]), np.float32)
            c = np.zeros((0, 3), np.int32)
            n = np.zeros((0,), np.int32)
        # the reference's C++ twin must agree with its numba voxelizer
        tp = torch.from_numpy(pts)
        tv = torch.zeros((case["max_voxels"], case["max_points"], pts.shape[1]))
        tc = torch.zeros((case["max_voxels"], 3), dtype=torch.int32)
        tn = torch.zeros((case["max_voxels"],), dtype=torch.int32)
        m = ref_cpp.hard_voxelize(tp, tv, tc, tn, [float(x) for x in vs], [float(x) for x in rg], case["max_points"],
                                  case["max_voxels"], 3)
        assert m == v.shape[0], (name, m, v.shape)
        assert np.array_equal(tv[:m].numpy(), v) and np.array_equal(tc[:m].numpy(), c) and \
            np.array_equal(tn[:m].numpy(), n), name
        dyn = torch.zeros((pts.shape[0], 3), dtype=torch.int32)
        ref_cpp.dynamic_voxelize(tp, dyn, [float(x) for x in vs], [float(x) for x in rg], 3)
        np.savez_compressed(os.path.join(HERE, "voxelize_%s.npz" % name), points=pts, voxel_size=vs, coors_range=rg,
                            max_points=case["max_points"], max_voxels=case["max_voxels"], voxels=v, coors=c,
                            num_points_per_voxel=n, dynamic_coors=dyn.numpy())
        print("voxelize", name, "N=%d M=%d" % (pts.shape[0], v.shape[0]))


def gen_box_attn():
    stub = types.ModuleType("efg")
    stub._C = types.ModuleType("efg._C")
    sys.modules.setdefault("efg", stub)
    sys.modules.setdefault("efg._C", stub._C)
    msda = load_by_path("ref_ms_deform_attn", os.path.join(REF, "efg/operators/ms_deform_attn.py"))
    g = torch.Generator().manual_seed(1234)
    cases = {
        # Voxel-DETR-like: 1 level, 8 heads x 32 channels, 25 points; locations spill over the borders
        "single_level": dict(B=2, H=8, C=32, shapes=[(12, 14)], LQ=37, P=25),
        # multi-level, odd head_dim
        "multi_level": dict(B=1, H=4, C=16, shapes=[(9, 7), (5, 4), (3, 2)], LQ=21, P=4),
        "wide_head": dict(B=1, H=2, C=64, shapes=[(6, 6)], LQ=9, P=9),
    }
    for name, cs in cases.items():
        shapes = torch.tensor(cs["shapes"], dtype=torch.int64)
        lv = int((shapes[:, 0] * shapes[:, 1]).sum())
        L = shapes.shape[0]
        value = torch.randn(cs["B"], lv, cs["H"], cs["C"], generator=g)
        loc = torch.rand(cs["B"], cs["LQ"], cs["H"], L, cs["P"], 2, generator=g) * 1.3 - 0.15
        attn = torch.softmax(torch.randn(cs["B"], cs["LQ"], cs["H"], L * cs["P"], generator=g), -1)
        attn = attn.view(cs["B"], cs["LQ"], cs["H"], L, cs["P"])
        grad_out = torch.randn(cs["B"], cs["LQ"], cs["H"] * cs["C"], generator=g)
        v, l, a = value.clone().requires_grad_(), loc.clone().requires_grad_(), attn.clone().requires_grad_()
        out = msda.ms_deform_attn_core_pytorch(v, shapes, l, a)
        out.backward(grad_out)
        torch.save(dict(value=value, shapes=shapes, loc=loc, attn=attn, grad_out=grad_out, out=out.detach(),
                        grad_value=v.grad, grad_loc=l.grad, grad_attn=a.grad),
                   os.path.join(HERE, "box_attn_%s.pt" % name))
        print("box_attn", name, tuple(out.shape))


if __name__ == "__main__":
    gen_voxelize()
    gen_box_attn()
