"""CPU: pin the oracle against the golden vectors generated from the reference itself
(tests/golden/make_golden.py) and cross-check its independent restatements."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import box_attn as obox
from oracle import scatter as oscatter
from oracle import sparse_conv as sc
from oracle import voxelize as ovox

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
VOX_CASES = sorted(glob.glob(os.path.join(GOLDEN, "voxelize_*.npz")))
BOX_CASES = sorted(glob.glob(os.path.join(GOLDEN, "box_attn_*.pt")))


def test_golden_fixtures_present():
    assert len(VOX_CASES) >= 5 and len(BOX_CASES) >= 3


@pytest.mark.parametrize("path", VOX_CASES, ids=[os.path.basename(p)[9:-4] for p in VOX_CASES])
def test_voxelizer_oracle_matches_reference(path):
    g = np.load(path)
    v, c, n = ovox.hard_voxelize(g["points"], g["voxel_size"], g["coors_range"], int(g["max_points"]),
                                 int(g["max_voxels"]))
    assert v.shape == g["voxels"].shape
    assert np.array_equal(c, g["coors"])
    assert np.array_equal(n, g["num_points_per_voxel"])
    assert np.array_equal(v, g["voxels"])  # bit exact
    d = ovox.dynamic_voxelize(g["points"], g["voxel_size"], g["coors_range"])
    assert np.array_equal(d, g["dynamic_coors"])


def test_voxelizer_c_matches_python_loop():
    rng = np.random.default_rng(3)
    pts = rng.uniform(-3, 3, (3000, 4)).astype(np.float32)
    for mp, mv in ((3, 50), (1, 10 ** 6), (8, 200)):
        a = ovox.hard_voxelize(pts, [0.4, 0.5, 0.6], [-2, -2, -2, 2, 2, 2], mp, mv)
        b = ovox.hard_voxelize_py(pts, [0.4, 0.5, 0.6], [-2, -2, -2, 2, 2, 2], mp, mv)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


def test_reference_cpp_voxelizer_if_built():
    """oracle/_ref (the reference's own C++ file compiled here) agrees with the restatement."""
    from oracle import build_ref

    ref = build_ref.load()
    if ref is None:
        pytest.skip("oracle/_ref not built on this box")
    rng = np.random.default_rng(5)
    pts = rng.uniform(-80, 80, (30000, 5)).astype(np.float32)
    pts[:, 2] = rng.uniform(-3, 5, 30000)
    vs, rg = [0.1, 0.1, 0.15], [-75.2, -75.2, -2.0, 75.2, 75.2, 4.0]
    tv = torch.zeros((15000, 5, 5))
    tc = torch.zeros((15000, 3), dtype=torch.int32)
    tn = torch.zeros((15000,), dtype=torch.int32)
    m = ref.hard_voxelize(torch.from_numpy(pts), tv, tc, tn, vs, rg, 5, 15000, 3)
    v, c, n = ovox.hard_voxelize(pts, vs, rg, 5, 15000)
    assert m == v.shape[0] == 15000  # cut-off fired
    assert np.array_equal(tv.numpy()[:m], v) and np.array_equal(tc.numpy()[:m], c) and np.array_equal(tn.numpy()[:m], n)


@pytest.mark.parametrize("path", BOX_CASES, ids=[os.path.basename(p)[9:-3] for p in BOX_CASES])
def test_box_attn_oracle_matches_reference(path):
    g = torch.load(path)
    out, gv, gl, ga = obox.forward_backward(g["value"], g["shapes"], g["loc"], g["attn"], g["grad_out"])
    assert torch.equal(out, g["out"])
    assert torch.allclose(gv, g["grad_value"], atol=1e-6) and torch.allclose(gl, g["grad_loc"], atol=1e-5)
    assert torch.allclose(ga, g["grad_attn"], atol=1e-6)


def test_box_attn_scalar_loops_match_grid_sample():
    g = torch.load(os.path.join(GOLDEN, "box_attn_multi_level.pt"))
    shapes = g["shapes"]
    start = torch.cat([shapes.new_zeros(1), (shapes[:, 0] * shapes[:, 1]).cumsum(0)[:-1]])
    out = obox.forward_loops(g["value"], shapes, start, g["loc"], g["attn"])
    assert torch.allclose(out, g["out"], atol=2e-6)


def _random_sites(rng, batch, dhw, m):
    cells = rng.choice(batch * dhw[0] * dhw[1] * dhw[2], size=m, replace=False)
    d, h, w = dhw
    return np.stack([cells // (d * h * w), (cells // (h * w)) % d, (cells // w) % h, cells % w], 1).astype(np.int32)


@pytest.mark.parametrize("geom", [(3, 2, 1), ((3, 1, 1), (2, 1, 1), (1, 0, 0)), (3, 2, (0, 1, 1)),
                                  ((3, 1, 1), (2, 1, 1), 0), (3, 1, 1)])
def test_sparse_conv_oracle_sparse_vs_dense(geom):
    k, s, p = geom
    rng = np.random.default_rng(11)
    batch, dhw = 2, [7, 13, 12]
    coords = _random_sites(rng, batch, dhw, 260)
    feats = torch.randn(260, 6)
    kk = sc._triple(k)
    w = torch.randn(9, *kk, 6) * 0.2
    b = torch.randn(9)
    oc, od, nbr, nbr_t = sc.sparse_rulebook(coords, batch, dhw, k, s, p)
    y = sc.conv(feats, w, b, nbr)
    yd = sc.dense_conv_at_sites(feats, coords, batch, dhw, w, b, s, p, oc)
    assert torch.allclose(y, yd, atol=1e-5)
    # every output site has at least one contributing input; nbr_t is the transpose of nbr
    assert (nbr >= 0).any(1).all()
    o, t = np.nonzero(nbr >= 0)
    assert np.array_equal(nbr_t[nbr[o, t], t], o)
    # ascending linear order
    keys = sc.linear_key(oc, od)
    assert np.all(np.diff(keys) > 0)


def test_subm_oracle_sparse_vs_dense_and_symmetry():
    rng = np.random.default_rng(12)
    batch, dhw = 2, [5, 9, 11]
    coords = _random_sites(rng, batch, dhw, 300)
    feats = torch.randn(300, 4)
    w = torch.randn(8, 3, 3, 3, 4) * 0.2
    nbr = sc.subm_rulebook(coords, batch, dhw, 3)
    y = sc.conv(feats, w, None, nbr)
    yd = sc.dense_conv_at_sites(feats, coords, batch, dhw, w, None, 1, 1, coords)
    assert torch.allclose(y, yd, atol=1e-5)
    assert np.array_equal(nbr[:, 13], np.arange(300))  # centre tap is the site itself
    i, t = np.nonzero(nbr >= 0)
    assert np.array_equal(nbr[nbr[i, t], 26 - t], i)  # tap symmetry used by the CUDA dgrad


def test_scatter_oracle_groupby():
    rng = np.random.default_rng(13)
    coors = rng.integers(-1, 6, (500, 3)).astype(np.int32)
    feats = rng.normal(size=(500, 4)).astype(np.float32)
    for red in ("sum", "mean", "max"):
        out, oc, p2v, count = oscatter.forward(feats, coors, red)
        valid = (coors >= 0).all(1)
        assert (p2v[~valid] == -1).all() and (p2v[valid] >= 0).all()
        assert np.array_equal(oc[p2v[valid]], coors[valid])
        for v in range(0, out.shape[0], 17):
            grp = feats[p2v == v]
            ref = {"sum": grp.sum(0), "mean": grp.mean(0), "max": grp.max(0)}[red]
            assert np.allclose(out[v], ref, atol=1e-5)
        g = rng.normal(size=out.shape).astype(np.float32)
        gf = oscatter.backward(g, feats, out, p2v, count, red)
        assert gf.shape == feats.shape and (gf[~valid] == 0).all()


def test_synthetic_scene_statistics():
    """The generator is accepted only if its voxel statistics are within +-20% of the real frame
    calibration: M/N ~ 0.52, SubM pairs/site ~ 6.7 at level 0."""
    from efg_b200.data import WAYMO, make_scene

    pts, ann = make_scene(150000, WAYMO, seed=1)
    assert pts.shape == (150000, 5) and pts.dtype == np.float32
    v, c, n = ovox.hard_voxelize(pts, WAYMO.voxel_size, WAYMO.pc_range, 5, 150000)
    ratio = v.shape[0] / pts.shape[0]
    assert 0.41 < ratio < 0.63, ratio  # 0.52 +- 20%
    coords = np.pad(c, ((0, 0), (1, 0)))
    nbr = sc.subm_rulebook(coords, 1, [41, 1504, 1504], 3)
    pairs = (nbr >= 0).sum() / len(coords)
    assert 5.3 < pairs < 8.1, pairs  # 6.7 +- 20%
    assert ann["gt_boxes"].shape[1] == 9 and ann["labels"].min() >= 1


# --------------------------------------------------------------------------------------------------
# linear sum assignment: the restatement (and the lane-parallel reduction order the CUDA kernel uses) vs scipy
# --------------------------------------------------------------------------------------------------
def _lsa_cases():
    rng = np.random.default_rng(0)
    for trial in range(90):
        nr, nc = int(rng.integers(1, 45)), int(rng.integers(1, 45))
        kind = trial % 3
        if kind == 0:
            yield rng.random((nr, nc)).astype(np.float32)
        elif kind == 1:
            yield rng.integers(0, 4, (nr, nc)).astype(np.float32)  # heavy ties
        else:
            yield np.round(rng.random((nr, nc)) * 3).astype(np.float32)
    yield np.zeros((9, 5), np.float32)
    yield np.ones((4, 11), np.float32)


def test_lsa_restatement_equals_scipy_including_ties():
    from scipy.optimize import linear_sum_assignment

    from oracle import lsa

    for c in _lsa_cases():
        i0, j0 = linear_sum_assignment(c)
        i1, j1 = lsa.lsa_sequential(c)
        assert np.array_equal(i0, i1) and np.array_equal(j0, j1), c.shape


def test_lsa_lane_parallel_reduction_order_equals_scipy():
    from scipy.optimize import linear_sum_assignment

    from oracle import lsa

    for c in _lsa_cases():
        i0, j0 = linear_sum_assignment(c)
        for lanes in (32, 4):
            i1, j1 = lsa.lsa_lane_parallel(c, lanes=lanes)
            assert np.array_equal(i0, i1) and np.array_equal(j0, j1), (c.shape, lanes)
