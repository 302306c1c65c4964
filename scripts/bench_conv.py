"""Micro-benchmark of the sparse-conv kernels on the real level geometry of a 2x150k-point batch."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from efg_b200 import ops
from efg_b200.data import WAYMO, make_batch

dev = torch.device("cuda:0")
scenes = make_batch(2, 150000, WAYMO, seed=1)
pts = torch.from_numpy(np.concatenate([s[0] for s in scenes], 0)).to(dev)
offs = torch.tensor([0, 150000, 300000], dtype=torch.int32, device=dev)
r = ops.hard_voxelize_batched(pts, offs, WAYMO.voxel_size, WAYMO.pc_range, 5, 150000, coors_dim=4, want_voxels=False)
m = int(r["counts"][-1].item())
coords, shape = r["coors"][:m].contiguous(), [41, 1504, 1504]
levels = []
for lvl in range(4):
    oc, od, nbr_s, nbr_t = ops.sparse_rulebook(coords, 2, shape, 3, 2, 1)
    nbr = ops.subm_rulebook(oc, 2, od, 3, rows_sorted=True)
    levels.append((oc, od, nbr, nbr_s, nbr_t, coords.shape[0]))
    coords, shape = oc, od

def timeit(fn, iters=int(os.environ.get("ITERS", "20"))):
    for _ in range(int(os.environ.get("WARM", "3"))): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3  # us

modes = sys.argv[1:] or ["fp32x3"]
print("level rows   C     kind        us     alg GB/s  issued TFLOP/s  pairs/row")
for mode in modes:
    ops.CONV_PRECISION = mode
    for lvl, c in zip(range(4), [16, 64, 128, 256]):
        oc, od, nbr, nbr_s, nbr_t, m_in = levels[lvl]
        mo = oc.shape[0]
        feats = torch.randn(mo, c, device=dev)
        w = torch.randn(c, 27, c, device=dev) * 0.05
        go = torch.randn(mo, c, device=dev)
        nbytes = 4 * (2 * mo * c + 27 * c * c + 27 * mo)
        flops = 2 * mo * 27 * c * c
        ppr = float((nbr >= 0).sum()) / mo
        from efg_b200 import _lib
        L = _lib.lib()
        split = 1 if mode == "fp32x3" else 0
        packed = torch.empty(L.efgb_spconv_tc_packed_bytes(27, c, c, split) // 4, dtype=torch.float32, device=dev)
        L.efgb_spconv_tc_pack(ops._p(w), c, 27, c, 0, split, ops._p(packed), ops._stream())
        out = torch.empty(mo, c, device=dev)
        dw = torch.empty(c, 27, c, device=dev)
        st = ops._stream()
        def fwd_only():
            L.efgb_spconv_tc_forward(ops._p(feats), mo, c, ops._p(packed), None, ops._p(nbr), mo, 27, c, split, ops._p(out), st)
        def wgrad_only():
            L.efgb_spconv_tc_wgrad(ops._p(feats), mo, c, ops._p(go), ops._p(nbr), mo, 27, c, split, ops._p(dw), st)
        for kind, fn in (("subm fwd", fwd_only), ("subm wgrad", wgrad_only)):
            us = timeit(fn)
            print("L%d %7d %4d  %-10s %8.1f  %8.1f  %8.1f  %5.1f  [%s]" % (lvl + 1, mo, c, kind, us, nbytes / us / 1e3, flops / us / 1e6, ppr, mode))
