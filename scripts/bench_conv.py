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

modes = sys.argv[1:] or ["bf16x3"]
print("level rows   C     kind        us     alg GB/s  issued TFLOP/s  pairs/row")
for mode in modes:
    ops.CONV_PRECISION = mode.split("-")[0]
    inline = mode.endswith("-inline")
    for lvl, c in zip(range(4), [16, 64, 128, 256]):
        if str(lvl) not in os.environ.get("LEVELS", "0,1,2,3").split(",") or __name__ != "__main__":
            continue
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
        split = {"fp32x3": 1, "tf32": 0, "bf16x3": 2}[ops.CONV_PRECISION]
        packed = torch.empty(L.efgb_spconv_tc_packed_bytes(27, c, c, split) // 4, dtype=torch.float32, device=dev)
        L.efgb_spconv_tc_pack(ops._p(w), c, 27, c, 0, split, ops._p(packed), ops._stream())
        out = torch.empty(mo, c, device=dev)
        dw = torch.empty(c, 27, c, device=dev)
        planes = torch.empty_like(feats)
        st = ops._stream()
        use_planes = split == 2 and not inline and L.efgb_spconv_tc_planes_supported(c, c, 27)
        def split_only():
            L.efgb_split_bf16(ops._p(feats), mo, c, ops._p(planes), st)
        def fwd_only():
            if use_planes:
                L.efgb_spconv_tc_forward_planes(ops._p(planes), mo, c, ops._p(packed), None, ops._p(nbr), mo, 27, c, 0, ops._p(out), st)
            else:
                L.efgb_spconv_tc_forward(ops._p(feats), mo, c, ops._p(packed), None, ops._p(nbr), mo, 27, c, split, ops._p(out), st)
        def fwd_total():
            split_only(); fwd_only()
        def wgrad_only():
            L.efgb_spconv_tc_wgrad(ops._p(feats), mo, c, ops._p(go), ops._p(nbr), mo, 27, c, min(split, 1), ops._p(dw), st)
        kinds = [("subm fwd", fwd_only), ("subm wgrad", wgrad_only)]
        if use_planes:
            split_only()
            kinds = [("subm split", split_only), ("subm fwd", fwd_only), ("subm split+fwd", fwd_total), ("subm wgrad", wgrad_only)]
        for kind, fn in kinds:
            if kind.split()[1].split("+")[-1] not in os.environ.get("KINDS", "fwd,wgrad,split").split(","):
                continue
            us = timeit(fn)
            print("L%d %7d %4d  %-14s %8.1f  %8.1f  %8.1f  %5.1f  [%s]" % (lvl + 1, mo, c, kind, us, nbytes / us / 1e3, flops / us / 1e6, ppr, mode))
