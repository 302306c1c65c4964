"""ORACLE (test infrastructure) — dynamic point-to-voxel scatter on the CPU (numpy).

Restates efg/operators/src/voxelize/scatter_points_cuda.cu:209-290 (forward) and :292-352
(backward): linear id over dims = coors.max(0)+1 (:220, :70-82, rows with a negative coordinate
get id -1 and are dropped), voxels in ascending id order (argsort + cumsum, :236-250), reduce
sum / mean / max; count is only accumulated for 'mean' (:122-124).
Parity unpinned by reference fixtures (the reference has no CPU build of this op bound,
efg/operators/src/voxelize/voxelization.h:106) — it is cross-checked against numpy group-by
identities in tests/test_oracle.py.
"""
import numpy as np


def forward(feats, coors, reduce_type):
    feats = np.asarray(feats, dtype=np.float32)
    coors = np.asarray(coors, dtype=np.int32)
    n, c = feats.shape
    if n == 0:
        return (np.zeros((0, c), np.float32), np.zeros((0, 3), np.int32), np.zeros((0,), np.int32),
                np.zeros((0,), np.int32))
    dims = coors.max(0).astype(np.int64) + 1
    valid = (coors >= 0).all(1)
    lin = (coors[:, 0].astype(np.int64) * dims[1] + coors[:, 1]) * dims[2] + coors[:, 2]
    lin = np.where(valid, lin, -1)
    uniq = np.unique(lin[valid])
    m = uniq.shape[0]
    p2v = np.full((n,), -1, dtype=np.int32)
    p2v[valid] = np.searchsorted(uniq, lin[valid]).astype(np.int32)
    out_coors = np.zeros((m, 3), dtype=np.int32)
    out_coors[:, 2] = uniq % dims[2]
    out_coors[:, 1] = (uniq // dims[2]) % dims[1]
    out_coors[:, 0] = uniq // (dims[2] * dims[1])
    count = np.zeros((m,), dtype=np.int32)
    if reduce_type == "max":
        out = np.full((m, c), -np.inf, dtype=np.float32)
        np.maximum.at(out, p2v[valid], feats[valid])
    else:
        out = np.zeros((m, c), dtype=np.float32)
        np.add.at(out, p2v[valid], feats[valid])
        if reduce_type == "mean":
            np.add.at(count, p2v[valid], 1)
            out = out / count.astype(np.float32)[:, None]
    return out, out_coors, p2v, count


def backward(grad_voxel, feats, voxel_feats, p2v, count, reduce_type):
    feats = np.asarray(feats, dtype=np.float32)
    n, c = feats.shape
    grad = np.zeros((n, c), dtype=np.float32)
    valid = p2v >= 0
    if reduce_type in ("sum", "mean"):
        g = grad_voxel[p2v[valid]]
        if reduce_type == "mean":
            g = g / count[p2v[valid]].astype(np.float32)[:, None]
        grad[valid] = g
        return grad
    # max: gradient goes to the arg-max point with the smallest index (:165-205)
    m = voxel_feats.shape[0]
    reduce_from = np.full((m, c), n, dtype=np.int64)
    for i in np.nonzero(valid)[0]:
        v = p2v[i]
        hit = feats[i] == voxel_feats[v]
        reduce_from[v, hit] = np.minimum(reduce_from[v, hit], i)
    for v in range(m):
        for ch in range(c):
            if reduce_from[v, ch] < n:
                grad[reduce_from[v, ch], ch] = grad_voxel[v, ch]
    return grad
