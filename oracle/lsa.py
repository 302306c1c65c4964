"""ORACLE (test infrastructure) — linear sum assignment, restated.

The reference's matcher calls ``scipy.optimize.linear_sum_assignment`` (VD/modules/matcher.py:1,86-89; scipy is a
third-party dependency of the reference, pinned only as "scipy" in its requirements).  scipy's solver is the
rectangular shortest-augmenting-path algorithm of Crouse (2016) in ``scipy/optimize/rectangular_lsap``; it is not
vendored in /root/reference, so its published algorithm is restated here from memory of that source and PINNED by
testing against scipy itself (tests/test_oracle.py::test_lsa_*: exact equality of the index pairs, tie-heavy
integer matrices included).  Two functions:

* ``lsa_sequential`` — the algorithm as scipy runs it (one scan over the ``remaining`` columns per step).
* ``lsa_lane_parallel`` — the same algorithm with each scan split over 32 lanes and combined by the total order
  the CUDA kernel (efg_b200/csrc/lsa.cu) uses: lower value first; among equal values an unassigned column beats an
  assigned one, the LARGER position wins among unassigned, the SMALLER among assigned.  Its equality with scipy is
  what justifies the kernel's butterfly reduction.

Pure-Python loops: small problems only.  Never imported by efg_b200.
"""
import numpy as np


def _prepare(cost):
    cost = np.asarray(cost, dtype=np.float64)
    transpose = cost.shape[1] < cost.shape[0]
    return (cost.T.copy() if transpose else cost.copy()), transpose


def _finish(col4row, nr, transpose):
    if transpose:
        order = np.argsort(col4row, kind="stable")
        return col4row[order].astype(np.int64), order.astype(np.int64)
    return np.arange(nr, dtype=np.int64), col4row.astype(np.int64)


def _better(a, b):
    """Should candidate b = (value, position, unassigned) replace a under the kernel's total order?"""
    if a is None:
        return True
    if b[0] != a[0]:
        return b[0] < a[0]
    if a[2] and b[2]:
        return b[1] > a[1]
    if a[2] != b[2]:
        return b[2]
    return b[1] < a[1]


def _solve(cost, lanes):
    c, transpose = _prepare(cost)
    nr, nc = c.shape
    u, v = np.zeros(nr), np.zeros(nc)
    spc = np.empty(nc)
    path = np.full(nc, -1)
    col4row = np.full(nr, -1)
    row4col = np.full(nc, -1)
    for cur in range(nr):
        min_val, num_remaining = 0.0, nc
        remaining = [nc - it - 1 for it in range(nc)]
        sr, sc = np.zeros(nr, bool), np.zeros(nc, bool)
        spc[:] = np.inf
        sink, i = -1, cur
        while sink == -1:
            sr[i] = True
            best = None
            for lane in range(lanes):
                local = None
                for it in range(lane, num_remaining, lanes):
                    j = remaining[it]
                    r = ((min_val + c[i, j]) - u[i]) - v[j]
                    if r < spc[j]:
                        path[j], spc[j] = i, r
                    un = row4col[j] == -1
                    # the sequential rule of scipy, applied to this lane's subsequence
                    if local is None or spc[j] < local[0] or (spc[j] == local[0] and un):
                        local = (spc[j], it, un)
                if local is not None and _better(best, local):
                    best = local
            if best is None or not np.isfinite(best[0]):
                raise ValueError("cost matrix is infeasible")
            min_val, index = best[0], best[1]
            j = remaining[index]
            if row4col[j] == -1:
                sink = j
            else:
                i = row4col[j]
            sc[j] = True
            num_remaining -= 1
            remaining[index] = remaining[num_remaining]
        u[cur] += min_val
        for i2 in range(nr):
            if sr[i2] and i2 != cur:
                u[i2] += min_val - spc[col4row[i2]]
        for j2 in range(nc):
            if sc[j2]:
                v[j2] -= min_val - spc[j2]
        j = sink
        while True:
            i2 = path[j]
            row4col[j] = i2
            col4row[i2], j = j, col4row[i2]
            if i2 == cur:
                break
    return _finish(col4row, nr, transpose)


def lsa_sequential(cost):
    """scipy.optimize.linear_sum_assignment(cost) for finite costs: (row_ind, col_ind), rows ascending."""
    return _solve(cost, lanes=1)


def lsa_lane_parallel(cost, lanes=32):
    return _solve(cost, lanes=lanes)
