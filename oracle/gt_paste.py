"""CPU oracle (TEST INFRASTRUCTURE ONLY — never imported by efg_b200/) of the GT-database paste of the reference's
training pipeline.  Restates, loop for loop:

  * box corners            efg/geometry/box_ops.py:139-166 (corners_nd), :169-182 (rotation_2d), :185-199 (rotation_3d),
                           :115-136 (center_to_corner_box3d), :561-577 (center_to_corner_box2d)
  * collision test         box_ops.py:27-95 (box_collision_test: stand-up overlap, edge crossings, full containment)
  * greedy acceptance      efg/data/samplers/gt_database_sampler.py:182-212 (DataBaseSampler.sample_class) and
                           :111-146 (sample_all: classes in turn, accepted boxes join the boxes to avoid)
  * points in rotated box  box_ops.py:98-112, :202-221 (surfaces), :285-310 (plane equations), :337-371 (sign test)

Pinned by tests/golden/gt_paste_seed*.npz, written by tests/golden/make_golden_gt_paste.py from the reference's own
functions (numba) — tests/test_gt_paste.py."""
import numpy as np


def corners_nd(dims, origin=0.5):
    ndim = int(dims.shape[1])
    norm = np.stack(np.unravel_index(np.arange(2 ** ndim), [2] * ndim), axis=1).astype(dims.dtype)
    norm = norm[[0, 1, 3, 2]] if ndim == 2 else norm[[0, 1, 3, 2, 4, 5, 7, 6]]
    norm = norm - np.array(origin, dtype=dims.dtype)
    return dims.reshape([-1, 1, ndim]) * norm.reshape([1, 2 ** ndim, ndim])


def center_to_corner_box2d(centers, dims, angles):
    corners = corners_nd(dims, origin=0.5)
    s, c = np.sin(angles), np.cos(angles)
    rot_t = np.stack([c, s, -s, c]).reshape([2, 2, -1])
    corners = np.einsum("aij,jka->aik", corners, rot_t)
    return corners + centers.reshape([-1, 1, 2])


def center_to_corner_box3d(centers, dims, angles):
    corners = corners_nd(dims, origin=(0.5, 0.5, 0.5))
    s, c = np.sin(angles), np.cos(angles)
    one, zero = np.ones_like(c), np.zeros_like(c)
    rot_t = np.stack([[c, s, zero], [-s, c, zero], [zero, zero, one]])
    corners = np.einsum("aij,jka->aik", corners, rot_t)
    return corners + centers.reshape([-1, 1, 3])


def box_collision_test(boxes, qboxes, clockwise=True):
    """boxes [N,4,2], qboxes [K,4,2] corner arrays -> [N,K] bool."""
    n, k = boxes.shape[0], qboxes.shape[0]
    ret = np.zeros((n, k), dtype=bool)
    nxt = [1, 2, 3, 0]
    lo_b, hi_b = boxes.min(1), boxes.max(1)
    lo_q, hi_q = qboxes.min(1), qboxes.max(1)
    for i in range(n):
        for j in range(k):
            if not (min(hi_b[i, 0], hi_q[j, 0]) - max(lo_b[i, 0], lo_q[j, 0]) > 0):
                continue
            if not (min(hi_b[i, 1], hi_q[j, 1]) - max(lo_b[i, 1], lo_q[j, 1]) > 0):
                continue
            hit = False
            for a in range(4):
                A, B = boxes[i, a], boxes[i, nxt[a]]
                for b in range(4):
                    C, D = qboxes[j, b], qboxes[j, nxt[b]]
                    acd = (D[1] - A[1]) * (C[0] - A[0]) > (C[1] - A[1]) * (D[0] - A[0])
                    bcd = (D[1] - B[1]) * (C[0] - B[0]) > (C[1] - B[1]) * (D[0] - B[0])
                    if acd != bcd:
                        abc = (C[1] - A[1]) * (B[0] - A[0]) > (B[1] - A[1]) * (C[0] - A[0])
                        abd = (D[1] - A[1]) * (B[0] - A[0]) > (B[1] - A[1]) * (D[0] - A[0])
                        if abc != abd:
                            hit = True
                            break
                if hit:
                    break
            if not hit:
                def contains(outer, inner):
                    for L in range(4):
                        for c in range(4):
                            vec = outer[c] - outer[(c + 1) % 4]
                            if clockwise:
                                vec = -vec
                            cross = vec[1] * (outer[c, 0] - inner[L, 0]) - vec[0] * (outer[c, 1] - inner[L, 1])
                            if cross >= 0:
                                return False
                    return True
                hit = contains(boxes[i], qboxes[j]) or contains(qboxes[j], boxes[i])
            ret[i, j] = hit
    return ret


def sample_class(gt_boxes, cand_boxes):
    """-> bool [num_cand]: the candidates DataBaseSampler.sample_class keeps, given the boxes to avoid."""
    num_gt = gt_boxes.shape[0]
    boxes = np.concatenate([gt_boxes, cand_boxes], axis=0)
    bv = center_to_corner_box2d(boxes[:, 0:2], boxes[:, 3:5], boxes[:, -1])
    coll = box_collision_test(bv, bv)
    idx = np.arange(boxes.shape[0])
    coll[idx, idx] = False
    keep = np.zeros(cand_boxes.shape[0], dtype=bool)
    for i in range(num_gt, boxes.shape[0]):
        if coll[i].any():
            coll[i] = False
            coll[:, i] = False
        else:
            keep[i - num_gt] = True
    return keep


def sample_all(gt_boxes, gt_names, classes, max_nums, db_boxes):
    """classes in turn (gt_database_sampler.py:111-146); the sampler hands out the FIRST `num` entries of a class
    (the golden's sample_func).  -> list of (class, index in the class's database) in paste order."""
    avoid = gt_boxes
    out = []
    for name, max_num in zip(classes, max_nums):
        num = int(max_num - np.sum([n == name for n in gt_names]))
        if num <= 0:
            continue
        cand = db_boxes[name][:num]
        keep = sample_class(avoid, cand)
        out += [(name, int(i)) for i in np.nonzero(keep)[0]]
        if keep.any():
            avoid = np.concatenate([avoid, cand[keep]], axis=0)
    return out


def points_in_rbbox(points, rbbox):
    """-> [N, M] bool (box_ops.py:98-112 with origin (0.5, 0.5, 0.5), z axis 2)."""
    corners = center_to_corner_box3d(rbbox[:, :3], rbbox[:, 3:6], rbbox[:, -1])
    surf = np.array([[corners[:, 0], corners[:, 1], corners[:, 2], corners[:, 3]],
                     [corners[:, 7], corners[:, 6], corners[:, 5], corners[:, 4]],
                     [corners[:, 0], corners[:, 3], corners[:, 7], corners[:, 4]],
                     [corners[:, 1], corners[:, 5], corners[:, 6], corners[:, 2]],
                     [corners[:, 0], corners[:, 4], corners[:, 5], corners[:, 1]],
                     [corners[:, 3], corners[:, 2], corners[:, 6], corners[:, 7]]]).transpose([2, 0, 1, 3])
    m = surf.shape[0]
    normal = np.zeros((m, 6, 3), dtype=surf.dtype)
    d = np.zeros((m, 6), dtype=surf.dtype)
    for i in range(m):
        for j in range(6):
            sv0 = surf[i, j, 0] - surf[i, j, 1]
            sv1 = surf[i, j, 1] - surf[i, j, 2]
            normal[i, j, 0] = sv0[1] * sv1[2] - sv0[2] * sv1[1]
            normal[i, j, 1] = sv0[2] * sv1[0] - sv0[0] * sv1[2]
            normal[i, j, 2] = sv0[0] * sv1[1] - sv0[1] * sv1[0]
            d[i, j] = -surf[i, j, 0, 0] * normal[i, j, 0] - surf[i, j, 0, 1] * normal[i, j, 1] - surf[i, j, 0, 2] * normal[i, j, 2]
    pts = points[:, :3]
    ret = np.ones((pts.shape[0], m), dtype=bool)
    for j in range(m):
        for k in range(6):
            sign = pts[:, 0] * normal[j, k, 0] + pts[:, 1] * normal[j, k, 1] + pts[:, 2] * normal[j, k, 2] + d[j, k]
            ret[:, j] &= ~(sign >= 0)
    return ret


def paste(points, gt_boxes, gt_names, classes, max_nums, db_boxes, db_counts, db_points, rm_points=False):
    """DatabaseSampling.__call__ (extend_3d.py:68-92) over an in-memory database -> (points, boxes, names)."""
    picked = sample_all(gt_boxes, gt_names, classes, max_nums, db_boxes)
    if not picked:
        return points, gt_boxes, gt_names
    pts_list, boxes, names = [], [], []
    for name, i in picked:
        start = int(db_counts[name][:i].sum())
        p = db_points[name][start:start + int(db_counts[name][i])].copy()
        p[:, :3] += db_boxes[name][i][:3]
        pts_list.append(p)
        boxes.append(db_boxes[name][i])
        names.append(name)
    boxes = np.stack(boxes)
    if rm_points:
        points = points[~points_in_rbbox(points, np.nan_to_num(boxes)).any(-1)]
    out = np.nan_to_num(np.concatenate(pts_list + [points], axis=0))
    return out, np.nan_to_num(np.concatenate([gt_boxes, boxes], 0)), np.concatenate([gt_names, np.array(names)])
