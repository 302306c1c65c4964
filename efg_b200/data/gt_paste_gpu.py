"""GT-database paste ("copy-paste" augmentation) with the point data resident on the device.

Mirror of the reference's `DatabaseSampling` processor + `DataBaseSampler` (efg/data/augmentations/extend_3d.py:49-92,
efg/data/samplers/gt_database_sampler.py:69-212):

  * which objects to try, and the collision test among them and the scene's boxes, are host logic over at most a few
    dozen boxes whose result (the pasted boxes and names) feeds the host-side target encoding — they stay on the host,
    vectorised (`collision_matrix`, `select`);
  * the points — every database object's points, packed once into one device tensor — are gathered, translated and
    concatenated with the scene's points by ONE kernel (`csrc/augment.cu: paste_points_kernel`); with
    `rm_points_after_sample` the scene points inside a pasted box are moved out of range in the same pass (the voxelizer
    drops them) instead of being compacted away, so nothing is read back from the device.

The database itself is given in memory (`{class: [{"box3d_lidar", "points" | "path", ...}]}`); `from_infos` reads the
reference's on-disk format (`np.frombuffer(..., float32).reshape(-1, points_dim)`, gt_database_sampler.py:156-163).
"""
import numpy as np
import torch

from .. import _lib, ops


def corners_2d(boxes):
    """[N,7] (x,y,z,l,w,h,yaw) -> BEV corners [N,4,2], clockwise from the minimum corner (box_ops.py:139-182,561-577)."""
    dims = boxes[:, 3:5]
    norm = np.array([[0, 0], [0, 1], [1, 1], [1, 0]], dtype=boxes.dtype) - np.array(0.5, dtype=boxes.dtype)
    corners = dims[:, None, :] * norm[None]
    s, c = np.sin(boxes[:, -1]), np.cos(boxes[:, -1])
    rot_t = np.stack([c, s, -s, c]).reshape(2, 2, -1)
    return np.einsum("aij,jka->aik", corners, rot_t) + boxes[:, None, 0:2]


def collision_matrix(corners):
    """[N,4,2] -> [N,N] bool, the reference's `box_collision_test(c, c)` (box_ops.py:27-95) vectorised: stand-up boxes
    overlap AND (two edges cross OR one box contains all corners of the other)."""
    n = corners.shape[0]
    lo, hi = corners.min(1), corners.max(1)
    iw = np.minimum(hi[:, None, 0], hi[None, :, 0]) - np.maximum(lo[:, None, 0], lo[None, :, 0])
    ih = np.minimum(hi[:, None, 1], hi[None, :, 1]) - np.maximum(lo[:, None, 1], lo[None, :, 1])
    near = (iw > 0) & (ih > 0)
    nxt = [1, 2, 3, 0]
    A = corners[:, None, :, None, :]                 # [N,1,4,1,2] edge starts of box i
    B = corners[:, nxt][:, None, :, None, :]
    C = corners[None, :, None, :, :]                 # [1,N,1,4,2] edge starts of box j
    D = corners[:, nxt][None, :, None, :, :]

    def ccw(p, q, r):   # the reference's strict orientation predicate
        return (r[..., 1] - p[..., 1]) * (q[..., 0] - p[..., 0]) > (q[..., 1] - p[..., 1]) * (r[..., 0] - p[..., 0])

    cross = ((ccw(A, C, D) != ccw(B, C, D)) & (ccw(A, B, C) != ccw(A, B, D))).any((2, 3))

    def contains(outer, inner):   # outer [N,4,2], inner [N,4,2] -> [N(outer), N(inner)]
        vec = -(outer - outer[:, nxt])                                        # clockwise
        ox = outer[:, None, :, None, 0] - inner[None, :, None, :, 0]          # [No,Ni,4 corners of outer,4 points of inner]
        oy = outer[:, None, :, None, 1] - inner[None, :, None, :, 1]
        crossp = vec[:, None, :, None, 1] * ox - vec[:, None, :, None, 0] * oy
        return (crossp < 0).all((2, 3))

    inside = contains(corners, corners)
    return near & (cross | inside | inside.T) & ~np.eye(n, dtype=bool)


def select(avoid_boxes, cand_boxes):
    """DataBaseSampler.sample_class (gt_database_sampler.py:182-212): a candidate that collides with anything still in
    play — the boxes to avoid, earlier accepted candidates, LATER candidates not yet rejected — is dropped."""
    num_gt = avoid_boxes.shape[0]
    boxes = np.concatenate([avoid_boxes, cand_boxes], axis=0)
    coll = collision_matrix(corners_2d(boxes))
    keep = np.zeros(cand_boxes.shape[0], dtype=bool)
    for i in range(num_gt, boxes.shape[0]):
        if coll[i].any():
            coll[i] = False
            coll[:, i] = False
        else:
            keep[i - num_gt] = True
    return keep


def box_planes(boxes):
    """[M,7] -> [M,6,4] float32 (normal, d) of the six faces, normals pointing inwards, built exactly as the reference
    builds them (center_to_corner_box3d -> corner_to_surfaces_3d -> surface_equ_3d_jitv2, box_ops.py:115-136,202-221,
    285-310) so that the sign test of a point agrees bit for bit."""
    boxes = np.asarray(boxes, dtype=np.float32)
    dims = boxes[:, 3:6]
    norm = np.stack(np.unravel_index(np.arange(8), [2] * 3), axis=1).astype(np.float32)[[0, 1, 3, 2, 4, 5, 7, 6]]
    corners = dims.reshape(-1, 1, 3) * (norm - np.array((0.5, 0.5, 0.5), dtype=np.float32)).reshape(1, 8, 3)
    s, c = np.sin(boxes[:, -1]), np.cos(boxes[:, -1])
    one, zero = np.ones_like(c), np.zeros_like(c)
    rot_t = np.stack([[c, s, zero], [-s, c, zero], [zero, zero, one]])
    corners = np.einsum("aij,jka->aik", corners, rot_t) + boxes[:, None, :3]
    faces = [(0, 1, 2), (7, 6, 5), (0, 3, 7), (1, 5, 6), (0, 4, 5), (3, 2, 6)]   # first three corners of each surface
    out = np.zeros((boxes.shape[0], 6, 4), dtype=np.float32)
    for k, (a, b, d_) in enumerate(faces):
        p0, p1, p2 = corners[:, a], corners[:, b], corners[:, d_]
        sv0, sv1 = p0 - p1, p1 - p2
        n = np.stack([sv0[:, 1] * sv1[:, 2] - sv0[:, 2] * sv1[:, 1], sv0[:, 2] * sv1[:, 0] - sv0[:, 0] * sv1[:, 2],
                      sv0[:, 0] * sv1[:, 1] - sv0[:, 1] * sv1[:, 0]], 1)
        out[:, k, :3] = n
        out[:, k, 3] = -p0[:, 0] * n[:, 0] - p0[:, 1] * n[:, 1] - p0[:, 2] * n[:, 2]
    return out


class BatchSampler:
    """Order in which database entries of one class are handed out (gt_database_sampler.py:16-66): a shuffled index list,
    padded to a multiple of the world size and sharded by rank, consumed `num` at a time; when fewer than `num + 1`
    entries are left the rest is returned (possibly fewer than asked for) and the shard is reshuffled.  Draws from
    `np.random` exactly as the reference does, so a seeded run picks the same objects."""

    def __init__(self, n, shuffle=True, rank=0, world=1):
        self.num_samples = int(np.ceil(n * 1.0 / world))
        total = self.num_samples * world
        indices = np.arange(n).tolist()
        if shuffle:
            np.random.shuffle(indices)
        indices += indices[:(total - n)]
        self._indices = indices[self.num_samples * rank:self.num_samples * (rank + 1)]
        self._idx = 0
        self._shuffle = shuffle

    def sample(self, num):
        if self._idx + num >= self.num_samples:
            ret = self._indices[self._idx:].copy()
            if self._shuffle:
                np.random.shuffle(self._indices)
            self._idx = 0
        else:
            ret = self._indices[self._idx:self._idx + num]
            self._idx += num
        return ret


class GpuGtDatabase:
    """`groups`: the reference's sample_groups, e.g. [{"VEHICLE": 15}, {"PEDESTRIAN": 10}];  `db`: {class: [info, ...]}
    with info = {"box3d_lidar": [7], "points": [n, F] float32, ...} (points relative to the box centre, as stored by the
    reference's database creation).  `pick(name, num)` decides which entries to try: default = the reference's
    BatchSampler order without shuffling; pass a callable to plug in a shuffled sampler."""

    def __init__(self, db, groups, device, pick=None):
        self.device = torch.device(device)
        self.classes = [list(g.keys())[0] for g in groups]
        self.max_nums = [int(list(g.values())[0]) for g in groups]
        self.boxes, self.counts, self.starts = {}, {}, {}
        chunks, at = [], 0
        for name in self.classes:
            infos = db.get(name, [])
            self.boxes[name] = np.stack([np.asarray(i["box3d_lidar"], dtype=np.float32) for i in infos]) if infos else np.zeros((0, 7), np.float32)
            cnt = np.array([int(np.asarray(i["points"]).shape[0]) for i in infos], dtype=np.int64)
            self.counts[name] = cnt
            self.starts[name] = at + np.concatenate([[0], np.cumsum(cnt)[:-1]]) if len(cnt) else np.zeros((0,), np.int64)
            at += int(cnt.sum())
            chunks += [np.asarray(i["points"], dtype=np.float32) for i in infos]
        self.nfeat = int(chunks[0].shape[1]) if chunks else 5
        packed = np.concatenate(chunks, 0) if chunks else np.zeros((0, self.nfeat), np.float32)
        self.points = torch.from_numpy(packed).to(self.device)      # resident for the whole run
        self._cursor = {name: 0 for name in self.classes}
        self._pick = pick or self._in_order

    def use_reference_sampling(self, rank=0, world=1, shuffle=True):
        """Hand out entries like the reference's per-class BatchSampler (shuffled, sharded by rank; draws np.random)."""
        samplers = {name: BatchSampler(self.boxes[name].shape[0], shuffle, rank, world) for name in self.classes}
        self._pick = lambda name, num: samplers[name].sample(num)
        return self

    @classmethod
    def from_infos(cls, db_infos, root_path, groups, device, points_dim, min_points=0, difficulty=-1, pick=None):
        """The reference's pickled `db_infos` (gt_database_sampler.py:83-109): filtered, point files read once."""
        import os

        db = {}
        for name, infos in db_infos.items():
            keep = [i for i in infos if i["num_points_in_gt"] >= min_points and i["difficulty"] >= difficulty]
            for i in keep:
                raw = open(os.path.join(root_path, i["path"]), "rb").read()
                i = dict(i, points=np.frombuffer(raw, np.float32).copy().reshape(-1, points_dim))
                db.setdefault(name, []).append(i)
        return cls(db, groups, device, pick=pick)

    def _in_order(self, name, num):
        n = self.boxes[name].shape[0]
        idx = [(self._cursor[name] + k) % max(n, 1) for k in range(min(num, n))]
        self._cursor[name] = (self._cursor[name] + len(idx)) % max(n, 1)
        return idx

    def sample(self, gt_boxes, gt_names):
        """DataBaseSampler.sample_all (gt_database_sampler.py:111-146) -> list of (class, database index) in paste order."""
        avoid = np.asarray(gt_boxes, dtype=np.float32).reshape(-1, 7)
        out = []
        for name, max_num in zip(self.classes, self.max_nums):
            num = int(max_num - np.sum([n == name for n in gt_names]))
            if num <= 0 or self.boxes[name].shape[0] == 0:
                continue
            idx = np.asarray(self._pick(name, num), dtype=np.int64)
            cand = self.boxes[name][idx]
            keep = select(avoid, cand)
            out += [(name, int(i)) for i in idx[keep]]
            if keep.any():
                avoid = np.concatenate([avoid, cand[keep]], axis=0)
        return out

    def paste(self, points, gt_boxes, gt_names, rm_points_after_sample=False):
        """points: device tensor [N, F].  -> (device points [n_paste + N, F], gt_boxes, gt_names) with the pasted objects
        appended to the annotations (DatabaseSampling.__call__, extend_3d.py:68-92).  No device-to-host traffic."""
        ops._check(points, "points", torch.float32)
        picked = self.sample(gt_boxes, gt_names)
        if not picked:
            return points, gt_boxes, gt_names
        boxes = np.stack([self.boxes[n][i] for n, i in picked])
        counts = np.array([self.counts[n][i] for n, i in picked], dtype=np.int64)
        starts = np.array([self.starts[n][i] for n, i in picked], dtype=np.int64)
        dst = np.concatenate([[0], np.cumsum(counts)[:-1]])
        n_paste, n_scene = int(counts.sum()), int(points.shape[0])
        table = torch.from_numpy(np.stack([starts, dst, counts], 1).astype(np.int32)).pin_memory().to(self.device, non_blocking=True)
        centers = torch.from_numpy(np.ascontiguousarray(boxes[:, :3])).pin_memory().to(self.device, non_blocking=True)
        planes, n_rm = None, 0
        if rm_points_after_sample:
            planes = torch.from_numpy(box_planes(np.nan_to_num(boxes))).pin_memory().to(self.device, non_blocking=True)
            n_rm = int(boxes.shape[0])
        out = torch.empty((n_paste + n_scene, self.nfeat), dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib().efgb_paste_points(ops._p(self.points), ops._p(table), ops._p(centers), len(picked), n_paste,
                                                ops._p(points), n_scene, self.nfeat, ops._p(planes), n_rm, ops._p(out),
                                                ops._stream()), "paste_points")
        names = np.concatenate([np.asarray(gt_names), np.array([n for n, _ in picked])])
        return out, np.nan_to_num(np.concatenate([np.asarray(gt_boxes, np.float32).reshape(-1, 7), boxes], 0)), names
