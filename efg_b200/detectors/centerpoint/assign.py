"""CenterPoint label assignment (CP/voxelnet.py:44-192, CP/center_utils.py:10-58): per task a class
heatmap with one Gaussian per object (radius from the CornerNet overlap rule), and per object the flat
BEV index, class and regression target (dx, dy, z, log l, log w, log h, vx, vy, sin, cos).

``assign_scene`` is the reference's host loop (numpy; used on the CPU path and as the checker of the device path).
``assign_batch_device`` is the CUDA path: the per-object scalars are computed vectorised on the host (a few hundred boxes,
no per-object Python loop) and uploaded in ONE pinned copy; the heatmaps are never built on the host — one kernel draws
every Gaussian of the batch into zero-initialised device maps (csrc/iou3d.cu: draw_gaussians_kernel)."""
import numpy as np


def gaussian_radius(det_size, min_overlap=0.5):
    height, width = det_size
    b1 = height + width
    c1 = width * height * (1 - min_overlap) / (1 + min_overlap)
    r1 = (b1 + np.sqrt(b1 ** 2 - 4 * c1)) / 2
    b2 = 2 * (height + width)
    c2 = (1 - min_overlap) * width * height
    r2 = (b2 + np.sqrt(b2 ** 2 - 16 * c2)) / 2
    a3 = 4 * min_overlap
    b3 = -2 * min_overlap * (height + width)
    c3 = (min_overlap - 1) * width * height
    r3 = (b3 + np.sqrt(b3 ** 2 - 4 * a3 * c3)) / 2
    return min(r1, r2, r3)


def gaussian2d(shape, sigma=1.0):
    m, n = [(ss - 1.0) / 2.0 for ss in shape]
    y, x = np.ogrid[-m:m + 1, -n:n + 1]
    h = np.exp(-(x * x + y * y) / (2 * sigma * sigma))
    h[h < np.finfo(h.dtype).eps * h.max()] = 0
    return h


def draw_gaussian(heatmap, center, radius, k=1):
    diameter = 2 * radius + 1
    gaussian = gaussian2d((diameter, diameter), sigma=diameter / 6)
    x, y = int(center[0]), int(center[1])
    height, width = heatmap.shape[0:2]
    left, right = min(x, radius), min(width - x, radius + 1)
    top, bottom = min(y, radius), min(height - y, radius + 1)
    dst = heatmap[y - top:y + bottom, x - left:x + right]
    src = gaussian[radius - top:radius + bottom, radius - left:radius + right]
    if min(src.shape) > 0 and min(dst.shape) > 0:
        np.maximum(dst, src * k, out=dst)
    return heatmap


def limit_period(val, offset=0.5, period=np.pi):
    return val - np.floor(val / period + offset) * period


def assign_scene(annotations, tasks, grid_size, pc_range, voxel_size, out_size_factor, gaussian_overlap, max_objs,
                 min_radius):
    """-> dict(hm=[task][C,H,W], anno_box=[task][max_objs,10], ind, mask, cat) for one scene."""
    class_names_by_task = [list(t["class_names"]) for t in tasks]
    plain = [n for names in class_names_by_task for n in names]
    names = np.asarray(annotations["gt_names"])
    keep = np.array([n in plain for n in names], dtype=bool)
    boxes = np.asarray(annotations["gt_boxes"], dtype=np.float32)[keep]
    names = names[keep]
    classes = np.array([plain.index(n) + 1 for n in names], dtype=np.int32)
    fmap = np.asarray(grid_size[:2]) // out_size_factor  # (W, H)

    out = {"hm": [], "anno_box": [], "ind": [], "mask": [], "cat": []}
    flag = 0
    for cnames in class_names_by_task:
        sel = [np.where(classes == cnames.index(c) + 1 + flag)[0] for c in cnames]
        order = np.concatenate(sel) if sel else np.zeros((0,), dtype=np.int64)
        tboxes = boxes[order].copy()
        tcls = classes[order] - flag
        flag += len(cnames)
        if tboxes.shape[0]:
            tboxes[:, -1] = limit_period(tboxes[:, -1], offset=0.5, period=np.pi * 2)
        hm = np.zeros((len(cnames), int(fmap[1]), int(fmap[0])), dtype=np.float32)
        anno = np.zeros((max_objs, 10), dtype=np.float32)
        ind = np.zeros((max_objs,), dtype=np.int64)
        mask = np.zeros((max_objs,), dtype=np.uint8)
        cat = np.zeros((max_objs,), dtype=np.int64)
        for k in range(min(tboxes.shape[0], max_objs)):
            cls_id = int(tcls[k]) - 1
            L = tboxes[k, 3] / voxel_size[0] / out_size_factor
            W = tboxes[k, 4] / voxel_size[1] / out_size_factor
            if not (L > 0 and W > 0):
                continue
            radius = max(min_radius, int(gaussian_radius((L, W), min_overlap=gaussian_overlap)))
            x, y, z = tboxes[k, 0], tboxes[k, 1], tboxes[k, 2]
            ct = np.array([(x - pc_range[0]) / voxel_size[0] / out_size_factor,
                           (y - pc_range[1]) / voxel_size[1] / out_size_factor], dtype=np.float32)
            ct_int = ct.astype(np.int32)
            if not (0 <= ct_int[0] < fmap[0] and 0 <= ct_int[1] < fmap[1]):
                continue
            draw_gaussian(hm[cls_id], ct, radius)
            cat[k] = cls_id
            ind[k] = ct_int[1] * fmap[0] + ct_int[0]
            mask[k] = 1
            rot = tboxes[k, -1]
            anno[k] = np.concatenate((ct - ct_int, [z], np.log(tboxes[k, 3:6]), tboxes[k, 6:8], [np.sin(rot)], [np.cos(rot)]),
                                     axis=None)
        for key, val in (("hm", hm), ("anno_box", anno), ("ind", ind), ("mask", mask), ("cat", cat)):
            out[key].append(val)
    return out


def _gaussian_radius_vec(h, w, min_overlap):
    b1 = h + w
    c1 = w * h * (1 - min_overlap) / (1 + min_overlap)
    r1 = (b1 + np.sqrt(b1 ** 2 - 4 * c1)) / 2
    b2 = 2 * (h + w)
    c2 = (1 - min_overlap) * w * h
    r2 = (b2 + np.sqrt(b2 ** 2 - 16 * c2)) / 2
    a3 = 4 * min_overlap
    b3 = -2 * min_overlap * (h + w)
    c3 = (min_overlap - 1) * w * h
    r3 = (b3 + np.sqrt(b3 ** 2 - 4 * a3 * c3)) / 2
    return np.minimum(np.minimum(r1, r2), r3)


def assign_batch_device(infos, tasks, grid_size, pc_range, voxel_size, out_size_factor, gaussian_overlap, max_objs,
                        min_radius, device):
    """The targets of ``assign_scene`` for a whole batch as device tensors:
    {"hm": [task][B,C,H,W], "anno_box": [task][B,max_objs,10], "ind", "mask", "cat"}."""
    import torch

    from ... import _lib, ops

    class_names_by_task = [list(t["class_names"]) for t in tasks]
    plain = [n for names in class_names_by_task for n in names]
    lookup = {n: i + 1 for i, n in enumerate(plain)}
    fmap = np.asarray(grid_size[:2]) // out_size_factor  # (W, H)
    W_, H_ = int(fmap[0]), int(fmap[1])
    B, T = len(infos), len(tasks)
    # small per-object targets: one float and one int block for the whole batch, uploaded once
    fblock = np.zeros((T, B, max_objs, 10), dtype=np.float32)
    iblock = np.zeros((T, 3, B, max_objs), dtype=np.int64)     # ind, mask, cat
    plane_base, planes = [], 0
    for cnames in class_names_by_task:
        plane_base.append(planes)
        planes += B * len(cnames)
    objs = []
    for b, info in enumerate(infos):
        ann = info["annotations"]
        names = np.asarray(ann["gt_names"])
        cls_all = np.array([lookup.get(n, 0) for n in names], dtype=np.int32)
        keep = cls_all > 0
        boxes = np.asarray(ann["gt_boxes"], dtype=np.float32)[keep]
        classes = cls_all[keep]
        flag = 0
        for t, cnames in enumerate(class_names_by_task):
            local = classes - flag
            in_task = (local >= 1) & (local <= len(cnames))
            # the reference concatenates the objects class by class (stable inside a class)
            order = np.argsort(np.where(in_task, local, len(cnames) + 1), kind="stable")[:int(in_task.sum())]
            flag += len(cnames)
            order = order[:max_objs]
            if order.size == 0:
                continue
            tb = boxes[order].copy()
            tcls = local[order] - 1
            tb[:, -1] = limit_period(tb[:, -1], offset=0.5, period=np.pi * 2)
            L = tb[:, 3] / voxel_size[0] / out_size_factor
            Wd = tb[:, 4] / voxel_size[1] / out_size_factor
            ok = (L > 0) & (Wd > 0)
            with np.errstate(invalid="ignore"):
                rad = np.maximum(min_radius, _gaussian_radius_vec(L, Wd, gaussian_overlap).astype(np.int64)).astype(np.int32)
            ct = np.stack([(tb[:, 0] - pc_range[0]) / voxel_size[0] / out_size_factor,
                           (tb[:, 1] - pc_range[1]) / voxel_size[1] / out_size_factor], 1).astype(np.float32)
            cti = ct.astype(np.int32)
            ok &= (cti[:, 0] >= 0) & (cti[:, 0] < W_) & (cti[:, 1] >= 0) & (cti[:, 1] < H_)
            k = np.nonzero(ok)[0]
            if k.size == 0:
                continue
            rot = tb[k, -1]
            fblock[t, b, k] = np.concatenate([ct[k] - cti[k], tb[k, 2:3], np.log(tb[k, 3:6]), tb[k, 6:8], np.sin(rot)[:, None],
                                              np.cos(rot)[:, None]], 1)
            iblock[t, 0, b, k] = cti[k, 1].astype(np.int64) * W_ + cti[k, 0]
            iblock[t, 1, b, k] = 1
            iblock[t, 2, b, k] = tcls[k]
            plane = plane_base[t] + b * len(cnames) + tcls[k]
            objs.append(np.stack([plane, cti[k, 0], cti[k, 1], rad[k]], 1).astype(np.int32))
    obj = np.concatenate(objs, 0) if objs else np.zeros((0, 4), dtype=np.int32)
    dev = torch.device(device)
    up = [torch.from_numpy(fblock), torch.from_numpy(iblock), torch.from_numpy(np.ascontiguousarray(obj))]
    if dev.type == "cuda":
        up = [u.pin_memory().to(dev, non_blocking=True) for u in up]
    f_dev, i_dev, o_dev = up
    hm_all = torch.zeros((planes, H_, W_), dtype=torch.float32, device=dev)
    _lib.check(_lib.lib().efgb_draw_gaussians(ops._p(o_dev), int(obj.shape[0]), H_, W_, ops._p(hm_all), ops._stream()),
               "draw_gaussians")
    out = {"hm": [], "anno_box": [], "ind": [], "mask": [], "cat": []}
    for t, cnames in enumerate(class_names_by_task):
        c = len(cnames)
        out["hm"].append(hm_all[plane_base[t]:plane_base[t] + B * c].view(B, c, H_, W_))
        out["anno_box"].append(f_dev[t])
        out["ind"].append(i_dev[t, 0])
        out["mask"].append(i_dev[t, 1].to(torch.uint8))
        out["cat"].append(i_dev[t, 2])
    return out
