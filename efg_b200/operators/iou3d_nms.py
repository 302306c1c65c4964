"""BEV IoU of rotated boxes and rotated NMS: the Python surface of efg/operators/iou3d_nms.py (boxes_iou_bev :39-53,
boxes_iou3d_gpu :56-88, nms_gpu :91-108, nms_normal_gpu :111-126) over this repo's kernels (csrc/iou3d.cu).

`nms_gpu` / `nms_normal_gpu` keep the reference's return convention `(selected indices, None)`; the selection itself
never leaves the device here (the reference copies an N x N/64 mask to the host and scans it there), only the count is
read back to size the result."""
import torch

from .. import ops


def boxes_iou_bev(boxes_a, boxes_b):
    """boxes [N,7] / [M,7] (x, y, z, dx, dy, dz, heading) -> [N,M] BEV IoU."""
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    return ops.boxes_bev(boxes_a.float().contiguous(), boxes_b.float().contiguous())


def boxes_iou3d_gpu(boxes_a, boxes_b):
    """3-D IoU = BEV overlap x height overlap / union volume (iou3d_nms.py:56-88)."""
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    a_max, a_min = (boxes_a[:, 2] + boxes_a[:, 5] / 2).view(-1, 1), (boxes_a[:, 2] - boxes_a[:, 5] / 2).view(-1, 1)
    b_max, b_min = (boxes_b[:, 2] + boxes_b[:, 5] / 2).view(1, -1), (boxes_b[:, 2] - boxes_b[:, 5] / 2).view(1, -1)
    overlaps_bev = ops.boxes_bev(boxes_a.float().contiguous(), boxes_b.float().contiguous(), overlap=True)
    overlaps_h = torch.clamp(torch.min(a_max, b_max) - torch.max(a_min, b_min), min=0)
    overlaps_3d = overlaps_bev * overlaps_h
    vol_a = (boxes_a[:, 3] * boxes_a[:, 4] * boxes_a[:, 5]).view(-1, 1)
    vol_b = (boxes_b[:, 3] * boxes_b[:, 4] * boxes_b[:, 5]).view(1, -1)
    return overlaps_3d / torch.clamp(vol_a + vol_b - overlaps_3d, min=1e-6)


def _nms(boxes, scores, thresh, pre_maxsize, normal):
    assert boxes.shape[1] == 7
    order = scores.sort(0, descending=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    kept, count = ops.nms_bev(boxes[order].float().contiguous(), thresh, normal=normal)
    return order[kept[:int(count.item())]].contiguous(), None


def nms_gpu(boxes, scores, thresh, pre_maxsize=None, **kwargs):
    return _nms(boxes, scores, thresh, pre_maxsize, False)


def nms_normal_gpu(boxes, scores, thresh, **kwargs):
    return _nms(boxes, scores, thresh, None, True)


def rotate_nms_pcdet(boxes, scores, thresh, pre_maxsize=None, post_max_size=None):
    """CP/box_torch_ops.py:239-264: CenterPoint boxes (x, y, z, l, w, h, theta) -> the kernel's convention
    (dx = w, dy = l, heading = -theta - pi/2), rotated NMS, at most post_max_size survivors."""
    import math

    boxes = boxes[:, [0, 1, 2, 4, 3, 5, -1]].clone()
    boxes[:, -1] = -boxes[:, -1] - math.pi / 2
    if boxes.shape[0] == 0:
        return torch.zeros((0,), dtype=torch.int64, device=boxes.device)
    selected, _ = _nms(boxes, scores, thresh, pre_maxsize, False)
    return selected[:post_max_size] if post_max_size is not None else selected
