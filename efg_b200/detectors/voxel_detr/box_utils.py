"""Box helpers used by the Voxel-DETR matcher and losses (VD/modules/utils.py:16-89)."""
import torch


def cxcyczlwh_to_corners(x):
    """(cx, cy, cz, l, w, h) -> (x0, y0, z0, x1, y1, z1)."""
    c, s = x[..., :3], x[..., 3:6]
    return torch.cat([c - 0.5 * s, c + 0.5 * s], dim=-1)


def _vol(x):
    """Product of the last-dim extents, written out: Tensor.prod's backward synchronises to look for zeros."""
    return x[..., 0] * x[..., 1] * x[..., 2]


def generalized_box3d_iou(a, b):
    """Axis-aligned 3-D GIoU between every pair: a [N,6], b [M,6] corner boxes -> [N,M]."""
    a = torch.nan_to_num(a)
    b = torch.nan_to_num(b)
    vol_a = _vol(a[:, 3:] - a[:, :3])
    vol_b = _vol(b[:, 3:] - b[:, :3])
    inter = _vol((torch.min(a[:, None, 3:], b[:, 3:]) - torch.max(a[:, None, :3], b[:, :3])).clamp(min=0))
    union = vol_a[:, None] + vol_b - inter
    iou = inter / union
    hull = _vol((torch.max(a[:, None, 3:], b[:, 3:]) - torch.min(a[:, None, :3], b[:, :3])).clamp(min=0))
    return iou - (hull - union) / hull


def generalized_box3d_iou_paired(a, b):
    """Same measure for matched pairs a[i] <-> b[i] ([N,6] each -> [N]); equals diag of the full matrix."""
    a = torch.nan_to_num(a)
    b = torch.nan_to_num(b)
    vol_a = _vol(a[:, 3:] - a[:, :3])
    vol_b = _vol(b[:, 3:] - b[:, :3])
    inter = _vol((torch.min(a[:, 3:], b[:, 3:]) - torch.max(a[:, :3], b[:, :3])).clamp(min=0))
    union = vol_a + vol_b - inter
    hull = _vol((torch.max(a[:, 3:], b[:, 3:]) - torch.min(a[:, :3], b[:, :3])).clamp(min=0))
    return inter / union - (hull - union) / hull


def inverse_sigmoid(x, eps=1e-5):
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def limit_period(val, offset=0.5, period=3.141592653589793):
    """efg/geometry/box_ops_torch.py:229."""
    return val - torch.floor(val / period + offset) * period


def sigmoid_focal_loss(logits, targets, alpha: float = -1, gamma: float = 2, reduction: str = "none"):
    """efg/modeling/losses/focal_loss.py:5-45."""
    p = torch.sigmoid(logits)
    ce = torch.nn.functional.binary_cross_entropy_with_logits(logits, targets, reduction="none")
    p_t = p * targets + (1 - p) * (1 - targets)
    loss = ce * ((1 - p_t) ** gamma)
    if alpha >= 0:
        loss = (alpha * targets + (1 - alpha) * (1 - targets)) * loss
    if reduction == "mean":
        return loss.mean()
    if reduction == "sum":
        return loss.sum()
    return loss
