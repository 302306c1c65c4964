"""Set-prediction losses of Voxel-DETR (VD/losses.py:11-156): sigmoid focal classification,
L1 on the 6 box parameters, axis-aligned 3-D GIoU, L1 on the heading — each normalised by the
(world-averaged) number of ground-truth boxes; auxiliary copies for every intermediate decoder
layer."""
import torch
import torch.distributed as dist
from torch import nn
from torch.nn import functional as F

from .box_utils import cxcyczlwh_to_corners, generalized_box3d_iou_paired, sigmoid_focal_loss


def _world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _src_idx(indices):
    batch = torch.cat([torch.full_like(src, i) for i, (src, _) in enumerate(indices)])
    src = torch.cat([src for src, _ in indices])
    return batch, src


class TargetList(list):
    """Per-scene target dicts plus their concatenation (``labels_cat`` [sum K], ``boxes_cat`` [sum K, 7],
    ``offsets`` python ints), so that losses index ground truth with ONE device index tensor per layer."""

    labels_cat = None
    boxes_cat = None
    offsets = None


class MatchIndex:
    """The Hungarian assignments of one output layer as device tensors: batch index, query index and the
    index into the concatenated ground truth.  All layers of a step are uploaded in one pinned,
    non-blocking copy (``upload_matches``) instead of one blocking copy per (layer, scene, loss)."""

    def __init__(self, batch, src, tgt):
        self.batch, self.src, self.tgt = batch, src, tgt


def upload_matches(solved_layers, offsets, device):
    """solved_layers: list (per layer) of lists (per scene) of (src_idx, tgt_idx) CPU int64 tensors."""
    rows, sizes = [], []
    for layer in solved_layers:
        b = torch.cat([torch.full_like(src, i) for i, (src, _) in enumerate(layer)])
        s_ = torch.cat([src for src, _ in layer])
        t = torch.cat([tgt + offsets[i] for i, (_, tgt) in enumerate(layer)])
        rows.append(torch.stack([b, s_, t]))
        sizes.append(b.numel())
    flat = torch.cat(rows, dim=1) if rows else torch.zeros((3, 0), dtype=torch.int64)
    if device.type == "cuda":
        flat = flat.pin_memory().to(device, non_blocking=True)
    out, off = [], 0
    for n in sizes:
        out.append(MatchIndex(flat[0, off:off + n], flat[1, off:off + n], flat[2, off:off + n]))
        off += n
    return out


def device_matches(mats, num_layers, offsets, device):
    """The assignments of every layer solved ON THE DEVICE (efg_b200.ops.lsa_batched — scipy's algorithm and
    tie-breaking, csrc/lsa.cu): no device-to-host copy of the cost matrices, no pipeline drain.  ``mats`` in
    layer-major, scene-minor order; returns one MatchIndex per layer.  The batch index and the ground-truth
    offsets of the pairs depend only on the (host-known) matrix shapes."""
    from ... import ops

    rows, cols, sizes = ops.lsa_batched(mats)
    bs = len(mats) // max(num_layers, 1)
    # pair counts may differ between layers (min(Q_layer, K): encoder top-k count vs decoder queries): per-layer
    # segments come from a cumulative sum of `sizes`, not from a uniform stride
    b_parts, o_parts, bounds, off = [], [], [], 0
    for l in range(num_layers):
        per_layer = sizes[l * bs:(l + 1) * bs]
        n = sum(per_layer)
        bounds.append((off, off + n))
        off += n
        for i, k in enumerate(per_layer):
            b_parts.append(torch.full((k,), i, dtype=torch.int64))
            o_parts.append(torch.full((k,), offsets[i], dtype=torch.int64))
    if off:
        both = torch.stack([torch.cat(b_parts), torch.cat(o_parts)])
    else:
        both = torch.zeros((2, 0), dtype=torch.int64)
    both = both.pin_memory().to(device, non_blocking=True) if torch.device(device).type == "cuda" else both.to(device)
    return [MatchIndex(both[0, a:b], rows[a:b], cols[a:b] + both[1, a:b]) for a, b in bounds]


class ClassificationLoss(nn.Module):
    def __init__(self, focal_alpha):
        super().__init__()
        self.focal_alpha = focal_alpha
        self.target_classes = None
        self.src_logits = None

    def forward(self, outputs, targets, indices, num_boxes):
        logits = outputs["pred_logits"]
        dev = logits.device
        onehot = torch.zeros_like(logits)
        if isinstance(indices, MatchIndex):
            b_idx, s_idx = indices.batch, indices.src
            labels_cat = outputs.get("_labels_cat", targets.labels_cat)
            tgt_cls = labels_cat[indices.tgt]
        else:
            b_idx, s_idx = _src_idx(indices)
            b_idx, s_idx = b_idx.to(dev), s_idx.to(dev)
            tgt_cls = torch.cat([t["labels"][j.to(dev)] for t, (_, j) in zip(targets, indices)])
        self.target_classes = tgt_cls
        if "topk_indexes" in outputs:
            topk = outputs["topk_indexes"]
            self.src_logits = torch.gather(logits, 1, topk.expand(-1, -1, logits.shape[-1]))[b_idx, s_idx]
            onehot.index_put_((b_idx, topk[b_idx, s_idx].squeeze(-1), tgt_cls), onehot.new_ones(()))
        else:
            self.src_logits = logits[b_idx, s_idx]
            onehot.index_put_((b_idx, s_idx, tgt_cls), onehot.new_ones(()))
        loss = sigmoid_focal_loss(logits, onehot, alpha=self.focal_alpha, gamma=2.0, reduction="sum") / num_boxes
        return {"loss_ce": loss}


class RegressionLoss(nn.Module):
    def forward(self, outputs, targets, indices, num_boxes):
        boxes = outputs["pred_boxes"]
        dev = boxes.device
        if "topk_indexes" in outputs:
            boxes = torch.gather(boxes, 1, outputs["topk_indexes"].expand(-1, -1, boxes.shape[-1]))
        if isinstance(indices, MatchIndex):
            b_idx, s_idx = indices.batch, indices.src
            tgt = targets.boxes_cat[indices.tgt]
        else:
            b_idx, s_idx = _src_idx(indices)
            b_idx, s_idx = b_idx.to(dev), s_idx.to(dev)
            tgt = torch.cat([t["gt_boxes"][j.to(dev)] for t, (_, j) in zip(targets, indices)], dim=0)
        src_box, src_rad = boxes[b_idx, s_idx].split(6, dim=-1)
        tgt_box, tgt_rad = tgt.split(6, dim=-1)
        giou = generalized_box3d_iou_paired(cxcyczlwh_to_corners(src_box), cxcyczlwh_to_corners(tgt_box))
        return {
            "loss_bbox": F.l1_loss(src_box, tgt_box, reduction="none").sum() / num_boxes,
            "loss_giou": (1 - giou).sum() / num_boxes,
            "loss_rad": F.l1_loss(src_rad, tgt_rad, reduction="none").sum() / num_boxes,
        }


class Det3DLoss(nn.Module):
    def __init__(self, matcher, weight_dict, losses):
        super().__init__()
        self.matcher = matcher
        self.weight_dict = weight_dict
        self.losses = losses
        self.det3d_losses = nn.ModuleDict()
        self.det3d_enc_losses = nn.ModuleDict()
        for loss in losses:
            if loss == "boxes":
                self.det3d_losses[loss] = RegressionLoss()
                self.det3d_enc_losses[loss + "_enc"] = RegressionLoss()
            elif loss == "focal_labels":
                self.det3d_losses[loss] = ClassificationLoss(0.25)
                self.det3d_enc_losses[loss + "_enc"] = ClassificationLoss(0.25)
            else:
                raise ValueError("Only boxes|focal_labels are supported for det3d losses. Found {}".format(loss))

    def get_target_classes(self):
        for k in self.det3d_losses.keys():
            if "labels" in k:
                return self.det3d_losses[k].src_logits, self.det3d_losses[k].target_classes

    @staticmethod
    def normaliser(targets, device):
        """Mean number of GT boxes per rank, >= 1 (VD/losses.py:121-125): a host float in a single process, a 0-dim
        device tensor (no host synchronisation) when the job is distributed over GPUs.  If the model requested the
        all-reduce at the start of forward (``request_normaliser``) that result is used."""
        pending = getattr(targets, "_num_boxes_async", None)
        if pending is None:
            pending = Det3DLoss.request_normaliser(targets, device)
        return pending.result(floor=1.0)

    @staticmethod
    def request_normaliser(targets, device):
        """Issue the 1-float all-reduce asynchronously; it depends only on the targets, so it can run under the whole
        forward pass instead of blocking in the middle of the loss."""
        from ...parallel import AsyncMean

        pending = AsyncMean(float(sum(len(t["labels"]) for t in targets)), device)
        try:
            targets._num_boxes_async = pending
        except AttributeError:  # a plain list: nothing to attach to
            pass
        return pending

    @staticmethod
    def layers_of(outputs):
        return list(outputs.get("aux_outputs", [])) + [{k: v for k, v in outputs.items() if k != "aux_outputs"}]

    def prepare(self, outputs, targets):
        """Cost matrices (device tensors) of every output layer, in layer-major, scene-minor order."""
        mats = []
        for lo in self.layers_of(outputs):
            mats.extend(self.matcher.cost_matrices(lo, targets))
        return mats

    def finish_stacked(self, logits, boxes, targets, solved, num_boxes):
        """Losses of ALL decoder layers in one pass.  logits [L,B,Q,C], boxes [L,B,Q,7]; solved: L MatchIndex
        objects (device).  Same elementwise expressions as ``finish`` (ClassificationLoss / RegressionLoss layer by
        layer, VD/losses.py:11-96) evaluated on the layer-stacked tensors, followed by per-layer sums — L-fold fewer
        kernel launches in the launch-bound tail of the step.  Keys and suffixes as in ``finish``."""
        L = logits.shape[0]
        n = solved[0].batch.numel()
        assert all(m.batch.numel() == n for m in solved), "every layer matches the same ground truth"
        dev = logits.device
        layer = torch.arange(L, device=dev).repeat_interleave(n)
        b_idx = torch.cat([m.batch for m in solved])
        s_idx = torch.cat([m.src for m in solved])
        t_idx = torch.cat([m.tgt for m in solved])
        out = {}
        for loss in self.losses:
            if loss == "focal_labels":
                tgt_cls = targets.labels_cat[t_idx]
                onehot = torch.zeros_like(logits)
                onehot.index_put_((layer, b_idx, s_idx, tgt_cls), onehot.new_ones(()))
                focal = sigmoid_focal_loss(logits, onehot, alpha=self.det3d_losses[loss].focal_alpha, gamma=2.0,
                                           reduction="none")
                out["loss_ce"] = focal.sum((1, 2, 3)) / num_boxes
                # what get_target_classes() reports: the matched logits / classes of the LAST layer
                self.det3d_losses[loss].src_logits = logits[L - 1][solved[-1].batch, solved[-1].src]
                self.det3d_losses[loss].target_classes = tgt_cls[(L - 1) * n:]
            else:
                src = boxes[layer, b_idx, s_idx]
                tgt = targets.boxes_cat[t_idx]
                src_box, src_rad = src.split(6, dim=-1)
                tgt_box, tgt_rad = tgt.split(6, dim=-1)
                giou = generalized_box3d_iou_paired(cxcyczlwh_to_corners(src_box), cxcyczlwh_to_corners(tgt_box))
                # explicit sizes: n may be 0 (a batch without ground truth), where view(L, -1) is ambiguous
                out["loss_bbox"] = F.l1_loss(src_box, tgt_box, reduction="none").view(L, n * src_box.shape[-1]).sum(1) / num_boxes
                out["loss_giou"] = (1 - giou).view(L, n).sum(1) / num_boxes
                out["loss_rad"] = F.l1_loss(src_rad, tgt_rad, reduction="none").view(L, n * src_rad.shape[-1]).sum(1) / num_boxes
        losses = {}
        for k, vec in out.items():
            for li in range(L):
                losses[k + ("" if li == L - 1 else "_{}".format(li))] = vec[li]
        return losses

    def finish(self, outputs, targets, solved, num_boxes):
        """solved: per layer, either a MatchIndex (device) or a list of per-scene (src, tgt) CPU pairs."""
        layers = self.layers_of(outputs)
        losses = {}
        for li, lo in enumerate(layers):
            suffix = "" if li == len(layers) - 1 else "_{}".format(li)
            if "_labels_cat" in outputs:
                lo = dict(lo, _labels_cat=outputs["_labels_cat"])
            for loss in self.losses:
                for k, v in self.det3d_losses[loss](lo, targets, solved[li], num_boxes).items():
                    losses[k + suffix] = v
        return losses

    def forward(self, outputs, targets, num_boxes=None):
        if num_boxes is None:
            num_boxes = self.normaliser(targets, next(iter(outputs.values())).device)
        # all cost matrices first (GPU), one host transfer, then the assignments
        solved = self.matcher.solve(self.prepare(outputs, targets))
        bs = len(targets)
        per_layer = [solved[i * bs:(i + 1) * bs] for i in range(len(solved) // max(bs, 1))]
        return self.finish(outputs, targets, per_layer, num_boxes)
