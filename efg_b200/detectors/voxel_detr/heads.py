"""Detection head of Voxel-DETR (VD/heads.py:14-96): per decoder layer a 3-layer class MLP and a
3-layer box MLP whose output refines the reference window in logit space."""
import math

import torch
from torch import nn

from .box_utils import inverse_sigmoid
from .losses import Det3DLoss
from .matcher import HungarianMatcher3d
from .transformer import MLP, get_clones


class Accuracy:
    """top-1 accuracy of the matched queries in percent (VD/modules/metrics.py:35-58)."""

    def __call__(self, logits, target):
        with torch.no_grad():
            if target.numel() == 0:
                return {"accuracy": torch.zeros([], device=logits.device)}
            pred = logits.argmax(dim=1)
            acc = (pred == target).float().sum() * (100.0 / target.shape[0])
            if torch.distributed.is_available() and torch.distributed.is_initialized() and \
                    torch.distributed.get_world_size() > 1:
                acc = acc.clone()
                torch.distributed.reduce(acc, dst=0)
                if torch.distributed.get_rank() == 0:
                    acc = acc / torch.distributed.get_world_size()
            return {"accuracy": acc}


class Det3DHead(nn.Module):
    def __init__(self, config, with_aux=False, with_metrics=False, num_classes=3, num_layers=1):
        super().__init__()
        hidden = config.model.hidden_dim
        cls = MLP(hidden, hidden, num_classes, 3)
        box = MLP(hidden, hidden, 7, 3)
        prior = 0.01
        cls.layers[-1].bias.data = torch.ones(num_classes) * (-math.log((1 - prior) / prior))
        nn.init.constant_(box.layers[-1].weight.data, 0)
        nn.init.constant_(box.layers[-1].bias.data, 0)
        self.class_embed = get_clones(cls, num_layers)
        self.bbox_embed = get_clones(box, num_layers)

        mc = config.model.loss.matcher
        matcher = HungarianMatcher3d(cost_class=mc.class_weight, cost_bbox=mc.bbox_weight, cost_giou=mc.giou_weight,
                                     cost_rad=mc.rad_weight)
        weight_dict = {"loss_ce": config.model.loss.class_loss_coef, "loss_bbox": config.model.loss.bbox_loss_coef,
                       "loss_giou": config.model.loss.giou_loss_coef, "loss_rad": config.model.loss.rad_loss_coef}
        self.losses = Det3DLoss(matcher=matcher, weight_dict=weight_dict, losses=["focal_labels", "boxes"])
        if with_aux:
            aux = {k + "_enc_0": v for k, v in weight_dict.items()}
            for i in range(config.model.transformer.dec_layers - 1):
                aux.update({k + "_{}".format(i): v for k, v in weight_dict.items()})
            self.losses.weight_dict.update(aux)
        if with_metrics:
            self.metrics = {"accuracy": Accuracy()}

    def forward(self, embed, anchors, layer_idx=0):
        logits = self.class_embed[layer_idx](embed)
        boxes = (self.bbox_embed[layer_idx](embed) + inverse_sigmoid(anchors)).sigmoid()
        return logits, boxes

    def compute_losses(self, outputs, targets, num_boxes=None, solved=None, stacked=None):
        """stacked = (logits [L,B,Q,C], boxes [L,B,Q,7]) selects the all-layers-at-once evaluation
        (Det3DLoss.finish_stacked); ``outputs`` is then only used for its keys."""
        if solved is None:
            loss_dict = self.losses(outputs, targets, num_boxes)
        elif stacked is not None:
            loss_dict = self.losses.finish_stacked(stacked[0], stacked[1], targets, solved, num_boxes)
        else:
            loss_dict = self.losses.finish(outputs, targets, solved, num_boxes)
        weights = self.losses.weight_dict
        for k in list(loss_dict.keys()):
            if k in weights:
                loss_dict[k] = loss_dict[k] * weights[k]
        if hasattr(self, "metrics"):
            loss_dict.update(self.metrics["accuracy"](*self.losses.get_target_classes()))
        return loss_dict
