"""CenterHead (CP/center_head.py:19-171) and its losses (CP/centernet_loss.py:8-56): a shared 3x3 conv,
then per task separate conv heads for the class heatmap and the box regressions; penalty-reduced focal
loss on the heatmap and masked L1 on the regression maps gathered at the object centres."""
import copy

import torch
import torch.nn.functional as F
from torch import nn

from ...modeling.norm import get_norm
from ...modeling.rpn import Sequential


def _gather_at(feat, ind):
    """feat [B,C,H,W], ind [B,M] flat (y*W+x) -> [B,M,C]."""
    b, c = feat.shape[:2]
    flat = feat.permute(0, 2, 3, 1).reshape(b, -1, c)
    return flat.gather(1, ind.unsqueeze(2).expand(-1, -1, c))


class RegLoss(nn.Module):
    def forward(self, output, mask, ind, target):
        pred = _gather_at(output, ind)
        mask = mask.float().unsqueeze(2)
        loss = F.l1_loss(pred * mask, target * mask, reduction="none") / (mask.sum() + 1e-4)
        return loss.transpose(2, 0).sum(dim=2).sum(dim=1)


class FastFocalLoss(nn.Module):
    def forward(self, out, target, ind, mask, cat):
        mask = mask.float()
        neg_loss = (torch.log(1 - out) * torch.pow(out, 2) * torch.pow(1 - target, 4)).sum()
        pos_pred = _gather_at(out, ind).gather(2, cat.unsqueeze(2))
        num_pos = mask.sum()
        pos_loss = (torch.log(pos_pred) * torch.pow(1 - pos_pred, 2) * mask.unsqueeze(2)).sum()
        # reference: `if num_pos == 0: return -neg_loss` (a host sync); same value without the branch
        return -(pos_loss + neg_loss) / torch.clamp(num_pos, min=1.0)


class SepHead(nn.Module):
    def __init__(self, in_channels, heads, head_conv=64, final_kernel=1, bn=None, init_bias=-2.19):
        super().__init__()
        self.heads = heads
        for head, (classes, num_conv) in self.heads.items():
            fc = Sequential()
            for _ in range(num_conv - 1):
                fc.add(nn.Conv2d(in_channels, head_conv, kernel_size=final_kernel, stride=1, padding=final_kernel // 2,
                                 bias=True))
                if bn is not None:
                    fc.add(get_norm(bn, head_conv))
                fc.add(nn.ReLU())
            fc.add(nn.Conv2d(head_conv, classes, kernel_size=final_kernel, stride=1, padding=final_kernel // 2, bias=True))
            if "hm" in head:
                fc[-1].bias.data.fill_(init_bias)
            else:
                for m in fc.modules():
                    if isinstance(m, nn.Conv2d):
                        nn.init.kaiming_normal_(m.weight, a=0, mode="fan_out", nonlinearity="relu")
                        nn.init.constant_(m.bias, 0)
            setattr(self, head, fc)

    def forward(self, x):
        return {head: getattr(self, head)(x) for head in self.heads}


class CenterHead(nn.Module):
    def __init__(self, config, init_bias=-2.19, share_conv_channel=64, num_hm_conv=2):
        super().__init__()
        head = config.model.head
        self.class_names = [list(t["class_names"]) for t in head.tasks]
        self.num_classes = [len(n) for n in self.class_names]
        self.code_weights = list(head.misc.code_weights)
        self.weight = head.misc.weight
        self.common_heads = {k: tuple(v) for k, v in head.misc.common_heads.items()}
        self.in_channels = head.in_channels
        self.criterion = FastFocalLoss()
        self.criterion_reg = RegLoss()
        self.box_n_dim = 9 if "vel" in self.common_heads else 7
        self.rotate_nms = None  # set by the detector from its backend; None = the CUDA operator
        norm = config.model.neck.norm
        self.shared_conv = nn.Sequential(
            nn.Conv2d(self.in_channels, share_conv_channel, kernel_size=3, padding=1, bias=True),
            get_norm(norm, share_conv_channel), nn.ReLU(inplace=True))
        self.tasks = nn.ModuleList()
        for num_cls in self.num_classes:
            heads = copy.deepcopy(self.common_heads)
            heads.update(dict(hm=(num_cls, num_hm_conv)))
            self.tasks.append(SepHead(share_conv_channel, heads, bn=norm, init_bias=init_bias, final_kernel=3))

    def forward(self, x):
        x = self.shared_conv(x)
        return [task(x) for task in self.tasks]

    def _code_weights_on(self, like):
        """code_weights as a tensor on `like`'s device, built once (a per-call new_tensor is a pageable host-to-device
        copy: a synchronisation, and illegal under CUDA-graph capture)."""
        cached = getattr(self, "_code_weights_cache", None)
        if cached is None or cached.device != like.device or cached.dtype != like.dtype:
            cached = torch.tensor(list(self.code_weights), dtype=like.dtype, device=like.device)
            self._code_weights_cache = cached
        return cached

    def loss(self, example, preds_dicts):
        out = {}
        for task_id, preds in enumerate(preds_dicts):
            hm = torch.clamp(preds["hm"].sigmoid(), min=1e-4, max=1 - 1e-4)
            hm_loss = self.criterion(hm, example["hm"][task_id], example["ind"][task_id], example["mask"][task_id],
                                     example["cat"][task_id])
            target_box = example["anno_box"][task_id]
            if "vel" in preds:
                anno = torch.cat((preds["reg"], preds["height"], preds["dim"], preds["vel"], preds["rot"]), dim=1)
            else:
                anno = torch.cat((preds["reg"], preds["height"], preds["dim"], preds["rot"]), dim=1)
                target_box = torch.cat((target_box[..., :6], target_box[..., 8:]), dim=-1)  # drop the velocity target
            box_loss = self.criterion_reg(anno, example["mask"][task_id], example["ind"][task_id], target_box)
            loc_loss = (box_loss * self._code_weights_on(box_loss)[:box_loss.shape[0]]).sum()
            out["%d_loss" % task_id] = hm_loss + self.weight * loc_loss
            out["%d_hm_loss" % task_id] = hm_loss.detach()
            out["%d_loc_loss" % task_id] = loc_loss
            out["%d_num_positive" % task_id] = example["mask"][task_id].float().sum()
        return out

    @torch.no_grad()
    def decode(self, preds_dicts, post_cfg):
        """Heatmap -> boxes (x,y,z,l,w,h,[vx,vy],yaw); per task and scene: score threshold + centre range mask, rotated
        NMS (pre / post max size), then the tasks are concatenated with class offsets (CP/center_head.py:173-376).
        The NMS is csrc/iou3d.cu behind the reference's `rotate_nms_pcdet` convention (CP/box_torch_ops.py:239-264)."""
        rotate_nms_pcdet = self.rotate_nms
        if rotate_nms_pcdet is None:
            from ...operators.iou3d_nms import rotate_nms_pcdet

        pc_range, voxel, osf = post_cfg.pc_range, post_cfg.voxel_size, post_cfg.out_size_factor
        per_task = []
        for task_id, preds in enumerate(preds_dicts):
            hm = preds["hm"].sigmoid()
            b, c, h, w = hm.shape
            ys, xs = torch.meshgrid(torch.arange(h, device=hm.device), torch.arange(w, device=hm.device), indexing="ij")
            xs = (xs[None] + preds["reg"][:, 0]) * osf * voxel[0] + pc_range[0]
            ys = (ys[None] + preds["reg"][:, 1]) * osf * voxel[1] + pc_range[1]
            rot = torch.atan2(preds["rot"][:, 0], preds["rot"][:, 1])
            parts = [xs, ys, preds["height"][:, 0], *torch.exp(preds["dim"]).unbind(1)]
            if "vel" in preds:
                parts += list(preds["vel"].unbind(1))
            boxes = torch.stack(parts + [rot], dim=-1).reshape(b, h * w, -1)
            scores, labels = hm.permute(0, 2, 3, 1).reshape(b, h * w, c).max(dim=-1)
            limit = boxes.new_tensor(post_cfg.post_center_limit_range)
            scenes = []
            for bi in range(b):
                bx, sc, lb = boxes[bi], scores[bi], labels[bi]
                mask = (sc > post_cfg.score_threshold) & (bx[:, :3] >= limit[:3]).all(1) & (bx[:, :3] <= limit[3:]).all(1)
                bx, sc, lb = bx[mask], sc[mask], lb[mask]
                sel = rotate_nms_pcdet(bx[:, [0, 1, 2, 3, 4, 5, -1]].float(), sc.float(), thresh=post_cfg.nms.nms_iou_threshold,
                                       pre_maxsize=post_cfg.nms.nms_pre_max_size, post_max_size=post_cfg.nms.nms_post_max_size)
                scenes.append((bx[sel], sc[sel], lb[sel] + sum(self.num_classes[:task_id])))
            per_task.append(scenes)
        out = []
        for bi in range(len(per_task[0])):
            out.append({"boxes3d": torch.cat([t[bi][0] for t in per_task]).cpu(),
                        "scores": torch.cat([t[bi][1] for t in per_task]).cpu(),
                        "labels": (torch.cat([t[bi][2] for t in per_task]) + 1).cpu()})
        return out
