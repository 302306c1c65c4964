"""Normalised box code of Voxel-DETR (VD/modules/box_coder.py:34-80).

encode: (x,y) -> [0,1] over the point-cloud range, z over [-10,10], (l,w) over the range size,
h over 20, heading -> (limit_period(theta) + pi) / 2pi; labels become 0-based.  Unlike the
reference this does not mutate its input."""
import math

import torch

from .box_utils import limit_period


class VoxelBoxCoder3D:
    def __init__(self, voxel_size, pc_range, n_dim=7, device=torch.device("cpu")):
        self.device = device
        self.voxel_size = torch.tensor(list(voxel_size), dtype=torch.float32, device=device)
        self.pc_range = torch.tensor(list(pc_range), dtype=torch.float32, device=device)
        self.pc_size = self.pc_range[3:] - self.pc_range[:3]
        self.z_normalizer = 10.0
        self.n_dim = n_dim

    @property
    def code_size(self):
        return self.n_dim

    def encode(self, target):
        out = dict(target)
        out["labels"] = target["labels"] - 1
        b = target["gt_boxes"].to(torch.float32)
        xy = (b[:, :2] - self.pc_range[:2]) / self.pc_size[:2]
        z = (b[:, 2:3] + self.z_normalizer) / (2 * self.z_normalizer)
        lw = b[:, 3:5] / self.pc_size[:2]
        h = b[:, 5:6] / (2 * self.z_normalizer)
        theta = limit_period(b[:, -1:], offset=0.5, period=math.pi * 2)
        theta = (theta + 0.5 * math.pi * 2) / (math.pi * 2)
        code = torch.cat([xy, z, lw, h, theta], dim=1)
        if code.numel():
            assert bool(((code >= 0) & (code <= 1)).all()), "ground-truth box outside the normalised range"
        out["gt_boxes"] = code
        return out

    def decode(self, pred):
        out = pred.clone()
        out[..., :2] = pred[..., :2] * self.pc_size[:2] + self.pc_range[:2]
        out[..., 2] = pred[..., 2] * 2 * self.z_normalizer - self.z_normalizer
        out[..., 3:5] = pred[..., 3:5] * self.pc_size[:2]
        out[..., 5] = pred[..., 5] * 2 * self.z_normalizer
        out[..., -1] = pred[..., -1] * math.pi * 2 - math.pi
        return out
