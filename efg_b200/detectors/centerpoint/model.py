"""CenterPoint ``VoxelNet`` (CP/voxelnet.py:19-226): mean-VFE reader -> SpMiddleResNetFHD sparse encoder ->
RPN neck -> CenterHead; training returns the per-task loss dict.  Inputs as for Voxel-DETR: the
reference's CPU-voxelized sample dicts, or raw ``points`` that are voxelized on the GPU."""
import numpy as np
import torch
from torch import nn

from ... import ops
from ...backend import cuda_backend
from ...modeling.rpn import RPN
from ...modeling.sparse_backbone import SpMiddleResNetFHD
from ...modeling.voxel_reader import VoxelMeanFeatureExtractor
from ..voxel_detr.model import collate_voxels
from .assign import assign_scene
from .head import CenterHead


class VoxelNet(nn.Module):
    def __init__(self, config, backend=None):
        super().__init__()
        self.config = config
        self.backend = [backend or cuda_backend()]
        self.device = torch.device(config.model.device)
        self.reader = VoxelMeanFeatureExtractor(**config.model.reader)
        self.backbone = SpMiddleResNetFHD(**config.model.backbone, backend=self.backend[0])
        self.neck = RPN(config.model.neck)
        self.center_head = CenterHead(config)
        self.center_head.rotate_nms = getattr(self.backend[0], "rotate_nms", None)
        a = config.model.loss
        self.out_size_factor = a.out_size_factor
        self.tasks = config.model.head.tasks
        self.gaussian_overlap, self._max_objs, self._min_radius = a.gaussian_overlap, a.max_objs, a.min_radius
        pr = np.asarray(config.dataset.pc_range, dtype=np.float32)
        vs = np.asarray(config.dataset.voxel_size, dtype=np.float32)
        self.grid_size = np.round((pr[3:] - pr[:3]) / vs).astype(np.int64)
        self.to(self.device)

    def voxelize_on_device(self, samples):
        ds = self.config.dataset
        pts = [s["points"] if isinstance(s["points"], torch.Tensor) else
               torch.from_numpy(np.ascontiguousarray(s["points"], dtype=np.float32)) for s in samples]
        sizes = [p.shape[0] for p in pts]
        points = torch.cat([p.to(self.device, non_blocking=True) for p in pts], 0)
        offs = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32)
        if self.device.type == "cuda":
            offs = offs.pin_memory()
        offs = offs.to(self.device, non_blocking=True)
        r = ops.hard_voxelize_batched(points.contiguous(), offs, ds.voxel_size, ds.pc_range,
                                      ds.get("max_points_in_voxel", 5), ds.get("max_voxel_num", 150000), coors_dim=4,
                                      want_voxels=False, want_mean=True)
        m = int(r["counts"][-1].item())
        return r["mean"][:m], r["coors"][:m], r["num_points_per_voxel"][:m], self.grid_size

    def label_assign(self, infos):
        ds = self.config.dataset
        if self.device.type == "cuda" and getattr(self.backend[0], "name", "") == "efgb200-cuda":
            from .assign import assign_batch_device

            return assign_batch_device(infos, self.tasks, self.grid_size, ds.pc_range, ds.voxel_size, self.out_size_factor,
                                       self.gaussian_overlap, self._max_objs, self._min_radius, self.device)
        per_scene = [assign_scene(info["annotations"], self.tasks, self.grid_size, ds.pc_range, ds.voxel_size,
                                  self.out_size_factor, self.gaussian_overlap, self._max_objs, self._min_radius)
                     for info in infos]
        targets = {}
        for key in ("hm", "anno_box", "ind", "mask", "cat"):
            targets[key] = []
            for t in range(len(self.tasks)):
                arr = torch.from_numpy(np.stack([s[key][t] for s in per_scene], axis=0))
                if self.device.type == "cuda":
                    arr = arr.pin_memory()
                targets[key].append(arr.to(self.device, non_blocking=True))
        return targets

    def forward(self, batched_inputs):
        if self.training and self.device.type == "cuda":
            ops.refresh_packs()   # every weight image the optimizer step made stale, in one launch
        samples = [bi[0] for bi in batched_inputs]
        infos = [bi[1] for bi in batched_inputs]
        batch_size = len(samples)
        with torch.no_grad():
            if "voxels" in samples[0]:
                voxels, coords, npv, input_shape = collate_voxels(samples, self.device)
            else:
                voxels, coords, npv, input_shape = self.voxelize_on_device(samples)
        x = self.reader(voxels, npv)
        x = self.backbone(x, coords, batch_size, input_shape)
        x = self.neck(x)
        preds = self.center_head(x)
        if self.training:
            with torch.no_grad():
                targets = self.label_assign(infos)
            return self.center_head.loss(targets, preds)
        return self.center_head.decode(preds, self.config.model.post_process)


def build_model(self, config, backend=None):
    return VoxelNet(config, backend=backend)
