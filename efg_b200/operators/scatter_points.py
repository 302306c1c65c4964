"""``dynamic_scatter`` / ``DynamicScatter`` — mirror of efg/operators/scatter_points.py:8-96."""
import torch
from torch import nn
from torch.autograd import Function

from .. import _C


class _dynamic_scatter(Function):
    @staticmethod
    def forward(ctx, feats, coors, reduce_type="max"):
        """feats [N,C], coors [N,3] int -> (voxel_feats [M,C], voxel_coors [M,3]); rows that share a
        coordinate are reduced with 'max' | 'sum' | 'mean'; rows with a negative coordinate are dropped."""
        voxel_feats, voxel_coors, point2voxel_map, voxel_points_count = _C.dynamic_point_to_voxel_forward(
            feats, coors, reduce_type)
        ctx.reduce_type = reduce_type
        ctx.save_for_backward(feats, voxel_feats, point2voxel_map, voxel_points_count)
        ctx.mark_non_differentiable(voxel_coors)
        return voxel_feats, voxel_coors

    @staticmethod
    def backward(ctx, grad_voxel_feats, grad_voxel_coors=None):
        feats, voxel_feats, point2voxel_map, voxel_points_count = ctx.saved_tensors
        grad_feats = torch.zeros_like(feats)
        _C.dynamic_point_to_voxel_backward(grad_feats, grad_voxel_feats.contiguous(), feats, voxel_feats,
                                           point2voxel_map, voxel_points_count, ctx.reduce_type)
        return grad_feats, None, None


dynamic_scatter = _dynamic_scatter.apply


class DynamicScatter(nn.Module):
    def __init__(self, voxel_size, point_cloud_range, average_points: bool):
        """Scatter points into voxels (mean if ``average_points`` else max)."""
        super().__init__()
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.average_points = average_points

    def forward_single(self, points, coors):
        reduce = "mean" if self.average_points else "max"
        return dynamic_scatter(points.contiguous(), coors.contiguous(), reduce)

    def forward(self, points, coors):
        if coors.size(-1) == 3:
            return self.forward_single(points, coors)
        batch_size = int(coors[-1, 0]) + 1
        voxels, voxel_coors = [], []
        for i in range(batch_size):
            inds = torch.where(coors[:, 0] == i)
            voxel, voxel_coor = self.forward_single(points[inds], coors[inds][:, 1:])
            voxel_coors.append(nn.functional.pad(voxel_coor, (1, 0), mode="constant", value=i))
            voxels.append(voxel)
        return torch.cat(voxels, dim=0), torch.cat(voxel_coors, dim=0)

    def __repr__(self):
        return "%s(voxel_size=%s, point_cloud_range=%s, average_points=%s)" % (
            self.__class__.__name__, self.voxel_size, self.point_cloud_range, self.average_points)
