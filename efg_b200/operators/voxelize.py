"""``Voxelization`` / ``voxelization`` — mirror of efg/operators/voxelize.py:9-96 over the
hash-and-scatter CUDA voxelizer (csrc/voxelize.cu)."""
import torch
from torch import nn
from torch.autograd import Function
from torch.nn.modules.utils import _pair

from .. import _C


class _Voxelization(Function):
    """Non-differentiable. max_points == -1 or max_voxels == -1 selects dynamic voxelization
    (per-point (z,y,x), -1 for dropped points), otherwise first-come hard voxelization."""

    @staticmethod
    def forward(ctx, points, voxel_size, coors_range, max_points=35, max_voxels=20000):
        if max_points == -1 or max_voxels == -1:
            coors = points.new_zeros(size=(points.size(0), 3), dtype=torch.int)
            _C.dynamic_voxelize(points, coors, voxel_size, coors_range, 3)
            return coors
        # the kernel writes every row it reports (zero padding included), so empty() is enough
        voxels = points.new_empty(size=(max_voxels, max_points, points.size(1)))
        coors = points.new_empty(size=(max_voxels, 3), dtype=torch.int)
        num_points_per_voxel = points.new_empty(size=(max_voxels,), dtype=torch.int)
        voxel_num = _C.hard_voxelize(points, voxels, coors, num_points_per_voxel, voxel_size, coors_range, max_points,
                                     max_voxels, 3)
        return voxels[:voxel_num], coors[:voxel_num], num_points_per_voxel[:voxel_num]


voxelization = _Voxelization.apply


class Voxelization(nn.Module):
    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000):
        """
        Args:
            voxel_size (list): [x, y, z] voxel edge lengths
            point_cloud_range (list): [x_min, y_min, z_min, x_max, y_max, z_max]
            max_num_points (int): max points kept per voxel
            max_voxels (tuple or int): max voxels at (training, testing) time
        """
        super().__init__()
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.max_num_points = max_num_points
        self.max_voxels = max_voxels if isinstance(max_voxels, tuple) else _pair(max_voxels)

        pc_range = torch.tensor(point_cloud_range, dtype=torch.float32)
        vsize = torch.tensor(voxel_size, dtype=torch.float32)
        grid_size = torch.round((pc_range[3:] - pc_range[:3]) / vsize).long()
        self.grid_size = grid_size
        # [w, h, d] -> [d, h, w] with d collapsed, as the reference exposes it
        self.pcd_shape = [*grid_size[:2], 1][::-1]

    def forward(self, input):
        max_voxels = self.max_voxels[0] if self.training else self.max_voxels[1]
        return voxelization(input, self.voxel_size, self.point_cloud_range, self.max_num_points, max_voxels)

    def __repr__(self):
        return "%s(voxel_size=%s, point_cloud_range=%s, max_num_points=%s, max_voxels=%s)" % (
            self.__class__.__name__, self.voxel_size, self.point_cloud_range, self.max_num_points, self.max_voxels)
