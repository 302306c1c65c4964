"""``VoxelMeanFeatureExtractor`` (efg/modeling/readers/voxel_reader.py:8-19).

Mean of the (<= max_points) points of every voxel.  When the voxels come from the fused CUDA
voxelizer the mean is already computed there (``mean`` output of efgb_hard_voxelize) and is passed
straight through; otherwise it is the reference expression."""
from torch import nn


class VoxelMeanFeatureExtractor(nn.Module):
    def __init__(self, num_input_features, norm="BN1d"):
        super().__init__()
        self.num_input_features = num_input_features

    def forward(self, features, num_voxels, coors=None):
        if features.dim() == 2:  # already reduced by the fused voxelizer: [M, C]
            return features[:, : self.num_input_features].contiguous()
        summed = features[:, :, : self.num_input_features].sum(dim=1, keepdim=False)
        return (summed / num_voxels.type_as(features).view(-1, 1)).contiguous()
