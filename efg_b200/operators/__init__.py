"""The slice of ``efg.operators`` that lies on the 3D-detection hot path
(efg/operators/__init__.py:1-5)."""
from .box_attention_func import BoxAttnFunction
from .scatter_points import DynamicScatter, dynamic_scatter
from .voxelize import Voxelization, voxelization

__all__ = ["BoxAttnFunction", "DynamicScatter", "dynamic_scatter", "Voxelization", "voxelization"]
