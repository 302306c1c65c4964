from .model import VoxelNet, build_model

__all__ = ["VoxelNet", "build_model"]
