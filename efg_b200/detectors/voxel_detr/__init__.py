from .model import VoxelDETR, build_model

__all__ = ["VoxelDETR", "build_model"]
