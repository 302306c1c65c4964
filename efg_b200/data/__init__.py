from .synthetic import SceneSpec, WAYMO, NUSCENES, make_batch, make_scene

__all__ = ["SceneSpec", "WAYMO", "NUSCENES", "make_batch", "make_scene"]
