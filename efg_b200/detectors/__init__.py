"""Detectors of the hot path: Voxel-DETR, ConQueR, CenterPoint."""
