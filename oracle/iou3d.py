"""ORACLE (test infrastructure) — BEV IoU of rotated boxes and rotated NMS on the CPU: ctypes front end of iou3d.c
(restates efg/operators/src/iou3d_nms/iou3d_cpu.cpp:61-214, iou3d_nms_kernel.cu:331-342, iou3d_nms.cpp:87-121).
Pinned by tests/golden/iou3d_*.npz, which hold the output of the reference's own iou3d_cpu.cpp."""
import ctypes

import numpy as np

from . import voxelize as _vox


def _lib():
    L = _vox._load()
    if not getattr(L, "_iou3d_bound", False):
        L.oracle_boxes_bev.restype = None
        L.oracle_boxes_bev.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
        L.oracle_nms.restype = ctypes.c_int64
        L.oracle_nms.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
        L._iou3d_bound = True
    return L


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def boxes_iou_bev(boxes_a, boxes_b, overlap=False):
    """[N,7] x [M,7] (x, y, z, dx, dy, dz, heading) -> [N,M] BEV IoU (or overlap area)."""
    a, b = _f32(boxes_a), _f32(boxes_b)
    out = np.zeros((a.shape[0], b.shape[0]), dtype=np.float32)
    _lib().oracle_boxes_bev(a.ctypes.data, a.shape[0], b.ctypes.data, b.shape[0], 1 if overlap else 0, out.ctypes.data)
    return out


def nms(boxes_sorted, thresh, normal=False):
    """Greedy NMS over boxes sorted by descending score -> kept indices (int64)."""
    b = _f32(boxes_sorted)
    keep = np.zeros((b.shape[0],), dtype=np.int64)
    n = _lib().oracle_nms(b.ctypes.data, b.shape[0], float(thresh), 1 if normal else 0, keep.ctypes.data)
    return keep[:n]
