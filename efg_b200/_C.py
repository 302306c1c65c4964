"""The callables of the reference's pybind module ``efg._C`` that lie on the hot path
(efg/operators/src/vision.cpp:70-122), re-implemented as thin Python over the C ABI.

Same names, argument order, return values and error behaviour (RuntimeError for CPU tensors or
non-contiguous inputs, as ``CHECK_INPUT`` / ``AT_ERROR`` produce in the reference).
"""
import ctypes

import torch

from . import _lib, ops


def hard_voxelize(points, voxels, coors, num_points_per_voxel, voxel_size, coors_range, max_points, max_voxels,
                  NDim=3):
    """voxelization.h:51-69.  Fills the caller's preallocated outputs, returns the voxel count."""
    if NDim != 3:
        raise RuntimeError("hard_voxelize: only NDim=3 is supported")
    if not points.is_cuda:
        raise RuntimeError("Not compiled with CPU support: efg_b200.hard_voxelize needs CUDA tensors")
    ops._check(points, "points", torch.float32)
    ops._check(voxels, "voxels", torch.float32)
    ops._check(coors, "coors", torch.int32)
    ops._check(num_points_per_voxel, "num_points_per_voxel", torch.int32)
    n, f = points.shape
    if voxels.shape[0] < min(n, max_voxels) or voxels.shape[1] != max_points or voxels.shape[2] != f:
        raise RuntimeError("hard_voxelize: voxels buffer %r does not match (max_voxels=%d, max_points=%d, F=%d)" %
                           (tuple(voxels.shape), max_voxels, max_points, f))
    dev = points.device
    L = _lib.lib()
    offsets = torch.tensor([0, n], dtype=torch.int32, device=dev)
    counts = torch.empty((2,), dtype=torch.int32, device=dev)
    ws = ops.workspace(L.efgb_voxelize_workspace_bytes(n, 1), dev)
    rc = L.efgb_hard_voxelize(ops._p(points), n, f, ops._p(offsets), 1, _lib.f32array(voxel_size),
                              _lib.f32array(coors_range), int(max_points), int(max_voxels), ops._p(voxels),
                              ops._p(coors), 3, ops._p(num_points_per_voxel), ctypes.c_void_p(0), ops._p(counts),
                              ops._p(ws), ws.numel(), ops._stream())
    _lib.check(rc, "hard_voxelize")
    return int(counts[1].item())


def dynamic_voxelize(points, coors, voxel_size, coors_range, NDim=3):
    """voxelization.h:71-83."""
    if NDim != 3:
        raise RuntimeError("dynamic_voxelize: only NDim=3 is supported")
    ops.dynamic_voxelize(points, coors, voxel_size, coors_range)


def dynamic_point_to_voxel_forward(feats, coors, reduce_type):
    """voxelization.h:96-108 -> [voxel_feats, voxel_coors, point2voxel_map, voxel_points_count]."""
    if not feats.is_cuda:
        raise RuntimeError("dynamic_point_to_voxel_forward: do not support cpu yet")
    return list(ops.dynamic_scatter_forward(feats, coors, reduce_type))


def dynamic_point_to_voxel_backward(grad_feats, grad_reduced_feats, feats, reduced_feats, coors_idx, reduce_count,
                                    reduce_type):
    """voxelization.h:110-128.  Fills grad_feats in place."""
    if not grad_feats.is_cuda:
        raise RuntimeError("dynamic_point_to_voxel_backward: do not support cpu yet")
    ops.dynamic_scatter_backward(grad_feats, grad_reduced_feats, feats, reduced_feats, coors_idx, reduce_count,
                                 reduce_type)


def _check_im2col(batch, im2col_step):
    step = min(batch, int(im2col_step))
    if step <= 0 or batch % step != 0:
        # box_attn.cu:41 AT_ASSERTM(batch % im2col_step_ == 0, ...)
        raise RuntimeError("batch(%d) must divide im2col_step(%d)" % (batch, step))


def box_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    """box_attn.h:29-54.  The im2col chunking of the reference is a launch detail; the whole
    batch runs in one launch here, but the divisibility contract is kept."""
    if not value.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    _check_im2col(value.shape[0], im2col_step)
    return ops.box_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)


def box_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output, im2col_step):
    """box_attn.h:56-83 -> [grad_value, grad_sampling_loc, grad_attn_weight]."""
    if not value.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    _check_im2col(value.shape[0], im2col_step)
    return list(ops.box_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                      grad_output))


# ---- BEV IoU / rotated NMS (vision.cpp:100-104, iou3d_nms.cpp:22-178) ---------------------------------------------
def _check_boxes(t, name):
    # CHECK_INPUT of the reference: CUDA + contiguous (iou3d_nms.cpp:11-13)
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDAtensor " % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous " % name)


def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    """iou3d_nms.cpp:22-40: fills ans_overlap [N, M] with the BEV overlap areas; returns 1."""
    for t, n in ((boxes_a, "boxes_a"), (boxes_b, "boxes_b"), (ans_overlap, "ans_overlap")):
        _check_boxes(t, n)
    ans_overlap.copy_(ops.boxes_bev(boxes_a.float(), boxes_b.float(), overlap=True))
    return 1


def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou):
    """iou3d_nms.cpp:42-58: fills ans_iou [N, M] with the BEV IoU of rotated boxes; returns 1."""
    for t, n in ((boxes_a, "boxes_a"), (boxes_b, "boxes_b"), (ans_iou, "ans_iou")):
        _check_boxes(t, n)
    ans_iou.copy_(ops.boxes_bev(boxes_a.float(), boxes_b.float(), overlap=False))
    return 1


def _nms(boxes, keep, thresh, normal):
    _check_boxes(boxes, "boxes")
    if not keep.is_contiguous():
        raise RuntimeError("keep must be contiguous ")
    kept, count = ops.nms_bev(boxes, thresh, normal=normal)
    n = int(count.item())                    # the reference's signature returns a host int
    keep[:n].copy_(kept[:n])                 # `keep` is a host LongTensor in the reference (iou3d_nms.py:113)
    return n


def nms_gpu(boxes, keep, nms_overlap_thresh):
    """iou3d_nms.cpp:60-121: boxes [N,7] sorted by score; writes the kept indices into `keep`, returns their number."""
    return _nms(boxes, keep, nms_overlap_thresh, False)


def nms_normal_gpu(boxes, keep, nms_overlap_thresh):
    """iou3d_nms.cpp:124-176: the same with axis-aligned IoU (heading ignored)."""
    return _nms(boxes, keep, nms_overlap_thresh, True)
