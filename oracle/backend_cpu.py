"""ORACLE (test infrastructure) — a backend object with the same two operator families as
efg_b200.backend.cuda_backend(), implemented by the CPU restatements.  It lets tests and bench.py's
CPU-baseline leg run the unmodified model wiring on the host.  Never imported by efg_b200."""
from . import box_attn as _box
from . import spconv_cpu


class CpuOracleBackend:
    name = "oracle-cpu"
    spconv = spconv_cpu

    @staticmethod
    def box_attn(value, spatial_shapes, level_start_index, sampling_locations, attention_weights, im2col_step):
        return _box.forward(value, spatial_shapes, sampling_locations, attention_weights)

    @staticmethod
    def rotate_nms(boxes, scores, thresh, pre_maxsize=None, post_max_size=None):
        """CP/box_torch_ops.py:239-264 over the CPU restatement of the rotated NMS (oracle/iou3d.c)."""
        import math

        import torch

        from . import iou3d

        b = boxes[:, [0, 1, 2, 4, 3, 5, -1]].clone()
        b[:, -1] = -b[:, -1] - math.pi / 2
        order = scores.sort(0, descending=True)[1]
        if pre_maxsize is not None:
            order = order[:pre_maxsize]
        if b.shape[0] == 0:
            return order[:0]
        keep = torch.from_numpy(iou3d.nms(b[order].detach().cpu().numpy(), thresh)).to(order.device)
        sel = order[keep]
        return sel[:post_max_size] if post_max_size is not None else sel

    def __deepcopy__(self, memo):
        return self


def cpu_backend():
    return CpuOracleBackend()


def voxelized_sample(points, spec_or_cfg, max_voxels=None):
    """What the reference's Voxelization processor emits for one scene (extend_3d.py:267-282),
    produced by the oracle voxelizer."""
    import numpy as np

    from . import voxelize as ov

    vs, rg = list(spec_or_cfg["voxel_size"]), list(spec_or_cfg["pc_range"])
    mp = spec_or_cfg.get("max_points_in_voxel", 5)
    mv = max_voxels if max_voxels is not None else spec_or_cfg.get("max_voxel_num", 150000)
    v, c, n = ov.hard_voxelize(points, vs, rg, mp, mv)
    grid = ov.grid_size(vs, rg).astype(np.int64)
    return {"voxels": v, "points": points, "coordinates": c, "num_points_per_voxel": n,
            "num_voxels": np.array([v.shape[0]], dtype=np.int64), "shape": grid,
            "range": np.asarray(rg, dtype=np.float32), "size": np.asarray(vs, dtype=np.float32)}
