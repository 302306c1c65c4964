"""Shared by make_golden_model.py (runs the REFERENCE) and tests/test_model_golden.py (runs this repo's port):
the small model configurations, the deterministic weight fill keyed by state_dict names, and batch construction.

The weight fill depends only on a parameter's NAME and SHAPE, so the 18 M-parameter state_dict never has to be
stored: both sides regenerate it, and a port whose state_dict keys or shapes differ from the reference's cannot load it.
"""
import copy
import zlib

import numpy as np
import torch

SMALL_RANGE = [-12.8, -12.8, -2.0, 12.8, 12.8, 4.0]
VOXEL = [0.1, 0.1, 0.15]


def fill_state_dict(sd):
    """name/shape -> values: N(0, 1/fan_in) for matrices and conv kernels, 1 + 0.1 N for norm weights, 0.1 N for
    biases, U(0.5, 1.5) for running_var; integer buffers are kept."""
    out = {}
    for k, v in sd.items():
        g = torch.Generator().manual_seed(zlib.crc32(k.encode()))
        if not v.dtype.is_floating_point:
            out[k] = v.clone()
        elif k.endswith("running_var"):
            out[k] = (torch.rand(v.shape, generator=g) + 0.5).to(v.dtype)
        elif v.dim() <= 1:
            out[k] = (torch.randn(v.shape, generator=g) * 0.1 + (1.0 if k.endswith("weight") else 0.0)).to(v.dtype)
        else:
            out[k] = (torch.randn(v.shape, generator=g) / max(v[0].numel(), 1) ** 0.5).to(v.dtype)
    return out


def model_overrides(kind):
    ds = {"pc_range": SMALL_RANGE, "voxel_size": VOXEL, "max_voxel_num": 20000}
    if kind in ("voxel_detr", "conquer"):
        return dict(dataset=ds, model={"device": "cpu", "transformer": {"num_queries": 100, "enc_layers": 1, "dec_layers": 2}})
    return dict(dataset=ds, model={"device": "cpu"})


def make_config(kind, device="cpu"):
    from efg_b200.config import centerpoint_config, conquer_config, voxel_detr_config

    ov = model_overrides(kind)
    ov["model"]["device"] = device
    return {"voxel_detr": voxel_detr_config, "conquer": conquer_config, "centerpoint": centerpoint_config}[kind](**ov)


def make_scenes(kind):
    """Seeded synthetic scenes (points [N,5] f32 + the reference's annotation dict) for one golden case."""
    from efg_b200.data import SceneSpec, make_scene

    spec = SceneSpec(pc_range=SMALL_RANGE, voxel_size=VOXEL)
    n_pts, seed = {"voxel_detr": (4000, 0), "conquer": (4000, 0), "centerpoint": (5000, 2)}[kind]
    out = []
    for i in range(2):
        pts, ann = make_scene(n_pts, spec, seed=seed + i, num_objects=6)
        keep = (np.abs(ann["gt_boxes"][:, 0]) < 12) & (np.abs(ann["gt_boxes"][:, 1]) < 12)
        out.append((pts, {k: v[keep] for k, v in ann.items()}))
    return out


def make_batch(scenes, dataset_cfg):
    """`batched_inputs` of the reference's forward (VD/voxel_detr.py:93-100): fresh copies, because the reference's
    box coder mutates the annotation arrays in place (VD/modules/box_coder.py:50-70)."""
    from oracle.backend_cpu import voxelized_sample

    return [(voxelized_sample(p.copy(), dataset_cfg), {"annotations": {k: copy.deepcopy(v) for k, v in a.items()}})
            for p, a in scenes]


# parameters whose full gradients are stored (small tensors from every part of the graph)
GRAD_KEYS = {
    "voxel_detr": ["backbone.extractor.bottom_up.stem.conv1.0.weight", "input_proj.0.1.weight",
                   "transformer.encoder.layers.0.self_attn.linear_box_bias", "transformer.decoder.layers.1.norm3.weight",
                   "transformer.decoder.detection_head.class_embed.1.layers.2.bias",
                   "transformer.proposal_head.bbox_embed.0.layers.2.bias"],
    "conquer": ["backbone.extractor.bottom_up.stem.conv1.0.weight", "projector.2.bias", "predictor.0.bias",
                "transformer.decoder.layers.0.norm2.weight", "transformer.decoder.detection_head.class_embed.0.layers.2.bias"],
    "centerpoint": ["backbone.conv_input.0.weight", "neck.blocks.0.2.weight", "center_head.shared_conv.1.bias",
                    "center_head.tasks.0.hm.3.bias", "center_head.tasks.0.dim.3.bias"],
}
