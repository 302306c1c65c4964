"""Minimal attribute-style config tree.

The reference reads OmegaConf YAML (efg/config/__init__.py:11-132); omegaconf is not in this
image and the config system is out of scope, so the models here take a plain nested namespace
with the SAME keys as the playground YAMLs (``config.model.transformer.num_queries`` ...).
``load_yaml`` reads a playground config.yaml, resolving ``${a.b.c}`` references to other keys
of the same file; ``includes:`` and ``${oc.env:...}`` entries (dataset gallery) are ignored.
"""
import copy
import re


class Config(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return Config({k: copy.deepcopy(v, memo) for k, v in self.items()})


def to_config(obj):
    if isinstance(obj, dict):
        return Config({k: to_config(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return [to_config(v) for v in obj]
    return obj


def merge(base, override):
    out = copy.deepcopy(base)
    for k, v in override.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = merge(out[k], v)
        else:
            out[k] = copy.deepcopy(v)
    return out


_REF = re.compile(r"^\$\{([A-Za-z0-9_.]+)\}$")


def _resolve(node, root, depth=0):
    if depth > 16:
        raise ValueError("config interpolation too deep")
    if isinstance(node, dict):
        return {k: _resolve(v, root, depth) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve(v, root, depth) for v in node]
    if isinstance(node, str):
        m = _REF.match(node.strip())
        if m:
            cur = root
            for part in m.group(1).split("."):
                if not isinstance(cur, dict) or part not in cur:
                    return node  # unresolved (e.g. gallery keys): keep the literal
                cur = cur[part]
            return _resolve(cur, root, depth + 1)
    return node


def load_yaml(path, overrides=None):
    import yaml

    with open(path) as f:
        raw = yaml.safe_load(f)
    raw.pop("includes", None)
    if overrides:
        raw = merge(raw, overrides)
    return to_config(_resolve(raw, raw))


# Keys of playground/detection.3d/waymo/conquer/VoxelDETR.waymo.res18.p3.box_only_with_3cat.bs6.epoch6/config.yaml
# that the model reads (dataset.* :11-20, model.* :66-131), with num_queries at BASELINE.json's 300.
VOXEL_DETR_WAYMO = {
    "dataset": {
        "format": "XYZIT",
        "nsweeps": 1,
        "classes": ["VEHICLE", "PEDESTRIAN", "CYCLIST"],
        "pc_range": [-75.2, -75.2, -2.0, 75.2, 75.2, 4.0],
        "voxel_size": [0.1, 0.1, 0.15],
        "max_points_in_voxel": 5,
        "max_voxel_num": 120000,
    },
    "model": {
        "device": "cuda",
        "hidden_dim": 256,
        "aux_loss": True,
        "loss": {
            "bbox_loss_coef": 4, "giou_loss_coef": 2, "class_loss_coef": 1, "rad_loss_coef": 4,
            "matcher": {"class_weight": 1, "bbox_weight": 4, "giou_weight": 2, "rad_weight": 4},
        },
        "metrics": [{"type": "accuracy", "params": {}}],
        "sparse_resnets": {
            "depth": 18, "out_features": ["res2", "res3", "res4"], "num_groups": 1, "norm": "BN1d",
            "activation": {"type": "ReLU", "inplace": True}, "width_per_group": 64,
            "res1_out_channels": 64, "stem_out_channels": 32,
        },
        "fpn": {"in_features": ["res2", "res3", "res4"], "top_block_in_feature": "p4", "out_channels": 256,
                "norm": "BN", "fuse_type": "sum"},
        "backbone": {"type": "voxelnet", "hidden_dim": 256, "position_encoding": "sine", "out_features": ["p3"],
                     "reader": {"norm": "BN"}, "out_channels": 256},
        "transformer": {"hidden_dim": 256, "nhead": 8, "enc_layers": 3, "dec_layers": 3, "dim_feedforward": 1024,
                        "dropout": 0, "num_queries": 300},
    },
}


def voxel_detr_config(**overrides):
    cfg = copy.deepcopy(VOXEL_DETR_WAYMO)
    cfg["model"]["backbone"]["extractor"] = {"resnet": cfg["model"]["sparse_resnets"], "fpn": cfg["model"]["fpn"]}
    if overrides:
        cfg = merge(cfg, overrides)
    return to_config(cfg)


def conquer_config(**overrides):
    """ConQueR = Voxel-DETR + contrastive / denoising keys (CQ/config.yaml:133-144)."""
    extra = {"model": {"contrastive": {"mom": 0.999, "dim": 256, "eqco": 1000, "tau": 0.7, "loss_coeff": 0.2},
                       "dn": {"enabled": True, "dn_number": 3, "dn_box_noise_scale": 0.4, "dn_label_noise_ratio": 0.5}}}
    cfg = merge(voxel_detr_config(), extra)
    if overrides:
        cfg = merge(cfg, overrides)
    return to_config(cfg)


# playground/detection.3d/waymo/center_point/centerpoint.waymo.voxelnet...bs48.36e/config.yaml:60-115
CENTERPOINT_WAYMO = {
    "dataset": {"classes": ["VEHICLE", "PEDESTRIAN", "CYCLIST"], "pc_range": [-75.2, -75.2, -2.0, 75.2, 75.2, 4.0],
                "voxel_size": [0.1, 0.1, 0.15], "max_points_in_voxel": 5, "max_voxel_num": 150000},
    "model": {
        "device": "cuda",
        "reader": {"num_input_features": 5, "norm": "BN"},
        "backbone": {"num_input_features": 5, "norm": "BN1d"},
        "neck": {"num_input_features": 256, "layer_nums": [5, 5], "ds_layer_strides": [1, 2], "ds_num_filters": [128, 256],
                 "us_layer_strides": [1, 2], "us_num_filters": [256, 256], "norm": "BN"},
        "head": {"in_channels": 512, "norm": {"type": "BN"},
                 "tasks": [{"num_classes": 3, "class_names": ["VEHICLE", "PEDESTRIAN", "CYCLIST"]}],
                 "misc": {"dataset": "waymo", "weight": 2, "code_weights": [1.0] * 8,
                          "common_heads": {"reg": [2, 2], "height": [1, 2], "dim": [3, 2], "rot": [2, 2]}}},
        "loss": {"out_size_factor": 8, "dense_reg": 1, "gaussian_overlap": 0.1, "max_objs": 500, "min_radius": 2},
        "post_process": {"post_center_limit_range": [-80, -80, -10.0, 80, 80, 10.0],
                         "nms": {"nms_pre_max_size": 4096, "nms_post_max_size": 300, "nms_iou_threshold": 0.7},
                         "score_threshold": 0.1, "pc_range": [-75.2, -75.2, -2.0, 75.2, 75.2, 4.0], "out_size_factor": 8,
                         "voxel_size": [0.1, 0.1, 0.15]},
    },
}


def centerpoint_config(**overrides):
    cfg = copy.deepcopy(CENTERPOINT_WAYMO)
    if overrides:
        cfg = merge(cfg, overrides)
    cfg["model"]["post_process"]["pc_range"] = cfg["dataset"]["pc_range"]
    cfg["model"]["post_process"]["voxel_size"] = cfg["dataset"]["voxel_size"]
    return to_config(cfg)
