"""FPN neck over the sparse backbone's BEV maps (efg/modeling/backbones/fpn.py:18-224).

Dense 2-D convolutions stay torch/cuDNN library calls.  The reference
evaluates every level even though Voxel-DETR consumes only p3 (VD/config.yaml:116), which is why
it needs ``ddp.find_unused_parameters``; with ``needed`` set, levels whose outputs nobody reads
are skipped — identical outputs, parameters still registered (their grads stay None).
"""
import math

import torch.nn.functional as F
from torch import nn

from .norm import get_norm
from .sparse_backbone import build_sparse_resnet_backbone


class Conv2d(nn.Conv2d):
    """nn.Conv2d with an optional ``norm`` and ``activation`` applied after it (common/blocks.py:45-99)."""

    def __init__(self, *args, norm=None, activation=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.norm = norm
        self.activation = activation

    def forward(self, x):
        x = F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x


def c2_xavier_fill(module):
    nn.init.kaiming_uniform_(module.weight, a=1)
    if module.bias is not None:
        nn.init.constant_(module.bias, 0)


class LastLevelMaxPool(nn.Module):
    def __init__(self, in_feature="p5"):
        super().__init__()
        self.num_levels = 1
        self.in_feature = in_feature

    def forward(self, x):
        return [F.max_pool2d(x, kernel_size=1, stride=2, padding=0)]


class FPN(nn.Module):
    def __init__(self, bottom_up, in_features, out_channels, norm="", top_block=None, fuse_type="sum"):
        super().__init__()
        self.out_channels = out_channels
        shapes = bottom_up.output_shape()
        in_strides = [shapes[f]["stride"] for f in in_features]
        in_channels = [shapes[f]["channels"] for f in in_features]
        for i, s in enumerate(in_strides[1:], 1):
            assert s == 2 * in_strides[i - 1], "Strides {} {} are not log2 contiguous".format(s, in_strides[i - 1])
        use_bias = norm == ""
        laterals, outputs = [], []
        for idx, cin in enumerate(in_channels):
            lateral = Conv2d(cin, out_channels, kernel_size=1, bias=use_bias, norm=get_norm(norm, out_channels))
            output = Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=use_bias,
                            norm=get_norm(norm, out_channels))
            c2_xavier_fill(lateral)
            c2_xavier_fill(output)
            stage = int(math.log2(in_strides[idx]))
            self.add_module("fpn_lateral{}".format(stage), lateral)
            self.add_module("fpn_output{}".format(stage), output)
            laterals.append(lateral)
            outputs.append(output)
        self._stages = [int(math.log2(s)) for s in in_strides]
        self.lateral_convs = laterals[::-1]  # top-down order
        self.output_convs = outputs[::-1]
        self.top_block = top_block
        self.in_features = list(in_features)
        self.bottom_up = bottom_up
        self._out_feature_strides = {"p{}".format(int(math.log2(s))): s for s in in_strides}
        if self.top_block is not None:
            for s in range(stage, stage + self.top_block.num_levels):
                self._out_feature_strides["p{}".format(s + 1)] = 2 ** (s + 1)
        self._out_features = list(self._out_feature_strides.keys())
        self._out_feature_channels = {k: out_channels for k in self._out_features}
        assert fuse_type in {"avg", "sum"}
        self._fuse_type = fuse_type
        self.needed = None  # None = compute every level (reference behaviour)

    def set_needed(self, names):
        """Only produce these output levels (and only the backbone features they depend on)."""
        self.needed = list(names)
        top_down = self._stages[::-1]
        lowest = min(int(n[1:]) for n in self.needed if int(n[1:]) in top_down)
        feats = [f for f, st in zip(self.in_features, self._stages) if st >= lowest]
        if hasattr(self.bottom_up, "compute_features"):
            self.bottom_up.compute_features = feats

    def forward(self, *args, **kwargs):
        return self.forward_dense(self.bottom_up(*args, **kwargs), (args, kwargs))

    def forward_dense(self, feats, bottom_up_call=None):
        """The dense top-down part on the bottom-up feature maps (static shapes: the part a CUDA graph can hold)."""
        if self.needed is None:
            return self._forward_all(feats)
        top_down_stages = self._stages[::-1]
        top_down_names = self.in_features[::-1]
        wanted = set(self.needed)
        lowest = min(int(n[1:]) for n in wanted if int(n[1:]) in top_down_stages)
        results = {}
        prev = None
        for st, name, lateral, output in zip(top_down_stages, top_down_names, self.lateral_convs, self.output_convs):
            if st < lowest:
                break
            lat = lateral(feats[name])
            if prev is not None:
                lat = lat + F.interpolate(prev, scale_factor=2, mode="nearest")
                if self._fuse_type == "avg":
                    lat = lat / 2
            prev = lat
            if "p%d" % st in wanted:
                results["p%d" % st] = output(prev)
        missing = wanted - set(results)
        if missing:  # a top-block level was requested: fall back to the full computation
            if bottom_up_call is None:
                raise RuntimeError("FPN.forward_dense: level(s) %r need the full bottom-up pass" % sorted(missing))
            args, kwargs = bottom_up_call
            return {k: v for k, v in self._forward_all(self.bottom_up(*args, **kwargs)).items() if k in wanted}
        return results

    def _forward_all(self, feats):
        x = [feats[f] for f in self.in_features[::-1]]
        results = []
        prev = self.lateral_convs[0](x[0])
        results.append(self.output_convs[0](prev))
        for f, lateral, output in zip(x[1:], self.lateral_convs[1:], self.output_convs[1:]):
            prev = lateral(f) + F.interpolate(prev, scale_factor=2, mode="nearest")
            if self._fuse_type == "avg":
                prev = prev / 2
            results.insert(0, output(prev))
        if self.top_block is not None:
            if self.top_block.in_feature in feats:
                top_in = feats[self.top_block.in_feature]
            else:
                top_in = results[self._out_features.index(self.top_block.in_feature)]
            results.extend(self.top_block(top_in))
        assert len(self._out_features) == len(results)
        return dict(zip(self._out_features, results))

    def output_shape(self):
        return {n: {"channels": self._out_feature_channels[n], "stride": self._out_feature_strides[n]}
                for n in self._out_features}


def build_resnet_fpn_backbone(config, input_shape, backend=None):
    """config.resnet / config.fpn as in VD/config.yaml:88-108 (fpn.py:18-38)."""
    bottom_up = build_sparse_resnet_backbone(config["resnet"], input_shape, backend=backend)
    fpn = config["fpn"]
    return FPN(bottom_up=bottom_up, in_features=fpn["in_features"], out_channels=fpn["out_channels"], norm=fpn["norm"],
               top_block=LastLevelMaxPool(in_feature=fpn["top_block_in_feature"]), fuse_type=fpn["fuse_type"])
