"""Dense BEV neck of CenterPoint (efg/modeling/backbones/configurable_rpn.py:14-121): per stage a
stride-s 3x3 conv followed by ``layer_num`` 3x3 convs (BN + ReLU each), each stage upsampled back with a
transposed conv and concatenated.  torch/cuDNN library layers."""
import numpy as np
import torch
from torch import nn

from .norm import get_norm


class Sequential(nn.Sequential):
    """nn.Sequential with ``add`` (efg/modeling/utils.py:8); children are named by position."""

    def add(self, module, name=None):
        if module is None:
            return
        self.add_module(str(len(self._modules)) if name is None else name, module)


class RPN(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self._layer_strides = list(cfg["ds_layer_strides"])
        self._num_filters = list(cfg["ds_num_filters"])
        self._layer_nums = list(cfg["layer_nums"])
        self._upsample_strides = list(cfg["us_layer_strides"])
        self._num_upsample_filters = list(cfg["us_num_filters"])
        self._num_input_features = cfg["num_input_features"]
        self.num_channels = sum(self._num_upsample_filters)
        self._norm_cfg = cfg["norm"]
        assert len(self._layer_strides) == len(self._layer_nums) == len(self._num_filters)
        assert len(self._num_upsample_filters) == len(self._upsample_strides)
        self._upsample_start_idx = len(self._layer_nums) - len(self._upsample_strides)
        ratios = [self._upsample_strides[i] / np.prod(self._layer_strides[:i + self._upsample_start_idx + 1])
                  for i in range(len(self._upsample_strides))]
        assert all(r == ratios[0] for r in ratios)

        in_filters = [self._num_input_features, *self._num_filters[:-1]]
        blocks, deblocks = [], []
        for i, layer_num in enumerate(self._layer_nums):
            blocks.append(self._make_layer(in_filters[i], self._num_filters[i], layer_num, self._layer_strides[i]))
            j = i - self._upsample_start_idx
            if j >= 0:
                stride, cout = self._upsample_strides[j], self._num_upsample_filters[j]
                if stride > 1:
                    up = nn.ConvTranspose2d(self._num_filters[i], cout, stride, stride=stride, bias=False)
                else:
                    s = int(np.round(1 / stride))
                    up = nn.Conv2d(self._num_filters[i], cout, s, stride=s, bias=False)
                deblocks.append(Sequential(up, get_norm(self._norm_cfg, cout), nn.ReLU()))
        self.blocks = nn.ModuleList(blocks)
        self.deblocks = nn.ModuleList(deblocks)

    def _make_layer(self, inplanes, planes, num_blocks, stride=1):
        block = Sequential(nn.ZeroPad2d(1), nn.Conv2d(inplanes, planes, 3, stride=stride, bias=False),
                           get_norm(self._norm_cfg, planes), nn.ReLU())
        for _ in range(num_blocks):
            block.add(nn.Conv2d(planes, planes, 3, padding=1, bias=False))
            block.add(get_norm(self._norm_cfg, planes))
            block.add(nn.ReLU())
        return block

    def forward(self, x):
        ups = []
        for i, blk in enumerate(self.blocks):
            x = blk(x)
            if i - self._upsample_start_idx >= 0:
                ups.append(self.deblocks[i - self._upsample_start_idx](x))
        return torch.cat(ups, dim=1) if ups else x
