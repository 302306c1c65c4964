"""ORACLE (test infrastructure) — CPU stand-ins for the ``spconv.pytorch`` names, built on
oracle/sparse_conv.py.  Used to run the model wiring of sparse_net.py on the CPU as the checker
and as bench.py's CPU baseline; never imported by efg_b200.  See sparse_conv.py for what is
restated and why parity with upstream spconv is unpinned."""
import math
from collections import OrderedDict

import numpy as np
import torch
from torch import nn

from . import sparse_conv as sc


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, indice_dict=None):
        self.features = features
        self.indices = indices.int()
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)
        self.indice_dict = indice_dict if indice_dict is not None else {}

    def replace_feature(self, feature):
        return SparseConvTensor(feature, self.indices, self.spatial_shape, self.batch_size, self.indice_dict)

    def dense(self):
        return sc.to_dense(self.features, self.indices.cpu().numpy(), self.batch_size, self.spatial_shape)


class SparseModule(nn.Module):
    pass


class SparseSequential(SparseModule):
    def __init__(self, *args):
        super().__init__()
        if len(args) == 1 and isinstance(args[0], OrderedDict):
            for k, m in args[0].items():
                self.add_module(k, m)
        else:
            for i, m in enumerate(args):
                if m is not None:
                    self.add_module(str(i), m)

    def forward(self, x):
        for m in self._modules.values():
            if isinstance(m, SparseModule):
                x = m(x)
            elif isinstance(x, SparseConvTensor):
                if x.indices.shape[0] != 0:
                    x = x.replace_feature(m(x.features))
            else:
                x = m(x)
        return x


class _Conv(SparseModule):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, indice_key=None,
                 subm=False):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = sc._triple(kernel_size)
        self.stride, self.padding = sc._triple(stride), sc._triple(padding)
        self.subm, self.indice_key = subm, indice_key
        self.weight = nn.Parameter(torch.empty(out_channels, *self.kernel_size, in_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        bound = 1.0 / math.sqrt(in_channels * int(np.prod(self.kernel_size)))
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            if self.bias is not None:
                self.bias.uniform_(-bound, bound)

    def forward(self, x):
        coords = x.indices.cpu().numpy()
        if self.subm:
            key = self.indice_key
            cached = x.indice_dict.get(key) if key is not None else None
            if cached is not None and cached[0] is x.indices:
                nbr = cached[1]
            else:
                nbr = sc.subm_rulebook(coords, x.batch_size, x.spatial_shape, self.kernel_size)
                if key is not None:
                    x.indice_dict[key] = (x.indices, nbr)
            out_idx, out_shape = x.indices, x.spatial_shape
        else:
            oc, out_shape, nbr, _ = sc.sparse_rulebook(coords, x.batch_size, x.spatial_shape, self.kernel_size,
                                                       self.stride, self.padding)
            out_idx = torch.from_numpy(oc).to(x.indices.device)
        feats = sc.conv(x.features, self.weight, self.bias, nbr)
        return SparseConvTensor(feats, out_idx, out_shape, x.batch_size, x.indice_dict)


class SubMConv3d(_Conv):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None):
        super().__init__(in_channels, out_channels, kernel_size, 1, 0, bias, indice_key, subm=True)


class SparseConv3d(_Conv):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, bias, indice_key, subm=False)
