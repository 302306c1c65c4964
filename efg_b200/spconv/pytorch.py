"""Sparse tensors and 3-D sparse convolution modules with the spconv-2.x names and state_dict layout.

What the reference uses from spconv (efg/modeling/backbones/sparse_net.py):
  SparseConvTensor(features, indices, spatial_shape, batch_size)   :289, :531
  .features / .indices / .replace_feature(x) / .dense()            :19-25, :304, :541
  SparseModule, SparseSequential (plain nn.Modules act on .features) :79-95
  SubMConv3d / SparseConv3d(in, out, kernel_size, stride, padding, bias, indice_key) :86-92, :277, :513
Weights are ``[Cout, kD, kH, kW, Cin]`` (spconv 2.x) so checkpoints map 1:1.

Rulebooks are built on device (csrc/rulebook.cu) as dense neighbour tables and cached per
``indice_key`` on the tensor, the convolution itself is csrc/spconv.cu.
"""
import math
from collections import OrderedDict

import torch
from torch import nn
from torch.autograd import Function

from .. import ops


def _triple(v):
    if isinstance(v, (list, tuple)):
        assert len(v) == 3, v
        return [int(x) for x in v]
    return [int(v)] * 3


class _Rulebook:
    """Neighbour tables of one convolution geometry over one active set."""

    __slots__ = ("nbr", "nbr_t", "subm", "out_indices", "out_shape", "ksize", "in_indices")

    def __init__(self, nbr, nbr_t, subm, out_indices, out_shape, ksize, in_indices):
        self.nbr, self.nbr_t, self.subm = nbr, nbr_t, subm
        self.out_indices, self.out_shape, self.ksize = out_indices, out_shape, ksize
        self.in_indices = in_indices

    def pairs(self):
        """(tap, in_row, out_row) triples — the classic indice-pair list, for tests/inspection."""
        o, k = torch.nonzero(self.nbr >= 0, as_tuple=True)
        return k, self.nbr[o, k].long(), o


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, grid=None, voxel_num=None, indice_dict=None,
                 benchmark=False):
        """features [M, C] f32; indices [M, 4] i32 (batch, z, y, x); spatial_shape [D, H, W]."""
        self.features = features
        self.indices = indices if indices.dtype == torch.int32 else indices.int()
        if not self.indices.is_contiguous():
            self.indices = self.indices.contiguous()
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)
        self.indice_dict = indice_dict if indice_dict is not None else {}
        self.grid = grid
        self.voxel_num = voxel_num
        self.benchmark = benchmark
        self._rows_sorted = False
        self._planes = None  # bf16 operand planes of `features` (ops.split_bf16 layout) when a producer emitted them

    def replace_feature(self, feature):
        out = SparseConvTensor(feature, self.indices, self.spatial_shape, self.batch_size, self.grid, self.voxel_num,
                               self.indice_dict, self.benchmark)
        out._rows_sorted = self._rows_sorted
        return out

    @property
    def spatial_size(self):
        return int(self.spatial_shape[0] * self.spatial_shape[1] * self.spatial_shape[2])

    def find_indice_pair(self, key):
        return self.indice_dict.get(key) if key is not None else None

    def dense(self, channels_first=True):
        out = _ToDenseFn.apply(self.features, self.indices, self.batch_size, tuple(self.spatial_shape))
        if not channels_first:
            out = out.permute(0, 2, 3, 4, 1).contiguous()
        return out

    @property
    def sparity(self):
        return self.indices.shape[0] / max(self.spatial_size * self.batch_size, 1)


class _ToDenseFn(Function):
    @staticmethod
    def forward(ctx, features, indices, batch_size, spatial_shape):
        ctx.save_for_backward(indices)
        return ops.sparse_to_dense(features.contiguous(), indices, batch_size, spatial_shape)

    @staticmethod
    def backward(ctx, grad):
        (indices,) = ctx.saved_tensors
        return ops.dense_to_sparse(grad.contiguous(), indices), None, None, None


def _padded_channels(c_in, c_out, taps):
    """Narrow inputs (the 5 / 6 point features of the first convolution) do not fit the tensor-core kernels, whose
    reduction runs in 16-byte pieces; zero-padded to 16 channels they do.  Returns 16 or None."""
    if c_in >= 16 or ops.spconv_tc_supported(c_in, c_out, taps) or not ops.PAD_NARROW_INPUTS:
        return None
    return 16 if ops.spconv_tc_supported(16, c_out, taps) and ops.spconv_tc_wgrad_supported(16, c_out, taps) else None


class _SparseConvFn(Function):
    """out = conv(features) over a rulebook; backward = dgrad (same kernel on the transposed
    rulebook) + wgrad.  ``weight`` is the parameter viewed as [Cout, taps, Cin]."""

    @staticmethod
    def forward(ctx, features, weight, bias, rulebook, planes=None):
        features = features.contiguous()
        weight = weight.contiguous()
        c_out, taps, c_in = weight.shape
        pad_to = _padded_channels(c_in, c_out, taps) if features.is_cuda else None
        if pad_to is not None:
            # zero channels contribute nothing: same result from the tcgen05 kernels (forward here, wgrad in backward)
            features = torch.nn.functional.pad(features, (0, pad_to - c_in))
            out = ops.spconv_tc(features, ops.padded_weight(weight, pad_to), bias, rulebook.nbr, 0)
        elif ops.spconv_tc_supported(c_in, c_out, taps):
            # `planes`: bf16 operand planes of `features` already produced by the fused BN + ReLU pass upstream
            out = ops.spconv_tc(features, weight, bias, rulebook.nbr, 0, planes=planes)
        else:
            out = ops.spconv_forward(features, weight.permute(1, 2, 0).contiguous(), bias, rulebook.nbr)
        ctx.rulebook = rulebook
        ctx.has_bias = bias is not None
        ctx.pad_to = pad_to
        ctx.save_for_backward(features, weight)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        features, weight = ctx.saved_tensors
        rb = ctx.rulebook
        grad_out = grad_out.contiguous()
        c_out, taps, c_in = weight.shape
        d_feat = d_w = d_b = None
        if ctx.needs_input_grad[0]:
            # submanifold: nbr[j, K-1-k] is the output that input j feeds through tap k (odd kernels are
            # symmetric), so dgrad reuses the forward table with mirrored taps; regular: transposed table
            table = rb.nbr if rb.subm else rb.nbr_t
            if ops.spconv_tc_supported(c_out, c_in, taps):
                d_feat = ops.spconv_tc(grad_out, weight, None, table, 2 if rb.subm else 1)
            else:
                w_t = weight.permute(1, 0, 2)  # [taps, Cout, Cin]
                if rb.subm:
                    w_t = w_t.flip(0)
                d_feat = ops.spconv_forward(grad_out, w_t.contiguous(), None, table)
        if ctx.needs_input_grad[1]:
            if ctx.pad_to is not None:   # `features` are the padded rows saved by forward
                d_w = ops.spconv_tc_wgrad(features, grad_out, rb.nbr, taps, ctx.pad_to, c_out)[..., :c_in].contiguous()
            elif ops.spconv_tc_wgrad_supported(c_in, c_out, taps):
                d_w = ops.spconv_tc_wgrad(features, grad_out, rb.nbr, taps, c_in, c_out)
            else:
                d_w = ops.spconv_wgrad(features, grad_out, rb.nbr, taps, c_in, c_out).permute(2, 0, 1).contiguous()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            d_b = grad_out.sum(0)
        return d_feat, d_w, d_b, None, None


def strided_rulebook(x, kernel_size, stride, padding):
    """Rulebook of a regular sparse conv over ``x``'s active set, cached on the tensor by geometry: the
    main and the shortcut conv of a residual block (sparse_net.py:126,136) share one table (and one host
    sync), and a backbone can build all of them ahead of the feature pass (SparseResNet.forward)."""
    k, s_, p_ = _triple(kernel_size), _triple(stride), _triple(padding)
    key = ("__strided__", id(x.indices), tuple(k), tuple(s_), tuple(p_))
    cached = x.indice_dict.get(key)
    if cached is not None and cached.in_indices is x.indices:
        return cached
    out_indices, out_shape, nbr, nbr_t = ops.sparse_rulebook(x.indices, x.batch_size, x.spatial_shape, k, s_, p_)
    rb = _Rulebook(nbr, nbr_t, False, out_indices, out_shape, k, x.indices)
    x.indice_dict[key] = rb
    return rb


class SparseModule(nn.Module):
    """Marker base class: SparseSequential hands these the SparseConvTensor itself."""


def _is_sparse_module(m):
    return isinstance(m, SparseModule)


def _wants_planes(module, channels):
    """True when `module` is a sparse conv that will take the pre-split tensor-core path on `channels` inputs."""
    if not isinstance(module, SparseConvolution) or ops.CONV_PRECISION != "bf16x3" or not ops.USE_PLANES:
        return False
    taps = int(module.kernel_size[0] * module.kernel_size[1] * module.kernel_size[2])
    return taps > 1 and module.in_channels == channels and ops.spconv_tc_supported(channels, module.out_channels, taps) and \
        bool(ops._lib.lib().efgb_spconv_tc_planes_supported(channels, module.out_channels, taps))


class SparseSequential(SparseModule):
    """nn.Sequential for mixed sparse / dense modules: SparseModules see the SparseConvTensor,
    every other module (BatchNorm1d, ReLU, ...) is applied to ``.features``."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        if len(args) == 1 and isinstance(args[0], OrderedDict):
            for key, module in args[0].items():
                self.add_module(key, module)
        else:
            for idx, module in enumerate(args):
                if module is None:
                    continue
                self.add_module(str(idx), module)
        for name, module in kwargs.items():
            if name in self._modules:
                raise ValueError("name exists.")
            self.add_module(name, module)

    def __getitem__(self, idx):
        if not (-len(self) <= idx < len(self)):
            raise IndexError("index {} is out of range".format(idx))
        if idx < 0:
            idx += len(self)
        return list(self._modules.values())[idx]

    def __len__(self):
        return len(self._modules)

    def add(self, module, name=None):
        if name is None:
            name = str(len(self._modules))
            if name in self._modules:
                raise KeyError("name exists")
        self.add_module(name, module)

    def forward(self, input, residual=None, final_relu=False):
        """`residual` / `final_relu`: fold `relu(x + residual)` of a residual block into the LAST BatchNorm1d of the
        sequence (only valid when the sequence ends with that norm; the blocks of sparse_backbone.py use it)."""
        mods = list(self._modules.values())
        i = 0
        while i < len(mods):
            module = mods[i]
            if _is_sparse_module(module):
                input = module(input)
            elif isinstance(input, SparseConvTensor):
                if input.indices.shape[0] != 0:
                    if isinstance(module, nn.BatchNorm1d) and ops.bn_act_supported(input.features, module):
                        # BatchNorm1d [+ residual] [+ ReLU] in one fused pass each way (csrc/batchnorm.cu), emitting the
                        # operand planes of the result when a tensor-core conv consumes it next
                        span = 2 if (i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)) else 1
                        tail = i + span == len(mods)
                        relu = span == 2 or (tail and final_relu)
                        res = residual.features if (tail and residual is not None) else None
                        consumer = mods[i + span] if i + span < len(mods) else None
                        want = _wants_planes(consumer, input.features.shape[1])
                        out = ops.bn_act(input.features, module, residual=res, relu=relu, want_planes=want)
                        feats, planes = out if want else (out, None)
                        input = input.replace_feature(feats)
                        input._planes = planes
                        if tail:
                            residual, final_relu = None, False
                        i += span - 1
                    else:
                        input = input.replace_feature(module(input.features))
            else:
                input = module(input)
            i += 1
        if residual is not None:   # the sequence did not end with a fusable norm: plain add + activation
            input = input.replace_feature(input.features + residual.features)
        if final_relu:
            input = input.replace_feature(torch.relu(input.features))
        return input


class SparseConvolution(SparseModule):
    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, subm=False, output_padding=0, transposed=False, inverse=False, indice_key=None,
                 algo=None, fp32_accum=None, name=None):
        super().__init__()
        assert ndim == 3, "only 3-D sparse convolution is on the hot path"
        assert groups == 1 and not transposed and not inverse, "groups/transposed/inverse convs are not used by EFG"
        assert _triple(dilation) == [1, 1, 1], "dilation is not used by EFG"
        self.ndim = ndim
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _triple(kernel_size)
        self.stride = _triple(stride)
        self.padding = _triple(padding)
        self.dilation = _triple(dilation)
        self.subm = subm
        self.indice_key = indice_key
        self.conv1x1 = self.kernel_size == [1, 1, 1] and self.stride == [1, 1, 1]
        self.weight = nn.Parameter(torch.empty(out_channels, *self.kernel_size, in_channels))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def extra_repr(self):
        s = "{in_channels}, {out_channels}, kernel_size={kernel_size}, stride={stride}, padding={padding}"
        if self.bias is None:
            s += ", bias=False"
        if self.indice_key is not None:
            s += ", indice_key={indice_key}"
        return s.format(**self.__dict__)

    def reset_parameters(self):
        # same distribution as torch.nn.Conv3d / spconv: kaiming_uniform(a=sqrt(5)) with fan_in = Cin * K
        fan_in = self.in_channels * self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2]
        bound = 1.0 / math.sqrt(fan_in) if fan_in > 0 else 0.0
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            if self.bias is not None:
                self.bias.uniform_(-bound, bound)

    def _rulebook(self, x):
        if self.subm:
            cached = x.find_indice_pair(self.indice_key)
            if cached is not None and cached.subm and cached.ksize == self.kernel_size and \
                    cached.in_indices is x.indices:
                return cached
            nbr = ops.subm_rulebook(x.indices, x.batch_size, x.spatial_shape, self.kernel_size,
                                    rows_sorted=x._rows_sorted)
            rb = _Rulebook(nbr, None, True, x.indices, x.spatial_shape, self.kernel_size, x.indices)
        else:
            rb = strided_rulebook(x, self.kernel_size, self.stride, self.padding)
        if self.indice_key is not None:
            x.indice_dict[self.indice_key] = rb
        return rb

    def forward(self, x):
        assert isinstance(x, SparseConvTensor), "sparse convolution expects a SparseConvTensor"
        rb = self._rulebook(x)
        taps = self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2]
        feats = _SparseConvFn.apply(x.features, self.weight.view(self.out_channels, taps, self.in_channels), self.bias, rb,
                                    getattr(x, "_planes", None))
        out = SparseConvTensor(feats, rb.out_indices, rb.out_shape, x.batch_size, x.grid, x.voxel_num, x.indice_dict,
                               x.benchmark)
        out._rows_sorted = x._rows_sorted if self.subm else True
        return out


class SparseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, algo=None, fp32_accum=None, name=None):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                         indice_key=indice_key)


class SubMConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, algo=None, fp32_accum=None, name=None):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, subm=True,
                         indice_key=indice_key)


class ToDense(SparseModule):
    def forward(self, x):
        return x.dense()
