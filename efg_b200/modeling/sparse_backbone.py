"""Sparse 3-D backbones of the path, wired exactly like the reference but over an injectable
sparse-conv backend (default: the CUDA kernels in efg_b200.spconv).

  SparseResNet (+stem, residual blocks, z-collapse heads)  efg/modeling/backbones/sparse_net.py:79-309
  build_sparse_resnet_backbone                              sparse_net.py:318-397
  SpMiddleResNetFHD (CenterPoint)                           sparse_net.py:400-545

Parameter names follow the reference module tree so its checkpoints load unchanged
(e.g. ``stem.conv1.0.weight``, ``res2.0.shortcut.0.weight``, ``res3_out.1.running_mean``).

One deliberate, output-identical difference: ``out_features`` that no consumer reads can be
skipped (``compute_features``), see SURVEY.md §7 "dead compute".
"""
import numpy as np
from torch import nn

from ..backend import cuda_backend
from .norm import get_activation, get_norm


def _replace(x, feats):
    return x.replace_feature(feats)


def _fusable_tail(seq, activation):
    """The sequence is this repo's CUDA SparseSequential (it understands residual= / final_relu=), ends with a
    BatchNorm1d and the block's activation is a plain ReLU."""
    import inspect

    mods = list(seq._modules.values()) if hasattr(seq, "_modules") else []
    return (bool(mods) and isinstance(mods[-1], nn.BatchNorm1d) and isinstance(activation, nn.ReLU) and
            "residual" in inspect.signature(seq.forward).parameters)


class SparseBasicStem(nn.Module):
    """SparseConv3d s2 -> SubM -> SubM, each followed by norm + activation (sparse_net.py:79-95)."""

    def __init__(self, sp, in_channels=16, out_channels=32, stem_width=32, norm="BN1d", activation=None,
                 indice_key=None):
        super().__init__()
        self.out_channels = out_channels
        self.conv1 = sp.SparseSequential(
            sp.SparseConv3d(in_channels, stem_width, 3, 2, padding=1, bias=False),
            get_norm(norm, stem_width),
            get_activation(activation),
            sp.SubMConv3d(stem_width, stem_width, 3, padding=1, bias=False, indice_key=indice_key),
            get_norm(norm, stem_width),
            get_activation(activation),
            sp.SubMConv3d(stem_width, out_channels, 3, padding=1, bias=False, indice_key=indice_key),
            get_norm(norm, out_channels),
            get_activation(activation),
        )

    stride = 2

    def forward(self, x):
        return self.conv1(x)


class SparseBasicResBlock(nn.Module):
    """Two 3x3x3 convs with a residual; the first conv (and the shortcut) is a strided
    SparseConv3d when stride != 1, otherwise SubMConv3d (sparse_net.py:120-165)."""

    def __init__(self, sp, in_channels=32, out_channels=64, stride=1, norm="BN1d", activation=None, indice_key=None):
        super().__init__()
        self.in_channels, self.out_channels, self.stride = in_channels, out_channels, stride
        if in_channels != out_channels:
            self.shortcut = sp.SparseSequential(
                sp.SparseConv3d(in_channels, out_channels, 3, padding=1, stride=stride, bias=False),
                get_norm(norm, out_channels),
            )
        else:
            self.shortcut = None
        self.activation = get_activation(activation)
        if stride == 1:
            first = sp.SubMConv3d(in_channels, out_channels, kernel_size=3, stride=stride, padding=1, bias=False,
                                  indice_key=indice_key)
        else:
            first = sp.SparseConv3d(in_channels, out_channels, kernel_size=3, stride=stride, padding=1, bias=False)
        self.conv = sp.SparseSequential(
            first,
            get_norm(norm, out_channels),
            get_activation(activation),
            sp.SubMConv3d(out_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=False,
                          indice_key=indice_key),
            get_norm(norm, out_channels),
        )

    def forward(self, x):
        shortcut = self.shortcut(x) if self.shortcut is not None else x
        if _fusable_tail(self.conv, self.activation):
            # relu(bn2(conv2(.)) + shortcut): the add and the activation ride in the last norm's fused pass
            return self.conv(x, residual=shortcut, final_relu=True)
        out = self.conv(x)
        out = _replace(out, out.features + shortcut.features)
        return _replace(out, self.activation(out.features))


# mark the blocks as sparse modules of whichever backend builds them
def _sparse_class(sp, cls):
    return type(cls.__name__, (cls, sp.SparseModule), {})


class SparseResNet(nn.Module):
    def __init__(self, sp, stem, stages, out_features=None, norm=None):
        super().__init__()
        self._sp = [sp]  # list: keep the namespace out of the module registry
        self.stem = stem
        current_stride = self.stem.stride
        self._out_feature_strides = {"stem": current_stride}
        self._out_feature_channels = {"stem": self.stem.out_channels}
        self.stages_and_names = []
        for i, blocks in enumerate(stages):
            name = "res" + str(i + 2)
            stage = sp.SparseSequential(*blocks)
            self.add_module(name, stage)
            self.stages_and_names.append((stage, name))
            current_stride = int(current_stride * np.prod([b.stride for b in blocks]))
            self._out_feature_strides[name] = current_stride
            self._out_feature_channels[name] = blocks[-1].out_channels
        if out_features is None:
            out_features = [name]
        self._out_features = list(out_features)
        children = [n for n, _ in self.named_children()]
        for f in self._out_features:
            assert f in children, "Available children: {}".format(", ".join(children))
        # z-collapse heads: k(3,1,1) s(2,1,1) p(1,0,0) conv + norm + ReLU, then dense + fold D into C (:273-282)
        multipliers = [6, 3, 2]
        for idx, f in enumerate(self._out_features):
            ch = self._out_feature_channels[f]
            self.add_module(f + "_out", sp.SparseSequential(
                sp.SparseConv3d(ch, ch, (3, 1, 1), (2, 1, 1), padding=(1, 0, 0), bias=False),
                get_norm(norm, ch),
                nn.ReLU(),
            ))
            self._out_feature_channels[f] *= multipliers[idx]
        self.compute_features = None  # None = all of out_features (reference behaviour)

    def plan_geometry(self, coors, batch_size, input_shape):
        """Every strided rulebook of the forward pass for these voxel coordinates, WITHOUT features: the part of the
        pass that needs host reads (one output count per downsampling).  Returns the indice_dict to hand to forward();
        a data pipeline runs this one batch ahead on its own stream so that the feature pass never synchronises."""
        sp = self._sp[0]
        sparse_shape = np.array(input_shape[::-1]) + [1, 0, 0]
        x = sp.SparseConvTensor(None, coors.int(), sparse_shape, batch_size)
        wanted = self._out_features if self.compute_features is None else \
            [f for f in self._out_features if f in self.compute_features]
        self._plan_rulebooks(x, wanted)
        return x.indice_dict

    def forward(self, voxel_features, coors, batch_size, input_shape, indice_dict=None):
        sp = self._sp[0]
        sparse_shape = np.array(input_shape[::-1]) + [1, 0, 0]
        kw = {} if indice_dict is None else {"indice_dict": indice_dict}   # rulebooks planned ahead (plan_geometry)
        x = sp.SparseConvTensor(voxel_features, coors.int(), sparse_shape, batch_size, **kw)
        wanted = self._out_features if self.compute_features is None else \
            [f for f in self._out_features if f in self.compute_features]
        self._plan_rulebooks(x, wanted)
        stage_out = {}
        x = self.stem(x)
        if "stem" in wanted:
            stage_out["stem"] = x
        for stage, name in self.stages_and_names:
            x = stage(x)
            if name in wanted:
                stage_out[name] = x
        outputs = {}
        for f in wanted:
            out = getattr(self, f + "_out")(stage_out[f]).dense()
            n, c, d, h, w = out.shape
            outputs[f] = out.view(n, c * d, h, w)
        return outputs

    def _plan_rulebooks(self, x, wanted):
        """Build every strided rulebook of the forward pass up front (indices only).  Each one needs a host
        read of its output count; doing them back to back, before any feature kernel is queued, keeps
        those synchronisations from draining the GPU in the middle of the feature pipeline."""
        sp = self._sp[0]
        plan = getattr(sp, "strided_rulebook", None)
        if plan is None:
            return
        cur = x
        levels = {}
        rb = plan(cur, 3, 2, 1)  # stem conv
        cur = sp.SparseConvTensor(None, rb.out_indices, rb.out_shape, x.batch_size, indice_dict=x.indice_dict)
        cur._rows_sorted = True
        for _, name in self.stages_and_names:
            rb = plan(cur, 3, 2, 1)  # first block of the stage: main + shortcut conv share it
            cur = sp.SparseConvTensor(None, rb.out_indices, rb.out_shape, x.batch_size, indice_dict=x.indice_dict)
            cur._rows_sorted = True
            levels[name] = cur
        for f in wanted:
            if f in levels:
                plan(levels[f], (3, 1, 1), (2, 1, 1), (1, 0, 0))

    def output_shape(self):
        return {n: {"channels": self._out_feature_channels[n], "stride": self._out_feature_strides[n]}
                for n in self._out_features}


def build_sparse_resnet_backbone(config, in_channels, backend=None):
    """config: depth, norm, activation, stem_out_channels, res1_out_channels, out_features (sparse_net.py:318-397)."""
    sp = (backend or cuda_backend()).spconv
    depth = config["depth"]
    stem_width = {18: 16, "18b": 24, "18c": 32, 34: 16, "34b": 24, "34c": 32}[depth]
    blocks_per_stage = {18: [2, 2, 2, 2], "18b": [2, 2, 2, 2], "18c": [2, 2, 2, 2], 34: [3, 4, 6, 3],
                        "34b": [3, 4, 6, 3], "34c": [3, 4, 6, 3]}[depth]
    norm, activation = config["norm"], config["activation"]
    Stem = _sparse_class(sp, SparseBasicStem)
    Block = _sparse_class(sp, SparseBasicResBlock)
    stem = Stem(sp, in_channels=in_channels, out_channels=config["stem_out_channels"], norm=norm,
                activation=activation, stem_width=stem_width, indice_key="stem")
    out_features = list(config["out_features"])
    max_stage = max({"res2": 2, "res3": 3, "res4": 4, "res5": 5}[f] for f in out_features)
    c_in, c_out = config["stem_out_channels"], config["res1_out_channels"]
    stages = []
    for idx, stage_idx in enumerate(range(2, max_stage + 1)):
        blocks = []
        for i in range(blocks_per_stage[idx]):
            blocks.append(Block(sp, in_channels=c_in if i == 0 else c_out, out_channels=c_out,
                                stride=2 if i == 0 else 1, norm=norm, activation=activation,
                                indice_key="res" + str(stage_idx)))
        stages.append(blocks)
        c_in, c_out = c_out, c_out * 2
    return SparseResNet(sp, stem, stages, out_features=out_features, norm=norm)


# ------------------------------------------------------------------------------------------------
# CenterPoint middle encoder
# ------------------------------------------------------------------------------------------------
class SparseBasicBlock(nn.Module):
    """Legacy residual block of SpMiddleResNetFHD: two SubM convs WITH bias (bias = norm is not None),
    sparse_net.py:429-469."""

    expansion = 1

    def __init__(self, sp, inplanes, planes, stride=1, norm=None, indice_key=None):
        super().__init__()
        bias = norm is not None
        self.conv1 = sp.SubMConv3d(inplanes, planes, kernel_size=3, stride=stride, padding=1, bias=bias,
                                   indice_key=indice_key)
        self.bn1 = get_norm(norm, planes)
        self.relu = nn.ReLU()
        self.conv2 = sp.SubMConv3d(planes, planes, kernel_size=3, stride=stride, padding=1, bias=bias,
                                   indice_key=indice_key)
        self.bn2 = get_norm(norm, planes)
        self.stride = stride

    def forward(self, x):
        identity = x
        out = self.conv1(x)
        if out.features.is_cuda and isinstance(self.bn1, nn.BatchNorm1d):
            from .. import ops

            if ops.bn_act_supported(out.features, self.bn1):
                out = _replace(out, ops.bn_act(out.features, self.bn1, relu=True))
            else:
                out = _replace(out, self.relu(self.bn1(out.features)))
        else:
            out = _replace(out, self.relu(self.bn1(out.features)))
        out = self.conv2(out)
        from .. import ops

        if out.features.is_cuda and isinstance(self.bn2, nn.BatchNorm1d) and ops.bn_act_supported(out.features, self.bn2):
            return _replace(out, ops.bn_act(out.features, self.bn2, residual=identity.features, relu=True))
        out = _replace(out, self.bn2(out.features))
        out = _replace(out, out.features + identity.features)
        return _replace(out, self.relu(out.features))


class SpMiddleResNetFHD(nn.Module):
    """CenterPoint's sparse encoder (sparse_net.py:472-545): SubM 5->16, 2 blocks @16, then three
    stride-2 stages (32, 64, 128; the last with padding [0,1,1]) of 2 blocks each, a (3,1,1)/(2,1,1)
    z-collapse, dense, and D folded into C -> [B, 128*D', H/8, W/8]."""

    def __init__(self, num_input_features=128, out_features=("res3",), norm="BN1d", backend=None):
        super().__init__()
        sp = (backend or cuda_backend()).spconv
        self._sp = [sp]
        Block = _sparse_class(sp, SparseBasicBlock)

        def cbr(cin, cout, k, s, p):
            return [sp.SparseConv3d(cin, cout, k, s, padding=p, bias=False), get_norm(norm, cout), nn.ReLU(inplace=True)]

        self.conv_input = sp.SparseSequential(
            sp.SubMConv3d(num_input_features, 16, 3, bias=False, indice_key="res0"), get_norm(norm, 16),
            nn.ReLU(inplace=True))
        self.conv1 = sp.SparseSequential(Block(sp, 16, 16, norm=norm, indice_key="res0"),
                                         Block(sp, 16, 16, norm=norm, indice_key="res0"))
        self.conv2 = sp.SparseSequential(*cbr(16, 32, 3, 2, 1), Block(sp, 32, 32, norm=norm, indice_key="res1"),
                                         Block(sp, 32, 32, norm=norm, indice_key="res1"))
        self.conv3 = sp.SparseSequential(*cbr(32, 64, 3, 2, 1), Block(sp, 64, 64, norm=norm, indice_key="res2"),
                                         Block(sp, 64, 64, norm=norm, indice_key="res2"))
        self.conv4 = sp.SparseSequential(*cbr(64, 128, 3, 2, [0, 1, 1]),
                                         Block(sp, 128, 128, norm=norm, indice_key="res3"),
                                         Block(sp, 128, 128, norm=norm, indice_key="res3"))
        self.extra_conv = sp.SparseSequential(sp.SparseConv3d(128, 128, (3, 1, 1), (2, 1, 1), bias=False),
                                              get_norm(norm, 128), nn.ReLU())

    def plan_geometry(self, coors, batch_size, input_shape):
        """The strided rulebooks of the four downsampling convolutions for these voxel coordinates, without features
        (see SparseResNet.plan_geometry): returns the indice_dict to hand to forward()."""
        sp = self._sp[0]
        plan = getattr(sp, "strided_rulebook", None)
        if plan is None:
            return None
        sparse_shape = np.array(input_shape[::-1]) + [1, 0, 0]
        cur = sp.SparseConvTensor(None, coors.int(), sparse_shape, batch_size)
        indice_dict = cur.indice_dict
        for seq in (self.conv2, self.conv3, self.conv4, self.extra_conv):
            conv = seq[0]
            rb = plan(cur, conv.kernel_size, conv.stride, conv.padding)
            cur = sp.SparseConvTensor(None, rb.out_indices, rb.out_shape, batch_size, indice_dict=indice_dict)
            cur._rows_sorted = True
        return indice_dict

    def forward(self, voxel_features, coors, batch_size, input_shape, indice_dict=None):
        sp = self._sp[0]
        sparse_shape = np.array(input_shape[::-1]) + [1, 0, 0]
        kw = {} if indice_dict is None else {"indice_dict": indice_dict}   # rulebooks planned ahead (plan_geometry)
        x = sp.SparseConvTensor(voxel_features, coors.int(), sparse_shape, batch_size, **kw)
        x = self.conv_input(x)
        x_conv1 = self.conv1(x)
        x_conv2 = self.conv2(x_conv1)
        x_conv3 = self.conv3(x_conv2)
        x_conv4 = self.conv4(x_conv3)
        ret = self.extra_conv(x_conv4).dense()
        n, c, d, h, w = ret.shape
        ret = ret.view(n, c * d, h, w)
        return ret
