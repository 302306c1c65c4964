"""``get_norm`` / ``get_activation`` (efg/modeling/common/batch_norm.py:140-185) for the norm and
activation types the five 3D configs use."""
from torch import nn


def get_norm(norm, out_channels):
    """norm: "" | "BN" | "BN1d" | "GN" | ["BN", {kwargs}] | callable -> nn.Module or None."""
    args = None
    if isinstance(norm, (list, tuple)):
        norm, args = norm
    if isinstance(norm, str):
        if len(norm) == 0:
            return None
        table = {
            "BN": nn.BatchNorm2d,
            "BN1d": nn.BatchNorm1d,
            "GN": lambda channels: nn.GroupNorm(32, channels),
            "nnSyncBN": nn.SyncBatchNorm,
        }
        if norm not in table:
            raise KeyError("norm type %r is not on the 3D-detection path (supported: %s)" % (norm, sorted(table)))
        norm = table[norm]
    return norm(out_channels, **dict(args)) if args else norm(out_channels)


def get_activation(activation):
    """activation: None | {type: ReLU|ReLU6|LeakyReLU|..., inplace: bool} -> nn.Module or None."""
    if activation is None:
        return None
    atype = activation["type"] if isinstance(activation, dict) else activation.type
    inplace = activation["inplace"] if isinstance(activation, dict) else activation.inplace
    table = {"ReLU": nn.ReLU, "ReLU6": nn.ReLU6, "LeakyReLU": nn.LeakyReLU, "SiLU": nn.SiLU, "GELU": nn.GELU}
    if atype not in table:
        raise KeyError("activation %r not supported" % (atype,))
    if atype in ("GELU",):
        return table[atype]()
    return table[atype](inplace=inplace)


class TokenLinear(nn.Linear):
    """nn.Linear (same parameters / state_dict) whose forward runs on the tcgen05 gather-GEMM kernels with an
    identity rulebook when the input has many rows (the 70k-token linears of the box-attention encoder) and
    the module belongs to the CUDA backend; otherwise plain F.linear (cuBLAS / CPU)."""

    def __init__(self, in_features, out_features, bias=True, backend=None):
        super().__init__(in_features, out_features, bias=bias)
        self._tc = backend is not None and getattr(backend, "name", "") == "efgb200-cuda"

    def forward(self, x):
        if self._tc and x.is_cuda and x.dtype == self.weight.dtype:
            from .. import ops

            rows = x.numel() // x.shape[-1]
            if ops.dense_linear_supported(rows, self.in_features, self.out_features):
                return ops.dense_linear(x, self.weight, self.bias)
        return nn.functional.linear(x, self.weight, self.bias)
