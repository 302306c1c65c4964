"""The two operator families the models of the path are wired over.

The product has exactly one backend: the CUDA kernels behind ``efg_b200.spconv`` and
``efg_b200.operators.BoxAttnFunction``.  The model constructors accept a backend object so that the
test-suite and bench.py's CPU-baseline leg can wire the SAME module graph over the CPU oracle
(``oracle.backend_cpu``) — the oracle is injected from outside, never imported from here.
"""


class Backend:
    def __init__(self, name, spconv, box_attn):
        self.name = name
        self.spconv = spconv  # namespace with SparseConvTensor, SparseSequential, SubMConv3d, SparseConv3d, SparseModule
        self.box_attn = box_attn  # callable(value, shapes, level_start, loc, attn, im2col_step) -> [B, LQ, H*C]

    def __deepcopy__(self, memo):  # shared by module clones (get_clones deep-copies layers)
        return self


_cuda = None


def cuda_backend():
    global _cuda
    if _cuda is None:
        from . import spconv
        from .operators import BoxAttnFunction

        _cuda = Backend("efgb200-cuda", spconv, BoxAttnFunction.apply)
    return _cuda
