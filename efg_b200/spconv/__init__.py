"""``spconv`` / ``spconv.pytorch`` surface used by efg/modeling/backbones/sparse_net.py:6-11,
implemented on the efgb200 CUDA rulebook + gather-GEMM kernels (no upstream spconv involved)."""
from .pytorch import (SparseConv3d, SparseConvTensor, SparseModule, SparseSequential, SubMConv3d, ToDense,
                      strided_rulebook)

__all__ = ["SparseConv3d", "SparseConvTensor", "SparseModule", "SparseSequential", "SubMConv3d", "ToDense",
           "strided_rulebook"]
