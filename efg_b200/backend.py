"""The two operator families the models of the path are wired over.

The product has exactly one backend: the CUDA kernels behind ``efg_b200.spconv`` and
``efg_b200.operators.BoxAttnFunction``.  The model constructors accept a backend object so that the
test-suite and bench.py's CPU-baseline leg can wire the SAME module graph over the CPU oracle
(``oracle.backend_cpu``) — the oracle is injected from outside, never imported from here.
"""


class Backend:
    def __init__(self, name, spconv, box_attn, rotate_nms=None):
        self.name = name
        self.spconv = spconv  # namespace with SparseConvTensor, SparseSequential, SubMConv3d, SparseConv3d, SparseModule
        self.box_attn = box_attn  # callable(value, shapes, level_start, loc, attn, im2col_step) -> [B, LQ, H*C]
        # callable(boxes [N,7] (x,y,z,l,w,h,theta), scores, thresh, pre_maxsize, post_max_size) -> selected indices
        self.rotate_nms = rotate_nms

    def __deepcopy__(self, memo):  # shared by module clones (get_clones deep-copies layers)
        return self


_cuda = None


def cuda_backend():
    global _cuda
    if _cuda is None:
        from . import spconv
        from .operators import BoxAttnFunction
        from .operators.iou3d_nms import rotate_nms_pcdet

        _cuda = Backend("efgb200-cuda", spconv, BoxAttnFunction.apply, rotate_nms_pcdet)
    return _cuda
