"""ctypes binding of the C ABI declared in ``include/efgb200.h``.

The library is the product: there is no CPU or PyTorch fallback.  ``lib()`` raises if
``libefgb200.so`` has not been built (``python -m efg_b200._build`` or ``__graft_entry__.build()``).
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# EFGB_LIB_VARIANT=<name> loads libefgb200_<name>.so (an A/B build of the same sources, see _build.VARIANTS)
_VARIANT = os.environ.get("EFGB_LIB_VARIANT", "")
LIB_PATH = os.path.join(_HERE, "libefgb200%s.so" % ("_" + _VARIANT if _VARIANT else ""))

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_sz = ctypes.c_size_t
_host_i32x3 = ctypes.POINTER(ctypes.c_int32)
_host_f32 = ctypes.POINTER(ctypes.c_float)

# name -> (restype, argtypes); mirrors include/efgb200.h one to one
SIGNATURES = {
    "efgb_last_error": (ctypes.c_char_p, []),
    "efgb_version": (_int, []),
    "efgb_launch_count": (ctypes.c_uint64, []),
    "efgb_voxelize_workspace_bytes": (_sz, [_i64, _int]),
    "efgb_hard_voxelize": (_int, [_vp, _i64, _int, _vp, _int, _host_f32, _host_f32, _int, _int,
                                  _vp, _vp, _int, _vp, _vp, _vp, _vp, _sz, _vp]),
    "efgb_dynamic_voxelize": (_int, [_vp, _i64, _int, _host_f32, _host_f32, _vp, _vp]),
    "efgb_scatter_workspace_bytes": (_sz, [_i64, _host_i32x3]),
    "efgb_scatter_phase1": (_int, [_vp, _i64, _host_i32x3, _vp, _vp, _sz, _vp]),
    "efgb_scatter_phase2": (_int, [_vp, _vp, _i64, _int, _host_i32x3, _int, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "efgb_scatter_backward": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _int, _int, _i64, _vp, _vp, _sz, _vp]),
    "efgb_rulebook_workspace_bytes": (_sz, [_int, _host_i32x3, _i64]),
    "efgb_subm_rulebook": (_int, [_vp, _i64, _int, _host_i32x3, _host_i32x3, _int, _vp, _vp, _sz, _vp]),
    "efgb_sparse_rulebook_phase1": (_int, [_vp, _i64, _int, _host_i32x3, _host_i32x3, _host_i32x3, _host_i32x3,
                                           _host_i32x3, _vp, _vp, _sz, _vp]),
    "efgb_sparse_rulebook_phase2": (_int, [_vp, _i64, _int, _host_i32x3, _host_i32x3, _host_i32x3, _host_i32x3,
                                           _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "efgb_spconv_forward": (_int, [_vp, _i64, _int, _vp, _vp, _vp, _i64, _int, _int, _vp, _vp]),
    "efgb_spconv_tc_supported": (_int, [_int, _int, _int]),
    "efgb_spconv_tc_packed_bytes": (_sz, [_int, _int, _int, _int]),
    "efgb_spconv_tc_pack": (_int, [_vp, _int, _int, _int, _int, _int, _vp, _vp]),
    "efgb_spconv_tc_pack_blocks": (_i64, [_int, _int, _int, _int]),
    "efgb_spconv_tc_pack_batched": (_int, [_vp, _int, _i64, _vp]),
    "efgb_spconv_tc_forward": (_int, [_vp, _i64, _int, _vp, _vp, _vp, _i64, _int, _int, _int, _vp, _vp]),
    "efgb_split_bf16": (_int, [_vp, _i64, _int, _vp, _vp]),
    "efgb_spconv_tc_planes_supported": (_int, [_int, _int, _int]),
    "efgb_spconv_tc_forward_planes": (_int, [_vp, _i64, _int, _vp, _vp, _vp, _i64, _int, _int, _int, _vp, _vp]),
    "efgb_spconv_tc_forward_ex": (_int, [_vp, _i64, _int, _vp, _vp, _vp, _i64, _int, _int, _int, _int, _vp, _vp]),
    "efgb_spconv_tc_wgrad_supported": (_int, [_int, _int, _int]),
    "efgb_spconv_tc_wgrad": (_int, [_vp, _i64, _int, _vp, _vp, _i64, _int, _int, _int, _vp, _vp]),
    "efgb_spconv_wgrad": (_int, [_vp, _i64, _int, _vp, _vp, _i64, _int, _int, _vp, _vp]),
    "efgb_sparse_to_dense": (_int, [_vp, _vp, _i64, _int, _int, _host_i32x3, _vp, _vp]),
    "efgb_dense_to_sparse": (_int, [_vp, _vp, _i64, _int, _int, _host_i32x3, _vp, _vp]),
    "efgb_augment_workspace_bytes": (_sz, [_i64]),
    "efgb_paste_points": (_int, [_vp, _vp, _vp, _int, _i64, _vp, _i64, _int, _vp, _int, _vp, _vp]),
    "efgb_augment_points": (_int, [_vp, _i64, _int, _int, _int, ctypes.c_float, ctypes.c_float, ctypes.c_float, _host_f32, _host_f32,
                                   _vp, _vp, _vp, _sz, _vp]),
    "efgb_draw_gaussians": (_int, [_vp, _int, _int, _int, _vp, _vp]),
    "efgb_bn_supported": (_int, [_int]),
    "efgb_bn_workspace_bytes": (_sz, [_i64, _int]),
    "efgb_bn_forward": (_int, [_vp, _i64, _int, _vp, _vp, _vp, _int, ctypes.c_float, ctypes.c_float, _vp, _vp, _int, _vp, _vp, _vp,
                               _vp, _vp, _sz, _vp]),
    "efgb_bn_backward": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _int, _int, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "efgb_boxes_bev_workspace_bytes": (_sz, [_i64, _i64]),
    "efgb_boxes_bev": (_int, [_vp, _i64, _vp, _i64, _int, _vp, _vp, _sz, _vp]),
    "efgb_nms_bev_workspace_bytes": (_sz, [_i64]),
    "efgb_nms_bev": (_int, [_vp, _i64, ctypes.c_float, _int, _vp, _vp, _vp, _sz, _vp]),
    "efgb_lsa_batched": (_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32),
                                ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int64), _int, _vp, _vp, _vp]),
    "efgb_lsa_batched_status": (_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32),
                                       ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int64), _int, _vp, _vp, _vp, _vp]),
    "efgb_add_layernorm_supported": (_int, [_int]),
    "efgb_add_layernorm_workspace_bytes": (_sz, [_i64, _int]),
    "efgb_add_layernorm_forward": (_int, [_vp, _vp, _vp, _vp, _i64, _int, ctypes.c_float, _vp, _vp, _vp, _vp, _vp]),
    "efgb_add_layernorm_backward": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _int, _vp, _vp, _vp, _vp, _sz, _vp]),
    "efgb_colsum_workspace_bytes": (_sz, [_i64, _int]),
    "efgb_colsum": (_int, [_vp, _i64, _int, _vp, _vp, _sz, _vp]),
    "efgb_box_attn_forward": (_int, [_vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _int, _int, _vp, _vp]),
    "efgb_box_attn_fused_supported": (_int, [_int, _int, _int, _int]),
    "efgb_box_attn_fused_forward": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _i64, _i64, _int,
                                           _vp, _vp]),
    "efgb_box_attn_fused_backward": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _i64, _i64,
                                            _int, _vp, _vp, _vp, _vp]),
    "efgb_box_grid_softmax_forward": (_int, [_vp, _vp, _vp, _vp, _i64, _int, _int, _int, _int, _i64, _i64, _vp, _vp, _vp]),
    "efgb_box_grid_softmax_backward": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _int, _int, _int, _int, _i64, _i64, _vp, _vp,
                                             _vp]),
    "efgb_box_attn_backward": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _int, _int,
                                      _vp, _vp, _vp, _vp]),
}

_lock = threading.Lock()
_lib = None


def lib():
    """Return the loaded C-ABI library, loading it on first use.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    "efg_b200: %s is missing. The CUDA library is the only implementation of this "
                    "path (no CPU fallback); build it with `python -m efg_b200._build`." % LIB_PATH
                )
            handle = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)  # AttributeError if the .so is stale
                fn.restype = res
                fn.argtypes = args
            _lib = handle
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().efgb_last_error().decode("utf-8", "replace")
        raise RuntimeError("efg_b200 %s failed (code %d): %s" % (what, rc, msg))


def i32x3(values):
    a = (ctypes.c_int32 * 3)(*[int(v) for v in values])
    return a


def f32array(values):
    a = (ctypes.c_float * len(values))(*[float(v) for v in values])
    return a
