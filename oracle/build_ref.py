"""ORACLE (test infrastructure): build the reference's own C++ CPU voxelizer into oracle/_ref/.

Only possible where /root/reference exists (the build container).  The resulting
``oracle/_ref/efg_ref_voxelize*.so`` is git-ignored but travels to the GPU box with gpurun, where
``load()`` imports the prebuilt file.  Sources are compiled where they lie; nothing is copied.
"""
import glob
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference/efg/operators/src/voxelize/voxelization_cpu.cpp"
NAME = "efg_ref_voxelize"


def build():
    if not os.path.exists(REF_SRC):
        return None
    existing = glob.glob(os.path.join(OUT, NAME + "*.so"))
    if existing:
        return existing[0]
    os.makedirs(OUT, exist_ok=True)
    from torch.utils.cpp_extension import load

    load(name=NAME, sources=[os.path.join(HERE, "ref_shim.cpp"), REF_SRC], extra_cflags=["-O2", "-std=c++17"],
         build_directory=OUT, verbose=False)
    existing = glob.glob(os.path.join(OUT, NAME + "*.so"))
    return existing[0] if existing else None


def load():
    """Import the prebuilt reference module, or return None when it was never built."""
    existing = glob.glob(os.path.join(OUT, NAME + "*.so"))
    if not existing:
        return None
    import torch  # noqa: F401  (libtorch symbols must be loaded first)

    spec = importlib.util.spec_from_file_location(NAME, existing[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ---- the reference's box-attention CUDA kernels, compiled for sm_100a: the GPU comparator -------------------------
BOX_SRC = "/root/reference/efg/operators/src/box_attn/box_attn.cu"
BOX_NAME = "efg_ref_box_attn"


def build_box_attn():
    """nvcc cross-compiles the reference's box_attn.cu (+ box_attn_kernel.cuh) unmodified for sm_100a; ~6 minutes
    (ATen headers).  Returns the .so path, or None where /root/reference is absent."""
    if not os.path.exists(BOX_SRC):
        return None
    existing = glob.glob(os.path.join(OUT, BOX_NAME + "*.so"))
    if existing:
        return existing[0]
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load

    load(name=BOX_NAME, sources=[os.path.join(HERE, "ref_box_attn_shim.cpp"), BOX_SRC],
         extra_cflags=["-O2", "-std=c++17", "-DWITH_CUDA"],
         extra_cuda_cflags=["-O2", "-std=c++17", "-DWITH_CUDA", "-gencode", "arch=compute_100a,code=sm_100a",
                            "--expt-relaxed-constexpr"],
         extra_include_paths=["/root/reference/efg/operators/src"], build_directory=OUT, verbose=False)
    existing = glob.glob(os.path.join(OUT, BOX_NAME + "*.so"))
    return existing[0] if existing else None


def load_box_attn():
    existing = glob.glob(os.path.join(OUT, BOX_NAME + "*.so"))
    if not existing:
        return None
    import torch  # noqa: F401

    spec = importlib.util.spec_from_file_location(BOX_NAME, existing[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ---- the reference's CPU BEV IoU (iou3d_cpu.cpp): pins oracle/iou3d.c ---------------------------------------------
IOU_SRC = "/root/reference/efg/operators/src/iou3d_nms/iou3d_cpu.cpp"
IOU_NAME = "efg_ref_iou3d"


def build_iou3d():
    if not os.path.exists(IOU_SRC):
        return None
    existing = glob.glob(os.path.join(OUT, IOU_NAME + "*.so"))
    if existing:
        return existing[0]
    os.makedirs(OUT, exist_ok=True)
    from torch.utils.cpp_extension import load

    # the file includes <cuda.h> / <cuda_runtime_api.h> and marks its Point methods __device__ (ignored by g++)
    load(name=IOU_NAME, sources=[os.path.join(HERE, "ref_iou3d_shim.cpp"), IOU_SRC],
         extra_cflags=["-O2", "-std=c++17", "-ffp-contract=off"],
         extra_include_paths=["/root/reference/efg/operators/src", "/usr/local/cuda/include"], build_directory=OUT, verbose=False)
    existing = glob.glob(os.path.join(OUT, IOU_NAME + "*.so"))
    return existing[0] if existing else None


def load_iou3d():
    existing = glob.glob(os.path.join(OUT, IOU_NAME + "*.so"))
    if not existing:
        return None
    import torch  # noqa: F401

    spec = importlib.util.spec_from_file_location(IOU_NAME, existing[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build())
    print(build_iou3d())
    if "--box-attn" in sys.argv:
        print(build_box_attn())
