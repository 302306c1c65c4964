"""Build the in-tree C-ABI library ``efg_b200/libefgb200.so`` with nvcc for sm_100a.

No torch headers are involved (the library is ATen-free), so a full rebuild takes seconds.
The resulting .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(CSRC, "build")
LIB_PATH = os.path.join(HERE, "libefgb200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the efg_b200 CUDA library cannot be built")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path, deps):
    h = hashlib.sha1()
    for p in [path] + deps:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(nvcc, src, deps, verbose, defines=(), tag=""):
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + tag + ".o")
    stamp = obj + ".sha1"
    dig = _digest(src, deps) + "".join(defines)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, ""
    cmd = [nvcc] + NVCC_FLAGS + list(defines) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    with open(stamp, "w") as f:
        f.write(dig)
    with open(obj + ".ptxas.log", "w") as f:
        f.write(r.stderr)
    return obj, r.stderr if verbose else ""


# Build variants for A/B measurements on the GPU box: name -> extra nvcc defines.  The default library is the
# product; a variant is only loaded when EFGB_LIB_VARIANT names it (efg_b200/_lib.py).
VARIANTS = {
    "pw8": ["-DEFGB_TC_PRODUCER_WARPS=8"],
    "pw16": ["-DEFGB_TC_PRODUCER_WARPS=16"],
    "fmma": ["-DEFGB_TC_FENCE_AT_MMA=1"],
    "abl1": ["-DEFGB_TC_ABLATE=1"],   # perf ablations of spconv_tc.cu (results invalid)
    "abl2": ["-DEFGB_TC_ABLATE=2"],
    "abl4": ["-DEFGB_TC_ABLATE=4"],
    "abl6": ["-DEFGB_TC_ABLATE=6"],
    "abl7": ["-DEFGB_TC_ABLATE=7"],
    "abl8": ["-DEFGB_TC_ABLATE=8"],
    "sb2": ["-DEFGB_TC_SB_MID=2"],
    "trace": ["-DEFGB_TC_TRACE=1"],    # pipeline time stamps of CTA 0 (scripts/trace_conv.py)
}


def build(verbose=False, force=False, variant=None):
    """Compile every .cu under csrc/ and link libefgb200.so (or libefgb200_<variant>.so). Returns the library path."""
    defines = tuple(VARIANTS[variant]) if variant else ()
    tag = "_" + variant if variant else ""
    lib_path = LIB_PATH[:-3] + tag + ".so"
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJ_DIR):
            os.remove(os.path.join(OBJ_DIR, f))
    deps = sorted(
        [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
        + [os.path.join(HERE, "..", "include", "efgb200.h")]
    )
    srcs = sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile_one(nvcc, s, deps, verbose, defines, tag), srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                sys.stderr.write(log)
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(lib_path) or os.path.getmtime(lib_path) < newest:
        # shared cudart: the library must use the SAME runtime instance as PyTorch (already loaded in the process)
        # so that stream capture (CUDA graphs) sees one consistent runtime; rpath is the fallback when torch is
        # not imported first
        cmd = [nvcc, "-shared", "-cudart", "shared", "-o", lib_path] + objs + [
            "-gencode", "arch=compute_100a,code=sm_100a", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return lib_path


if __name__ == "__main__":
    var = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")]
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv, variant=var[0] if var else None))
