"""Build the sm_100a shared library in-tree: opencorr_b200/lib/libopencorr_b200.so.

nvcc cross-compiles without a GPU.  The library is a plain C-ABI .so (static cudart, no torch).
"""
import glob
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libopencorr_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libopencorr_b200.so")


def sources():
    csrc = os.path.join(_HERE, "csrc")
    return sorted(glob.glob(os.path.join(csrc, "*.cu"))), sorted(
        glob.glob(os.path.join(csrc, "*.cuh")) + glob.glob(os.path.join(csrc, "*.h"))
        + [os.path.join(_HERE, "..", "include", "opencorr_b200.h")])


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    cu, hdr = sources()
    return any(os.path.exists(p) and os.path.getmtime(p) > t for p in cu + hdr)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into one shared library (objects in parallel, then one link)."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cu, _ = sources()
    nvcc = [_nvcc()]
    # the image's CXX points at a gcc without OpenMP specs; nvcc only needs a host g++
    if os.path.exists("/usr/bin/g++"):
        nvcc += ["-ccbin", "/usr/bin/g++"]
    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])

    def compile_one(src):
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        subprocess.check_call(nvcc + compile_flags + ["-c", "-o", obj, src])
        return obj

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(cu), os.cpu_count() or 4)) as pool:
        objs = list(pool.map(compile_one, cu))
    subprocess.check_call(nvcc + ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", LIB_PATH] + objs)
    shutil.rmtree(obj_dir, ignore_errors=True)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
