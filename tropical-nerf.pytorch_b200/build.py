"""Build libtropical_b200.so (hand-written sm_100a kernels + the C ABI) in-tree with nvcc.

The library has no torch dependency: it is a plain CUDA shared object loaded through
ctypes (tropical/_native.py).  The float flags are part of the numerical contract with
the CPU checker (see csrc/common.cuh): no implicit FMA contraction, IEEE division and
square root, no flush-to-zero.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(OUT_DIR, "libtropical_b200.so")
SOURCES = ["net_kernels.cu", "complex.cu", "faces.cu", "grid_train.cu", "compat.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--expt-relaxed-constexpr",
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build the CUDA library")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "tropical_b200.h"))
    headers.append(os.path.abspath(__file__))
    nvcc = _nvcc()
    objs, jobs = [], []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        if force or _stale(obj, [path] + headers):
            jobs.append([nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj])
        objs.append(obj)
    rebuilt = bool(jobs)
    if jobs:   # the translation units are independent: compile them side by side (complex.cu alone takes minutes)
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=len(jobs)) as pool:
            list(pool.map(subprocess.check_call, jobs))
    if rebuilt or not os.path.exists(LIB):
        subprocess.check_call([nvcc, "-shared", "-o", LIB] + objs + ["-lcudart"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
