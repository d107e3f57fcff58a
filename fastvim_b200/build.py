"""Builds ``libfastvim_b200.so`` (the C-ABI library of include/fastvim_b200.h) in-tree.

Plain ``nvcc`` for sm_100a only -- no torch headers, no pybind: the library is a C ABI over
raw device pointers.  The built ``.so`` is git-ignored but travels to the GPU box with the
gpurun snapshot.  ``python -m fastvim_b200.build [--force] [--verbose]``.
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libfastvim_b200.so")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-I", os.path.join(ROOT, "include"),
              "-I", CSRC]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "fastvim_b200.h"))
    headers.append(os.path.abspath(__file__))
    srcs = sources()

    def compile_one(f):
        src, obj = os.path.join(CSRC, f), os.path.join(OBJ, f + ".o")
        if not force and not _stale(obj, [src] + headers):
            return obj, ""
        cmd = ["nvcc", "-c", src, "-o", obj] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else [])
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(f"nvcc failed on {f}:\n{r.stderr[-6000:]}")
        return obj, r.stderr

    with ThreadPoolExecutor(max(1, min(8, os.cpu_count() or 1))) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                sys.stderr.write(log)
    if force or _stale(LIB, objs):
        r = subprocess.run(["nvcc", "-shared", "-o", LIB] + objs +
                           ["-gencode", "arch=compute_100a,code=sm_100a"],
                           capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(f"link failed:\n{r.stderr[-4000:]}")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
