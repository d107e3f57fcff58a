"""TEST INFRASTRUCTURE ONLY -- builds the reference's own CUDA selective scan for sm_100a.

Compiles the UNMODIFIED sources where they lie under
``/root/reference/mamba-1p1p1/csrc/selective_scan`` (the ten files of the reference
``setup.py:118-129``, flags ``:131-150`` with the gencode swapped for
``compute_100a/sm_100a`` -- the reference ships sm_70/80/90 only, ``:102-108``) into
``oracle/_ref/selective_scan_cuda.so``.  No source is copied into this repo; only the
built ``.so`` (git-ignored, not gpurun-ignored) travels to the GPU box, where it is

* the GPU-side parity checker ("results must match the reference's own CUDA
  selective_scan") in ``tests/test_gpu_vs_reference_cuda.py``, and
* the same-box GPU competitor timed by ``bench.py`` (reported, never the product path).

Usage:  python oracle/build_ref.py [-j N]     (needs /root/reference; ~4 min on 8 cores)
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SRC = os.path.join(os.environ.get("FASTVIM_REFERENCE_ROOT", "/root/reference"),
                   "mamba-1p1p1", "csrc", "selective_scan")
SOURCES = ["selective_scan.cpp", "selective_scan_fwd_fp32.cu", "selective_scan_fwd_fp16.cu",
           "selective_scan_fwd_bf16.cu", "selective_scan_bwd_fp32_real.cu",
           "selective_scan_bwd_fp32_complex.cu", "selective_scan_bwd_fp16_real.cu",
           "selective_scan_bwd_fp16_complex.cu", "selective_scan_bwd_bf16_real.cu",
           "selective_scan_bwd_bf16_complex.cu"]


PYREF = os.path.join(OUT, "pyref")
REF_ROOT = os.environ.get("FASTVIM_REFERENCE_ROOT", "/root/reference")


def stage_python_reference() -> int:
    """Stages the reference's own PYTHON packages (``mamba-1p1p1/mamba_ssm`` and ``models``, .py files only, unmodified)
    into the git-ignored ``oracle/_ref/pyref/`` so that they exist on the GPU box, where /root/reference does not:
      * ``bench.py --impl reference`` / ``cpu_baseline`` then time the REFERENCE'S OWN CPU path (``selective_scan_ref`` /
        ``mamba_inner_ref`` through its ``VisionMamba``) instead of the oracle port (kind "reference");
      * ``tests/test_gpu_reference_model_over_shims.py`` drives the reference's own ``models/fastvim.py`` forward over the
        ``fastvim_b200/compat`` import shims on the GPU.
    Like the compiled ``selective_scan_cuda.so`` next to it this is a build artefact of the checker: never committed
    (``oracle/_ref/`` is git-ignored), never imported by the product package."""
    import shutil

    n = 0
    for sub in (os.path.join("mamba-1p1p1", "mamba_ssm"), "models"):
        src_root = os.path.join(REF_ROOT, sub)
        if not os.path.isdir(src_root):
            continue
        for dirpath, _dirs, files in os.walk(src_root):
            for f in files:
                if not f.endswith(".py"):
                    continue
                rel = os.path.relpath(os.path.join(dirpath, f), REF_ROOT)
                dst = os.path.join(PYREF, rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(os.path.join(dirpath, f)):
                    shutil.copyfile(os.path.join(dirpath, f), dst)
                n += 1
    print(f"[build_ref] staged {n} reference .py files under {PYREF}")
    return 0


def main(jobs: int) -> int:
    if os.path.isdir(os.path.join(REF_ROOT, "models")):
        stage_python_reference()
    if not os.path.isdir(SRC):
        print(f"[build_ref] {SRC} not present (GPU box or stripped container): nothing to do")
        return 0
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(os.path.join(OUT, "obj"), exist_ok=True)
    so = os.path.join(OUT, "selective_scan_cuda.so")
    newest_src = max(os.path.getmtime(os.path.join(SRC, f)) for f in os.listdir(SRC))
    if os.path.exists(so) and os.path.getmtime(so) > newest_src:
        print("[build_ref] up to date:", so)
        return 0
    inc = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{SRC}", f"-I{sysconfig.get_paths()['include']}"]
    common = ["-O3", "-std=c++17", "-DTORCH_EXTENSION_NAME=selective_scan_cuda",
              "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    nvcc_flags = ["-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
                  "-U__CUDA_NO_BFLOAT16_OPERATORS__", "-U__CUDA_NO_BFLOAT16_CONVERSIONS__",
                  "-U__CUDA_NO_BFLOAT162_OPERATORS__", "-U__CUDA_NO_BFLOAT162_CONVERSIONS__",
                  "--expt-relaxed-constexpr", "--expt-extended-lambda", "--use_fast_math", "-lineinfo",
                  "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-w"]

    def compile_one(f):
        obj = os.path.join(OUT, "obj", f + ".o")
        if f.endswith(".cu"):
            cmd = ["nvcc", "-c", os.path.join(SRC, f), "-o", obj] + common + nvcc_flags + inc
        else:
            cmd = ["g++", "-c", os.path.join(SRC, f), "-o", obj, "-fPIC", "-w"] + common + inc
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stderr[-4000:])
            raise RuntimeError(f"compile failed: {f}")
        return obj

    with ThreadPoolExecutor(jobs) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    libdirs = ce.library_paths("cuda")
    link = ["g++", "-shared", "-o", so] + objs + [f"-L{p}" for p in libdirs] + \
           [f"-Wl,-rpath,{p}" for p in libdirs] + \
           ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stderr[-4000:])
        return 1
    print("[build_ref] built", so)
    return 0


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("-j", type=int, default=max(1, (os.cpu_count() or 2) - 2))
    sys.exit(main(ap.parse_args().j))
