"""ctypes binding of ``libfastvim_b200.so`` (the C ABI declared in include/fastvim_b200.h).

PyTorch is plumbing here: it owns device memory and streams; every hot-path op goes through
the ``fv_*`` entry points below.  There is no CPU or PyTorch fallback: if the library is not
built, ``lib()`` raises, and every op in this package fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libfastvim_b200.so")

FV_F32, FV_BF16 = 0, 1
FV_POOL_MEAN, FV_POOL_MAX = 0, 1


class fv_geom(C.Structure):
    _fields_ = [("batch", C.c_int32), ("dim", C.c_int32), ("outer", C.c_int32), ("pool", C.c_int32),
                ("inner", C.c_int32), ("tok_stride_outer", C.c_int64), ("tok_stride_pool", C.c_int64),
                ("tok_stride_inner", C.c_int64)]


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float
_G = C.POINTER(fv_geom)

# name -> argtypes; every symbol declared in include/fastvim_b200.h must be listed here
# (tests/test_abi.py cross-checks this table against the header).
SIGNATURES = {
    "fv_conv_pool_fwd": [_G, _I, _P, _L, _L, _P, _P, _F, _I, _P, _P],
    "fv_scan_fwd": [_G, _I, _P, _P, _L, _I, _I, _P, _P, _P, _I, _P, _P],
    "fv_gate_fwd": [_G, _I, _P, _P, _L, _L, _P, _P, _P, _P, _P, _P, _F, _P, _L, _L, _P, _P],
    "fv_norm_gate_apply": [_G, _I, _I, _P, _L, _L, _P, _L, _L, _P, _P, _P, _F, _P],
    "fv_block_fwd_supported": [_G, _I, _I, _I],
    "fv_block_pack_xproj_bytes": [_I, _I],
    "fv_block_pack_xproj": [_I, _I, _P, _P, _P],
    "fv_block_fwd": [_G, _I, _P, _P, _L, _L, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _F, _F, _P, _L, _L,
                     _P, _P, _P, _P, _P, _P],
    "fv_block_fwd_saves_v": [_G, _I, _I, _I],
    "fv_gate_bwd_v_supported": [_G, _I],
    "fv_gate_bwd_v": [_G, _I, _P, _P, _L, _L, _P, _L, _L, _P, _P, _F, _P, _P, _P, _P, _P, _P],
    "fv_add_norm_fwd": [_I, _L, _I, _P, _L, _P, _P, _P, _F, _I, _P, _L, _P, _P, _P, _P],
    "fv_selective_scan_fwd": [_I, _I, _I, _L, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P],
    "fv_gemm_bf16_tn": [_L, _I, _I, _P, _L, _P, _L, _P, _P, _L, _P],
    "fv_gemm_supported": [_L, _I, _I],
    "fv_set_pdl": [_I],
    "fv_set_pdl_all": [_I],
    "fv_conv_pool_w_supported": [_G, _I],
    "fv_conv_pool_w_fwd": [_G, _I, _P, _L, _L, _P, _P, _F, _I, _P, _P, _P, _P],
    "fv_gate_w_fwd": [_G, _I, _P, _P, _L, _L, _P, _P, _P, _F, _P, _L, _L, _P],
    "fv_block_fwd_signal_supported": [_G, _I, _I, _I],
    "fv_block_fwd_signal": [_G, _I, _P, _P, _L, _L, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _F, _F, _P, _L, _L,
                            _P, _I, _P],
    "fv_gemm_out_norm_flow": [_L, _I, _I, _P, _L, _P, _L, _P, _L, _P, _P, _F, _P, _L, _P, _I, _I, _P],
    "fv_gemm_out_norm_supported": [_L, _I, _I],
    "fv_gemm_out_norm": [_L, _I, _I, _P, _L, _P, _L, _P, _L, _P, _P, _F, _P, _L, _P],
    "fv_gemm_bf16": [_L, _I, _L, _I, _P, _L, _I, _P, _L, _I, _P, _L, _I, _P],
    "fv_gemm_bf16_splits": [_L, _I, _L],
    "fv_gemm_bf16_batched": [_I, _L, _I, _L, _I, _P, _L, _L, _I, _P, _L, _L, _I, _P, _L, _L, _I, _P],
    "fv_ln_gate_fwd": [_I, _L, _I, _P, _L, _P, _L, _P, _P, _F, _P, _L, _P],
    "fv_peer_header_bytes": [],
    "fv_peer_sum_f32": [_I, _I, _P, _L, _L, _P, _P, _P],
    "fv_peer_copy2d": [_I, _I, _P, _I, _I, _L, _P, _L, _P, _L, _P, _P],
    "fv_peer_error": [_P],
    "fv_patchify_supported": [_I, _I, _I, _I, _I],
    "fv_patchify": [_I, _I, _I, _I, _I, _I, _I, _P, _P, _P],
    "fv_causal_conv1d_fwd": [_I, _I, _I, _L, _P, _L, _L, _P, _P, _I, _P, _P],
    "fv_pool_bdl_fwd": [_I, _I, _I, _I, _I, _I, _P, _I, _F, _P, _P],
    "fv_bcast_skip_bdl_fwd": [_I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    "fv_selective_scan_bwd_workspace_bytes": [_I, _I, _L, _I],
    "fv_selective_scan_bwd": [_I, _I, _I, _L, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P,
                              _P, _P, _P, _L, _P],
    "fv_causal_conv1d_bwd": [_I, _I, _I, _L, _P, _L, _L, _P, _P, _I, _P, _P, _L, _L, _P, _P, _P],
    "fv_rowdot_bdl": [_I, _I, _I, _L, _P, _P, _P, _P],
    "fv_bwd_tiles_per_group": [_G, _I],
    "fv_gate_bwd": [_G, _I, _P, _P, _L, _L, _P, _L, _L, _P, _P, _P, _P, _P, _P, _F, _P, _P, _P, _P, _P, _P, _P],
    "fv_gate_bwd_stream_supported": [_G, _L, _L],
    "fv_gate_bwd_stream": [_G, _I, _P, _P, _L, _L, _P, _L, _L, _P, _P, _P, _P, _P, _P, _F, _P, _P, _P, _P, _P, _P, _P, _P],
    "fv_scan_bwd_planes": [_G],
    "fv_scan_bwd_short_supported": [_G, _I],
    "fv_scan_bwd_short": [_G, _I, _I, _P, _P, _L, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P],
    "fv_scan_bwd": [_G, _I, _I, _P, _P, _L, _I, _I, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P],
    "fv_reduce_planes": [_I, _P, _I, _L, _P, _P],
    "fv_conv_pool_bwd": [_G, _I, _P, _L, _L, _P, _P, _P, _P, _P, _F, _I, _P, _P, _P, _P, _P],
    "fv_add_norm_bwd": [_I, _L, _I, _P, _L, _P, _P, _P, _F, _I, _P, _L, _P, _P, _P, _P],
}

_lib = None


class FastVimLibraryError(RuntimeError):
    pass


def lib():
    """The loaded C-ABI library.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FastVimLibraryError(
            f"{LIB_PATH} not found: build it with `python -m fastvim_b200.build` "
            "(fastvim_b200 has no CPU/PyTorch fallback for its CUDA kernels)")
    l = C.CDLL(LIB_PATH)
    l.fv_last_error.restype = C.c_char_p
    l.fv_last_error.argtypes = []
    l.fv_version.restype = C.c_int
    l.fv_launch_count.restype = C.c_int64
    l.fv_reset_launch_count.restype = None
    for name, args in SIGNATURES.items():
        fn = getattr(l, name)
        fn.argtypes = args
        fn.restype = C.c_int64 if name.endswith("_bytes") else C.c_int
    _lib = l
    return l


_profile = None


def set_profile(records) -> None:
    """bench.py's per-kernel timer: when `records` is a list, every C-ABI launch is bracketed by
    CUDA events on the current stream and ``(name, start, stop)`` is appended to it."""
    global _profile
    _profile = records


def call(name: str, *args) -> None:
    l = lib()
    if _profile is not None:
        import torch

        n0 = int(l.fv_launch_count())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(l, name)(*args)
        e1.record()
        tag = name
        if name in ("fv_gemm_bf16_tn", "fv_gemm_out_norm", "fv_gemm_out_norm_flow", "fv_gemm_bf16"):   # several GEMM shapes share one entry point: tag with (M, N, K)
            tag = "%s[%dx%dx%d]" % (name, int(args[0]), int(args[1]), int(args[2]))
        _profile.append((tag, e0, e1, int(l.fv_launch_count()) - n0))
    else:
        rc = getattr(l, name)(*args)
    if rc != 0:
        raise FastVimLibraryError(f"{name} failed (rc={rc}): {l.fv_last_error().decode()}")


def launch_count() -> int:
    return int(lib().fv_launch_count())


def reset_launch_count() -> None:
    lib().fv_reset_launch_count()


class pdl_all:
    """``with pdl_all():`` -- programmatic dependent launch for every kernel of the library inside the block (process-wide
    switch, restored on exit).  ``FASTVIM_TRAIN_PDL=1`` makes the training path (``fastvim_b200.autograd``) use it."""

    def __init__(self, on: bool = True):
        self.on = on

    def __enter__(self):
        self.prev = int(lib().fv_set_pdl_all(1 if self.on else 0))
        return self

    def __exit__(self, *exc):
        lib().fv_set_pdl_all(self.prev)
        return False
