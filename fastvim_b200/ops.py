"""Tensor-level wrappers over the C ABI (one function per ``fv_*`` entry point).

These take CUDA ``torch.Tensor`` arguments, allocate outputs with torch, and pass raw
pointers / strides / the current CUDA stream to ``libfastvim_b200.so``.  No arithmetic
happens in Python.  Layout is token-major ``(B, L, D)`` (see include/fastvim_b200.h).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import FV_BF16, FV_F32, FV_POOL_MAX, FV_POOL_MEAN, fv_geom

Tensor = torch.Tensor


@dataclass(frozen=True)
class Geometry:
    """Sequence geometry of one mixer call (mirrors ``fv_geom``).

    The mixer sees a ``rows x cols`` token grid and pools over ``cols``
    (``mamba_simple_faster.py:287-297``).  ``rotated=True`` is the odd-layer case of
    ``models/fastvim.py:192-210``: the mixer's sequence is the column-major walk of a
    ``cols x rows`` row-major grid in memory, expressed here as strides instead of a copy.
    """
    outer: int
    pool: int
    inner: int = 1
    stride_outer: int = 0
    stride_pool: int = 1
    stride_inner: int = 0

    @staticmethod
    def grid(rows: int, cols: int, rotated: bool = False) -> "Geometry":
        if rotated:  # memory grid is (cols, rows) row-major; sequence t = r*cols + c lives at c*rows + r
            return Geometry(rows, cols, 1, 1, rows, 0)
        return Geometry(rows, cols, 1, cols, 1, 0)

    @property
    def L(self) -> int:
        return self.outer * self.pool * self.inner

    @property
    def Lp(self) -> int:
        return self.outer * self.inner

    def c_struct(self, batch: int, dim: int) -> fv_geom:
        return fv_geom(batch, dim, self.outer, self.pool, self.inner, self.stride_outer,
                       self.stride_pool, self.stride_inner)


def _dt(t: Tensor) -> int:
    if t.dtype == torch.float32:
        return FV_F32
    if t.dtype == torch.bfloat16:
        return FV_BF16
    raise TypeError(f"fastvim_b200 supports float32 and bfloat16 activations, got {t.dtype}")


def _p(t: Optional[Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(t: Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _check_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.FastVimLibraryError("fastvim_b200 ops need CUDA tensors (there is no CPU fallback)")


def _f32c(t: Optional[Tensor]) -> Optional[Tensor]:
    if t is None:
        return None
    return t.detach().to(torch.float32).contiguous()


def _tokmajor(t: Tensor, name: str) -> Tuple[int, int]:
    """(row stride, image stride) of a (B, L, D) view with unit channel stride."""
    if t.dim() != 3 or t.stride(2) != 1:
        raise ValueError(f"{name} must be (B, L, D) with contiguous channels, got strides {t.stride()}")
    return t.stride(1), t.stride(0)


def conv_pool_fwd(x: Tensor, geom: Geometry, conv_w: Tensor, conv_b: Optional[Tensor],
                  scale: float = 1.0, mode: str = "mean") -> Tensor:
    """K1.  x (B, L, D) -> pooled conv output u (2, B, Lp, D), [0]=forward, [1]=backward direction."""
    _check_cuda(x, conv_w)
    B, L, D = x.shape
    assert L == geom.L, (L, geom)
    ldx, bs = _tokmajor(x, "x")
    u = torch.empty((2, B, geom.Lp, D), device=x.device, dtype=x.dtype)
    g = geom.c_struct(B, D)
    _lib.call("fv_conv_pool_fwd", C.byref(g), _dt(x), _p(x), ldx, bs, _p(conv_w), _p(conv_b), float(scale),
              FV_POOL_MAX if mode == "max" else FV_POOL_MEAN, _p(u), _stream(x))
    return u


def scan_fwd(u: Tensor, xdbl: Tensor, geom: Geometry, dt_rank: int, d_state: int, dt_w: Tensor,
             dt_bias: Tensor, A: Tensor, a_is_log: bool = False) -> Tensor:
    """K2a.  u (2, B, Lp, D), xdbl (2, B*Lp, >=R+2N) -> s (2, B, Lp, D) fp32 (one plane per direction)."""
    _check_cuda(u, xdbl)
    _, B, Lp, D = u.shape
    assert u.is_contiguous() and xdbl.dim() == 3 and xdbl.stride(2) == 1 and xdbl.shape[1] == B * Lp
    assert xdbl.stride(0) == xdbl.shape[1] * xdbl.stride(1)
    s = torch.empty((2, B, Lp, D), device=u.device, dtype=torch.float32)
    g = geom.c_struct(B, D)
    _lib.call("fv_scan_fwd", C.byref(g), _dt(u), _p(u), _p(xdbl), xdbl.stride(1), dt_rank, d_state, _p(dt_w),
              _p(dt_bias), _p(A), int(a_is_log), _p(s), _stream(u))
    return s


def gate_fwd(x: Tensor, z: Tensor, s: Tensor, geom: Geometry, conv_w: Tensor, conv_b: Optional[Tensor],
             Dskip: Tensor, ln_w: Optional[Tensor], ln_b: Optional[Tensor], eps: float = 1e-5,
             out: Optional[Tensor] = None, stats: Optional[Tensor] = None) -> Tensor:
    """K2b.  -> y (B, L, D): LayerNorm((s[j] + D_f xc_f + D_b xc_b) / 2) * silu(z)."""
    _check_cuda(x, z, s)
    B, L, D = x.shape
    ldx, bs = _tokmajor(x, "x")
    ldz, bsz = _tokmajor(z, "z")
    assert (ldx, bs) == (ldz, bsz), "x and z must be the two halves of one in_proj output"
    y = out if out is not None else torch.empty((B, L, D), device=x.device, dtype=x.dtype)
    ldy, bsy = _tokmajor(y, "y")
    g = geom.c_struct(B, D)
    _lib.call("fv_gate_fwd", C.byref(g), _dt(x), _p(x), _p(z), ldx, bs, _p(s), _p(conv_w), _p(conv_b), _p(Dskip),
              _p(ln_w), _p(ln_b), float(eps), _p(y), ldy, bsy, _p(stats), _stream(x))
    return y


def norm_gate_apply(y: Tensor, z: Tensor, stats: Tensor, geom: Geometry, full_dim: int,
                    ln_w: Optional[Tensor], ln_b: Optional[Tensor], eps: float = 1e-5) -> Tensor:
    B, L, D = y.shape
    ldy, bsy = _tokmajor(y, "y")
    ldz, bsz = _tokmajor(z, "z")
    g = geom.c_struct(B, D)
    _lib.call("fv_norm_gate_apply", C.byref(g), _dt(y), int(full_dim), _p(y), ldy, bsy, _p(z), ldz, bsz,
              _p(stats), _p(ln_w), _p(ln_b), float(eps), _stream(y))
    return y


def add_norm_fwd(x: Tensor, residual: Optional[Tensor], weight: Tensor, bias: Optional[Tensor],
                 eps: float, is_rms: bool, want_residual: bool = True, want_stats: bool = False):
    """Fused (x + residual) -> fp32 residual_out, y = norm(residual_out).  x: (..., C)."""
    _check_cuda(x, weight)
    shape = x.shape
    x2 = x.reshape(-1, shape[-1])
    if x2.stride(1) != 1:
        x2 = x2.contiguous()
    rows, cols = x2.shape
    if residual is not None:
        residual = residual.reshape(rows, cols)
        assert residual.dtype == torch.float32 and residual.is_contiguous()
    y = torch.empty((rows, cols), device=x.device, dtype=x.dtype)
    res_out = torch.empty((rows, cols), device=x.device, dtype=torch.float32) if want_residual else None
    mean = torch.empty(rows, device=x.device, dtype=torch.float32) if (want_stats and not is_rms) else None
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32) if want_stats else None
    _lib.call("fv_add_norm_fwd", _dt(x), rows, cols, _p(x2), x2.stride(0), _p(residual), _p(weight), _p(bias),
              float(eps), int(is_rms), _p(y), cols, _p(res_out), _p(mean), _p(rstd), _stream(x))
    y = y.reshape(shape)
    res_out = res_out.reshape(shape) if res_out is not None else None
    return y, res_out, mean, rstd


def selective_scan_fwd(u: Tensor, delta: Tensor, A: Tensor, B: Tensor, Cm: Tensor, D: Optional[Tensor],
                       z: Optional[Tensor], delta_bias: Optional[Tensor], delta_softplus: bool,
                       want_last_state: bool = False):
    """Operator-API scan on (batch, dim, L), L contiguous; B, C: (batch, groups, N, L)."""
    _check_cuda(u, delta, A, B, Cm)
    batch, dim, L = u.shape
    groups, N = B.shape[1], B.shape[2]
    for t in (u, delta, B, Cm, z):
        assert t is None or t.is_contiguous()
    out = torch.empty_like(u)
    last = torch.empty((batch, dim, N), device=u.device, dtype=torch.float32) if want_last_state else None
    _lib.call("fv_selective_scan_fwd", _dt(u), batch, dim, L, N, groups, _p(u), _p(delta), _p(A), _p(B), _p(Cm),
              _p(D), _p(z), _p(delta_bias), int(delta_softplus), _p(out), _p(last), _stream(u))
    return out, last
