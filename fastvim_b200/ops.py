"""Tensor-level wrappers over the C ABI (one function per ``fv_*`` entry point).

These take CUDA ``torch.Tensor`` arguments, allocate outputs with torch, and pass raw
pointers / strides / the current CUDA stream to ``libfastvim_b200.so``.  No arithmetic
happens in Python.  Layout is token-major ``(B, L, D)`` (see include/fastvim_b200.h).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import FV_BF16, FV_F32, FV_POOL_MAX, FV_POOL_MEAN, fv_geom

Tensor = torch.Tensor

# streaming gate backward (statistics pass + apply pass, csrc/gate_bwd_stream.cu).  Parity-green but measured SLOWER than
# the tiled kernel (408 vs 386 us at FastVim-B, 260 vs 191 us at FastVim-T: DESIGN.md 3b), so it is opt-in ("1").
GATE_BWD_STREAM = os.environ.get("FASTVIM_GATE_BWD_STREAM", "0") == "1"
# short pooled sequences (Lp <= 16): fv_scan_bwd_short (two states per thread, dt_proj by GEMM); "0" = previous kernel
SCAN_BWD_SHORT = os.environ.get("FASTVIM_SCAN_BWD_SHORT", "1") != "0"
# split-K wgrad GEMMs: "1" = the K splits add into one fp32 output with TMA reduction stores (no workspace planes, no
# fv_reduce_planes pass; summation order over the splits not fixed), "0" = deterministic planes + reduction
WGRAD_ACC = os.environ.get("FASTVIM_WGRAD_ACC", "1") != "0"
WGRAD_ACC_MAX_SPLITS = 4


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
    cur = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.FastVimLibraryError("fastvim_b200 ops need CUDA tensors (there is no CPU fallback)")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            # the C ABI launches on the given stream of the CURRENT device (no device guard inside the library)
            raise _lib.FastVimLibraryError(
                f"tensor on cuda:{t.device.index} but the current device is cuda:{cur}; wrap the call in "
                "torch.cuda.device(tensor.device)")


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


def conv_pool_w_supported(geom: Geometry, batch: int, dim: int, dtype: torch.dtype) -> bool:
    """Channel layouts (inner > 1), bf16: the staged conv + pool kernel that also writes the D-skip term."""
    if dtype != torch.bfloat16:
        return False
    g = geom.c_struct(batch, dim)
    return bool(_lib.lib().fv_conv_pool_w_supported(C.byref(g), FV_BF16))


def conv_pool_w_fwd(x: Tensor, geom: Geometry, conv_w: Tensor, conv_b: Optional[Tensor], Dskip: Tensor,
                    scale: float = 1.0, mode: str = "mean", want_w: bool = True):
    """K1 for channel layouts.  x (B, L, D) bf16 -> (u (2, B, Lp, D), w (B, L, D) | None): pooled conv output and the
    D-skip term w = (D_f xc_f + D_b xc_b) / 2 of every token (memory-row order), for ``gate_w_fwd``."""
    _check_cuda(x, conv_w)
    B, L, D = x.shape
    assert L == geom.L and x.dtype == torch.bfloat16
    ldx, bs = _tokmajor(x, "x")
    u = torch.empty((2, B, geom.Lp, D), device=x.device, dtype=x.dtype)
    w = torch.empty((B, L, D), device=x.device, dtype=x.dtype) if want_w else None
    g = geom.c_struct(B, D)
    _lib.call("fv_conv_pool_w_fwd", C.byref(g), FV_BF16, _p(x), ldx, bs, _p(conv_w), _p(conv_b), float(scale),
              FV_POOL_MAX if mode == "max" else FV_POOL_MEAN, _p(Dskip), _p(u), _p(w), _stream(x))
    return u, w


def gate_w_fwd(w: Tensor, z: Tensor, s: Tensor, geom: Geometry, ln_w: Optional[Tensor], ln_b: Optional[Tensor],
               eps: float = 1e-5) -> Tensor:
    """K2b from the saved D-skip term: y (B, L, D) = LayerNorm(w + (s_f[j] + s_b[j]) / 2) * silu(z)."""
    _check_cuda(w, z, s)
    B, L, D = w.shape
    assert w.is_contiguous() and w.dtype == torch.bfloat16 and s.dtype == torch.float32 and s.is_contiguous()
    ldz, bsz = _tokmajor(z, "z")
    y = torch.empty((B, L, D), device=w.device, dtype=w.dtype)
    g = geom.c_struct(B, D)
    _lib.call("fv_gate_w_fwd", C.byref(g), FV_BF16, _p(w), _p(z), ldz, bsz, _p(s), _p(ln_w), _p(ln_b), float(eps), _p(y),
              y.stride(1), y.stride(0), _stream(w))
    return y


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


def block_fwd_supported(geom: Geometry, batch: int, dim: int, dtype: torch.dtype, dt_rank: int, d_state: int) -> bool:
    """True when the fused one-launch block interior (``fv_block_fwd``) handles this configuration."""
    if dtype != torch.bfloat16:
        return False
    g = geom.c_struct(batch, dim)
    return bool(_lib.lib().fv_block_fwd_supported(C.byref(g), FV_BF16, int(dt_rank), int(d_state)))


def block_fwd_signal_supported(geom: Geometry, batch: int, dim: int, dt_rank: int, d_state: int) -> bool:
    g = geom.c_struct(batch, dim)
    return bool(_lib.lib().fv_block_fwd_signal_supported(C.byref(g), FV_BF16, int(dt_rank), int(d_state)))


def block_pack_xproj(xproj_w: Tensor) -> Optional[Tensor]:
    """(2, R+2N, D) bf16 x_proj weights -> MMA-fragment order for ``block_fwd`` (None when D % 64 != 0)."""
    _check_cuda(xproj_w)
    _, ncols, D = xproj_w.shape
    nbytes = int(_lib.lib().fv_block_pack_xproj_bytes(D, ncols))
    if nbytes == 0:
        return None
    assert xproj_w.dtype == torch.bfloat16 and xproj_w.is_contiguous()
    packed = torch.empty(nbytes, device=xproj_w.device, dtype=torch.uint8)
    _lib.call("fv_block_pack_xproj", D, ncols, _p(xproj_w), _p(packed), _stream(xproj_w))
    return packed


def block_fwd(x: Tensor, z: Tensor, geom: Geometry, conv_w: Tensor, conv_b: Optional[Tensor], xproj_w: Tensor,
              dt_w: Tensor, dt_bias: Tensor, A: Tensor, Dskip: Tensor, ln_w: Optional[Tensor], ln_b: Optional[Tensor],
              eps: float, scale: float, dt_rank: int, d_state: int, a_is_log: bool = True, save: bool = False,
              xproj_w_packed: Optional[Tensor] = None, save_v: Optional[bool] = None, signal=None):
    """K-fused.  x, z (B, L, D) bf16 halves of the in_proj output -> y (B, L, D) bf16 (the out_proj input).
    xproj_w (2, R+2N, D) bf16, dt_w (2, D, R) fp32.  With ``save`` also returns the pooled intermediates
    (u (2, B, Lp, D) bf16, xdbl (2, B*Lp, R+2N) bf16, s (2, B, Lp, D) fp32) the backward kernels need.  With ``save_v``
    (implies ``save``) returns (y, u, xdbl, s, v): when the cluster kernel serves the configuration, v (B, L, D) bf16 --
    the pre-norm merged value, for ``gate_bwd_v`` -- and the fp32 dt_proj pre-activation (2, B*Lp, D), for ``scan_bwd``, are
    saved INSTEAD of s (s is None; the last element is the pair (v, pre)); otherwise the last element is None."""
    want5 = save_v is not None     # the 5-tuple form (training path) whenever the caller names save_v
    save_v = bool(save_v)
    save = save or want5
    _check_cuda(x, z, xproj_w, dt_w)
    B, L, D = x.shape
    assert L == geom.L and x.dtype == torch.bfloat16 and xproj_w.dtype == torch.bfloat16 and dt_w.dtype == torch.float32
    ldx, bs = _tokmajor(x, "x")
    assert _tokmajor(z, "z") == (ldx, bs), "x and z must be the two halves of one in_proj output"
    assert xproj_w.is_contiguous() and dt_w.is_contiguous()
    if xproj_w_packed is None:
        xproj_w_packed = block_pack_xproj(xproj_w)
    y = torch.empty((B, L, D), device=x.device, dtype=x.dtype)
    ncols = dt_rank + 2 * d_state
    u = xdbl = s = v = pre = None
    g = geom.c_struct(B, D)
    if save:
        u = torch.empty((2, B, geom.Lp, D), device=x.device, dtype=x.dtype)
        xdbl = torch.empty((2, B * geom.Lp, ncols), device=x.device, dtype=x.dtype)
        if save_v and _lib.lib().fv_block_fwd_saves_v(C.byref(g), FV_BF16, int(dt_rank), int(d_state)):
            # the streaming gate backward works from the saved pre-norm value; the per-direction scan outputs are not needed
            v = torch.empty((B, L, D), device=x.device, dtype=x.dtype)
            pre = torch.empty((2, B * geom.Lp, D), device=x.device, dtype=torch.float32)   # dt_proj pre-activation
        else:
            s = torch.empty((2, B, geom.Lp, D), device=x.device, dtype=torch.float32)
    if signal is not None:
        # inference dataflow form: publish each image's completion for fv_gemm_out_norm_flow (signal = (sync, launch_index))
        assert not save
        sync, launch_index = signal
        assert sync.dtype == torch.int32 and sync.numel() >= B + 1 and sync.is_contiguous()
        _lib.call("fv_block_fwd_signal", C.byref(g), FV_BF16, _p(x), _p(z), ldx, bs, _p(conv_w), _p(conv_b), _p(xproj_w),
                  _p(xproj_w_packed), _p(dt_w), _p(dt_bias), _p(A), int(a_is_log), int(dt_rank), int(d_state), _p(Dskip),
                  _p(ln_w), _p(ln_b), float(eps), float(scale), _p(y), y.stride(1), y.stride(0),
                  C.c_void_p(sync.data_ptr() + 4), int(launch_index) + 1, _stream(x))
        return y
    _lib.call("fv_block_fwd", C.byref(g), FV_BF16, _p(x), _p(z), ldx, bs, _p(conv_w), _p(conv_b), _p(xproj_w), _p(xproj_w_packed),
              _p(dt_w), _p(dt_bias), _p(A), int(a_is_log), int(dt_rank), int(d_state), _p(Dskip), _p(ln_w), _p(ln_b), float(eps),
              float(scale), _p(y), y.stride(1), y.stride(0), _p(u), _p(xdbl), _p(s), _p(v), _p(pre), _stream(x))
    if want5:
        return y, u, xdbl, s, (None if v is None else (v, pre))
    return (y, u, xdbl, s) if save else y


def norm_gate_apply(y: Tensor, z: Tensor, stats: Tensor, geom: Geometry, full_dim: int,
                    ln_w: Optional[Tensor], ln_b: Optional[Tensor], eps: float = 1e-5) -> Tensor:
    B, L, D = y.shape
    ldy, bsy = _tokmajor(y, "y")
    ldz, bsz = _tokmajor(z, "z")
    g = geom.c_struct(B, D)
    _lib.call("fv_norm_gate_apply", C.byref(g), _dt(y), int(full_dim), _p(y), ldy, bsy, _p(z), ldz, bsz,
              _p(stats), _p(ln_w), _p(ln_b), float(eps), _stream(y))
    return y


def ln_gate_fwd(v: Tensor, z: Tensor, ln_w: Optional[Tensor], ln_b: Optional[Tensor], eps: float) -> Tensor:
    """(rows, D) v, z (unit channel stride, any row stride) -> y (rows, D) = LayerNorm(v) * silu(z)."""
    _check_cuda(v, z)
    rows, D = v.shape
    assert v.stride(1) == 1 and z.stride(1) == 1 and z.shape == v.shape and z.dtype == v.dtype
    y = torch.empty((rows, D), device=v.device, dtype=v.dtype)
    _lib.call("fv_ln_gate_fwd", _dt(v), rows, D, _p(v), v.stride(0), _p(z), z.stride(0), _p(ln_w), _p(ln_b), float(eps),
              _p(y), D, _stream(v))
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


# --------------------------------------------------------------------------- tcgen05 GEMMs
def gemm_supported(M: int, N: int, K: int) -> bool:
    """Shapes the tcgen05 GEMM handles (K % 64 == 0, N % 64 == 0, the CTA's W block fits shared memory)."""
    return bool(_lib.lib().fv_gemm_supported(int(M), int(N), int(K)))


def gemm_bf16_tn(a: Tensor, w: Tensor, out: Optional[Tensor] = None, bias: Optional[Tensor] = None) -> Tensor:
    """a (..., K) bf16, w (N, K) bf16 [, bias (N)] -> (..., N) bf16 = a @ w.T + bias on the tcgen05 tensor cores
    (fp32 accumulate, bias added in fp32 before the bf16 rounding)."""
    _check_cuda(a, w)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and w.dim() == 2 and w.stride(1) == 1
    K = a.shape[-1]
    a2 = a.reshape(-1, K)
    if a2.stride(1) != 1:
        a2 = a2.contiguous()
    M, N = a2.shape[0], w.shape[0]
    c = out if out is not None else torch.empty((M, N), device=a.device, dtype=a.dtype)
    bias32 = _f32c(bias)   # keep the converted copy alive until after the launch (_p only takes the address)
    _lib.call("fv_gemm_bf16_tn", M, N, K, _p(a2), a2.stride(0), _p(w), w.stride(0), _p(bias32), _p(c), c.stride(0),
              _stream(a))
    return c.reshape(*a.shape[:-1], N)


def gemm_bf16(a: Tensor, b: Tensor, a_mn: bool = False, b_mn: bool = False, out_f32: bool = False,
              out: Optional[Tensor] = None, splits: Optional[int] = None) -> Tensor:
    """C (Mo, No) = op(a) @ op(b).T on the tcgen05 tensor cores (csrc/gemm_tc2.cu), bf16 operands, fp32 accumulate.

    ``a_mn=False``: a is (Mo, K); ``a_mn=True``: a is (K, Mo) and is used transposed WITHOUT a copy (MN-major UMMA
    operand).  ``b`` likewise with No.  ``out_f32``: fp32 result (split-K over the reduction when the output has few
    tiles: the partial planes are added by ``fv_reduce_planes``), else bf16.
        dgrad  dX = dY @ W        -> gemm_bf16(dY, W, b_mn=True)
        wgrad  dW = dY.T @ X      -> gemm_bf16(dY, X, a_mn=True, b_mn=True, out_f32=True)
    """
    _check_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.dim() == 2 and b.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1
    K, Mo = (a.shape[0], a.shape[1]) if a_mn else (a.shape[1], a.shape[0])
    Kb, No = (b.shape[0], b.shape[1]) if b_mn else (b.shape[1], b.shape[0])
    assert K == Kb, f"reduction lengths differ: {K} vs {Kb}"
    l = _lib.lib()
    if not out_f32:
        c = out if out is not None else torch.empty((Mo, No), device=a.device, dtype=torch.bfloat16)
        assert c.dtype == torch.bfloat16 and c.stride(1) == 1
        _lib.call("fv_gemm_bf16", Mo, No, K, int(a_mn), _p(a), a.stride(0), int(b_mn), _p(b), b.stride(0), FV_BF16, _p(c),
                  c.stride(0), 1, _stream(a))
        return c
    if splits is None:
        splits = int(l.fv_gemm_bf16_splits(Mo, No, K))
    c = out if out is not None else torch.empty((Mo, No), device=a.device, dtype=torch.float32)
    assert c.dtype == torch.float32 and c.is_contiguous()
    if splits == 1:
        _lib.call("fv_gemm_bf16", Mo, No, K, int(a_mn), _p(a), a.stride(0), int(b_mn), _p(b), b.stride(0), FV_F32, _p(c), No, 1,
                  _stream(a))
        return c
    # Few splits (FastVim-B: 2 / 4): every split adds into the one zeroed output with a TMA reduction store -- no planes, no
    # second pass (33.2 -> 32.6 ms per training step).  Many splits onto a small output (FastVim-T: 24 / 37 splits onto
    # <= 590 KB) contend in L2 and measured slower (11.1 vs 10.8 ms), so those keep the deterministic planes.
    if WGRAD_ACC and splits <= WGRAD_ACC_MAX_SPLITS:
        c.zero_()
        _lib.call("fv_gemm_bf16", Mo, No, K, int(a_mn), _p(a), a.stride(0), int(b_mn), _p(b), b.stride(0), 2, _p(c), No,
                  splits, _stream(a))
        return c
    ws = torch.empty((splits, Mo, No), device=a.device, dtype=torch.float32)
    _lib.call("fv_gemm_bf16", Mo, No, K, int(a_mn), _p(a), a.stride(0), int(b_mn), _p(b), b.stride(0), FV_F32, _p(ws), No,
              splits, _stream(a))
    _lib.call("fv_reduce_planes", FV_F32, _p(ws), splits, Mo * No, _p(c), _stream(a))
    return c


def gemm_bf16_batched(a: Tensor, b: Tensor, a_mn: bool = False, b_mn: bool = False, out: Optional[Tensor] = None,
                      out_f32: bool = False) -> Tensor:
    """``nbatch`` independent products ``op(a[i]) @ op(b[i]).T`` in one launch (``fv_gemm_bf16_batched``).  a, b: 3-D bf16 with
    unit inner stride (views with padded row / batch pitches allowed).  bf16 result, or fp32 (``out_f32``) accumulated over
    K splits into a zeroed buffer with TMA reduction stores."""
    _check_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.dim() == 3 and b.dim() == 3
    assert a.stride(2) == 1 and b.stride(2) == 1 and a.shape[0] == b.shape[0]
    nb = a.shape[0]
    K, Mo = (a.shape[1], a.shape[2]) if a_mn else (a.shape[2], a.shape[1])
    Kb, No = (b.shape[1], b.shape[2]) if b_mn else (b.shape[2], b.shape[1])
    assert K == Kb
    if out_f32:
        c = out if out is not None else torch.empty((nb, Mo, No), device=a.device, dtype=torch.float32)
        assert c.dtype == torch.float32 and c.stride(2) == 1
        c.zero_()
        splits = int(_lib.lib().fv_gemm_bf16_splits(Mo, No, K))
        splits = max(1, min(splits, max(1, int(_lib.lib().fv_gemm_bf16_splits(Mo * nb, No, K)))))   # nb batches share the SMs
        _lib.call("fv_gemm_bf16_batched", nb, Mo, No, K, int(a_mn), _p(a), a.stride(1), a.stride(0), int(b_mn), _p(b), b.stride(1),
                  b.stride(0), 2, _p(c), c.stride(1), c.stride(0), splits, _stream(a))
        return c
    c = out if out is not None else torch.empty((nb, Mo, No), device=a.device, dtype=torch.bfloat16)
    assert c.dtype == torch.bfloat16 and c.stride(2) == 1
    _lib.call("fv_gemm_bf16_batched", nb, Mo, No, K, int(a_mn), _p(a), a.stride(1), a.stride(0), int(b_mn), _p(b), b.stride(1),
              b.stride(0), FV_BF16, _p(c), c.stride(1), c.stride(0), 1, _stream(a))
    return c


def gemm_bf16_ok(*ts: Tensor) -> bool:
    """Operands the general tcgen05 GEMM accepts: CUDA bf16 2-D, unit inner stride, 16-byte aligned rows."""
    for t in ts:
        if not (t.is_cuda and t.dtype == torch.bfloat16 and t.dim() == 2 and t.stride(1) == 1 and t.stride(0) % 8 == 0
                and t.data_ptr() % 16 == 0):
            return False
    return True


def gemm_out_norm_supported(M: int, N: int, K: int) -> bool:
    return bool(_lib.lib().fv_gemm_out_norm_supported(int(M), int(N), int(K)))


def gemm_out_norm(a: Tensor, w: Tensor, residual: Tensor, norm_w: Tensor, eps: float, want_residual: bool = True,
                  inplace: bool = False, flow=None) -> Tuple[Tensor, Optional[Tensor]]:
    """out_proj + residual add + RMSNorm in one launch (``fv_gemm_out_norm``): a (..., K) bf16, w (N, K) bf16,
    residual (..., N) fp32 -> (y (..., N) bf16 = rmsnorm(residual + a @ w.T) * norm_w, new residual fp32 | None)."""
    _check_cuda(a, w, residual)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and residual.dtype == torch.float32
    K, N = a.shape[-1], w.shape[0]
    a2 = a.reshape(-1, K)
    if a2.stride(1) != 1:
        a2 = a2.contiguous()
    M = a2.shape[0]
    r2 = residual.reshape(M, N)
    assert r2.is_contiguous() and w.stride(1) == 1
    y = torch.empty((M, N), device=a.device, dtype=torch.bfloat16)
    res_out = None
    if want_residual:
        res_out = r2 if inplace else torch.empty((M, N), device=a.device, dtype=torch.float32)
    nw = _f32c(norm_w)
    if flow is not None:   # (sync, rows_per_flag, launch_index): `a` is being written by block_fwd(..., signal=) right now
        sync, rows_per_flag, launch_index = flow
        _lib.call("fv_gemm_out_norm_flow", M, N, K, _p(a2), a2.stride(0), _p(w), w.stride(0), _p(r2), r2.stride(0),
                  _p(res_out), _p(nw), float(eps), _p(y), y.stride(0), _p(sync), int(rows_per_flag), int(launch_index),
                  _stream(a))
    else:
        _lib.call("fv_gemm_out_norm", M, N, K, _p(a2), a2.stride(0), _p(w), w.stride(0), _p(r2), r2.stride(0), _p(res_out),
                  _p(nw), float(eps), _p(y), y.stride(0), _stream(a))
    shp = tuple(a.shape[:-1]) + (N,)
    return y.view(shp), (None if res_out is None else res_out.view(shp))


def x_proj(u: Tensor, x_w: Tensor, use_tc: bool = True) -> Tensor:
    """[dt | B | C] = u W_x^T per direction (``mamba_simple_faster.py:321-323, 377-379``).
    u (2, B, Lp, D), x_w (2, R+2N, D) -> xdbl (2, B*Lp, R+2N).  bf16 on the GPU: one batched launch of the general tcgen05 GEMM
    (ragged N = 44 / 56 / 80) writing into a buffer whose row pitch is rounded up to 16 bytes (the result is a view of it;
    the scan kernels take the pitch).  Otherwise ``torch.bmm``."""
    _, B, Lp, D = u.shape
    M, ncols = B * Lp, x_w.shape[1]
    if use_tc and u.is_cuda and u.dtype == torch.bfloat16 and x_w.dtype == torch.bfloat16 and D % 8 == 0 \
            and u.is_contiguous() and x_w.is_contiguous():
        ld = (ncols + 7) // 8 * 8
        buf = torch.empty((2, M, ld), device=u.device, dtype=u.dtype)
        _lib.call("fv_gemm_bf16_batched", 2, M, ncols, D, 0, _p(u), D, M * D, 0, _p(x_w), D, ncols * D, FV_BF16, _p(buf), ld,
                  M * ld, 1, _stream(u))     # both directions in one launch
        return buf[..., :ncols]
    return torch.bmm(u.reshape(2, M, D), x_w.transpose(1, 2))


_PATCH_DT = {torch.float32: 0, torch.bfloat16: 1, torch.uint8: 2}


def patchify_supported(img: Tensor, patch: int) -> bool:
    if img.dtype not in _PATCH_DT or img.dim() != 4 or not img.is_cuda or not img.is_contiguous():
        return False
    _, C, H, W = img.shape
    return bool(_lib.lib().fv_patchify_supported(_PATCH_DT[img.dtype], C, H, W, int(patch)))


def patchify(img: Tensor, patch: int, per_channel: bool = False) -> Tensor:
    """img (B, C, H, W) fp32 | bf16 | uint8 contiguous -> bf16 unfolded patches, the A operand of the patch-embedding
    GEMM: (B*gh*gw, C*p*p) or, per channel, (B*C*gh*gw, p*p).  One read of the image, one write."""
    _check_cuda(img)
    B, C, H, W = img.shape
    gh, gw = H // patch, W // patch
    rows, cols = (B * C * gh * gw, patch * patch) if per_channel else (B * gh * gw, C * patch * patch)
    out = torch.empty((rows, cols), device=img.device, dtype=torch.bfloat16)
    _lib.call("fv_patchify", _PATCH_DT[img.dtype], B, C, H, W, int(patch), int(per_channel), _p(img), _p(out), _stream(img))
    return out


# --------------------------------------------------------------------------- (B, D, L) operator-API helpers
def causal_conv1d_fwd(x: Tensor, weight: Tensor, bias: Optional[Tensor], silu: bool = True) -> Tensor:
    """x (B, D, L) with unit stride along L (any batch / channel strides), weight (D, 4) -> (B, D, L) contiguous."""
    _check_cuda(x, weight)
    B, D, L = x.shape
    assert x.stride(2) == 1 and weight.shape == (D, 4)
    out = torch.empty((B, D, L), device=x.device, dtype=x.dtype)
    w32, b32 = _f32c(weight), _f32c(bias)   # converted copies must outlive the launch (_p keeps only the address)
    _lib.call("fv_causal_conv1d_fwd", _dt(x), B, D, L, _p(x), x.stride(0), x.stride(1), _p(w32),
              _p(b32), int(silu), _p(out), _stream(x))
    return out


def pool_bdl_fwd(xc: Tensor, outer: int, pool: int, inner: int = 1, mode: str = "mean", scale: float = 1.0) -> Tensor:
    """(B, D, outer*pool*inner) contiguous -> (B, D, outer*inner): mean * scale, or max, over the pool axis."""
    _check_cuda(xc)
    B, D, L = xc.shape
    assert L == outer * pool * inner and xc.is_contiguous()
    out = torch.empty((B, D, outer * inner), device=xc.device, dtype=xc.dtype)
    _lib.call("fv_pool_bdl_fwd", _dt(xc), B, D, outer, pool, inner, _p(xc), FV_POOL_MAX if mode == "max" else FV_POOL_MEAN,
              float(scale), _p(out), _stream(xc))
    return out


def bcast_skip_bdl_fwd(s: Tensor, xc: Optional[Tensor], Dskip: Optional[Tensor], outer: int, pool: int,
                       inner: int = 1) -> Tensor:
    """out[b, d, t] = s[b, d, pool_index(t)] + D[d] * xc[b, d, t]; s (B, D, outer*inner), xc (B, D, L) contiguous."""
    _check_cuda(s, xc)
    B, D, Lp = s.shape
    assert Lp == outer * inner and s.is_contiguous() and (xc is None or xc.is_contiguous())
    out = torch.empty((B, D, outer * pool * inner), device=s.device, dtype=s.dtype)
    d32 = _f32c(Dskip)
    _lib.call("fv_bcast_skip_bdl_fwd", _dt(s), B, D, outer, pool, inner, _p(s), _p(xc), _p(d32), _p(out),
              _stream(s))
    return out


def selective_scan_bwd(dout: Tensor, u: Tensor, delta: Tensor, A: Tensor, B: Tensor, Cm: Tensor, D: Optional[Tensor],
                       z: Optional[Tensor], delta_bias: Optional[Tensor], delta_softplus: bool):
    """Backward of the operator-API scan.  Returns (du, ddelta, dA, dB, dC, dD, dz, ddelta_bias): du / ddelta / dz in the
    input dtype, dA / dD / ddelta_bias fp32, dB / dC (batch, groups, N, L) in the input dtype after fp32 accumulation."""
    _check_cuda(dout, u, delta, A, B, Cm)
    batch, dim, L = u.shape
    groups, N = B.shape[1], B.shape[2]
    for t in (dout, u, delta, B, Cm, z):
        assert t is None or (t.is_contiguous() and t.dtype == u.dtype)
    f32 = dict(device=u.device, dtype=torch.float32)
    du, ddelta = torch.empty_like(u), torch.empty_like(u)
    dz = torch.empty_like(u) if z is not None else None
    dA = torch.zeros((dim, N), **f32)
    dB, dC = torch.zeros(B.shape, **f32), torch.zeros(B.shape, **f32)
    dD = torch.zeros(dim, **f32) if D is not None else None
    dbias = torch.zeros(dim, **f32) if delta_bias is not None else None
    nbytes = int(_lib.lib().fv_selective_scan_bwd_workspace_bytes(batch, dim, L, N))
    ws = torch.empty(max(nbytes, 4) // 4, **f32)
    _lib.call("fv_selective_scan_bwd", _dt(u), batch, dim, L, N, groups, _p(u), _p(delta), _p(A), _p(B), _p(Cm), _p(D),
              _p(z), _p(delta_bias), int(delta_softplus), _p(dout), _p(du), _p(ddelta), _p(dA), _p(dB), _p(dC), _p(dD),
              _p(dz), _p(dbias), _p(ws), ws.numel() * 4, _stream(u))
    return du, ddelta, dA, dB.to(u.dtype), dC.to(u.dtype), dD, dz, dbias


def causal_conv1d_bwd(x: Tensor, weight: Tensor, bias: Optional[Tensor], dout: Tensor, silu: bool = True):
    """x (B, D, L) unit stride along L, dout (B, D, L) contiguous -> dx (B, D, L) contiguous, dw (D, 4) fp32, db (D) fp32 | None."""
    _check_cuda(x, weight, dout)
    B, D, L = x.shape
    assert x.stride(2) == 1 and weight.shape == (D, 4) and dout.is_contiguous() and dout.dtype == x.dtype
    dx = torch.empty((B, D, L), device=x.device, dtype=x.dtype)
    dw = torch.zeros((D, 4), device=x.device, dtype=torch.float32)
    db = torch.zeros(D, device=x.device, dtype=torch.float32) if bias is not None else None
    w32, b32 = _f32c(weight), _f32c(bias)
    _lib.call("fv_causal_conv1d_bwd", _dt(x), B, D, L, _p(x), x.stride(0), x.stride(1), _p(w32), _p(b32),
              int(silu), _p(dout), _p(dx), dx.stride(0), dx.stride(1), _p(dw), _p(db), _stream(x))
    return dx, dw, db


def rowdot_bdl(a: Tensor, c: Tensor) -> Tensor:
    """(B, D, L) x (B, D, L) contiguous -> (D,) fp32: sum over batch and L of a * c."""
    _check_cuda(a, c)
    B, D, L = a.shape
    assert a.is_contiguous() and c.is_contiguous() and a.dtype == c.dtype and a.shape == c.shape
    out = torch.zeros(D, device=a.device, dtype=torch.float32)
    _lib.call("fv_rowdot_bdl", _dt(a), B, D, L, _p(a), _p(c), _p(out), _stream(a))
    return out


# --------------------------------------------------------------------------- backward wrappers
def bwd_tiles_per_group(geom: Geometry, batch: int, dim: int, dtype: torch.dtype) -> int:
    g = geom.c_struct(batch, dim)
    n = _lib.lib().fv_bwd_tiles_per_group(C.byref(g), FV_F32 if dtype == torch.float32 else FV_BF16)
    if n <= 0:
        raise _lib.FastVimLibraryError("backward kernels need a plain (outer, pool, 1) geometry")
    return int(n)


def gate_bwd(x: Tensor, z: Tensor, dy: Tensor, s: Tensor, geom: Geometry, conv_w: Tensor, conv_b: Optional[Tensor],
             Dskip: Tensor, ln_w: Optional[Tensor], ln_b: Optional[Tensor], eps: float, dz: Tensor):
    """K2b-bwd.  Writes dz in place (the z half of the d(xz) buffer); returns
    (e (B, L, D), ds_planes (tpg, B, Lp, D) fp32, dDskip (2, D), dln_w, dln_b)."""
    _check_cuda(x, z, dy, s)
    B, L, D = x.shape
    ldx, bs = _tokmajor(x, "x")
    assert _tokmajor(z, "z") == (ldx, bs) and _tokmajor(dz, "dz") == (ldx, bs)
    lddy, dybs = _tokmajor(dy, "dy")
    g = geom.c_struct(B, D)
    if GATE_BWD_STREAM and geom.inner == 1 and _lib.lib().fv_gate_bwd_stream_supported(C.byref(g), ldx, lddy):
        f32 = dict(device=x.device, dtype=torch.float32)
        e = torch.empty((B, L, D), device=x.device, dtype=x.dtype)
        ds = torch.zeros((1, B, geom.Lp, D), **f32)
        dD = torch.zeros((2, D), **f32)
        dlw = torch.zeros(D, **f32) if ln_w is not None else None
        dlb = torch.zeros(D, **f32) if ln_w is not None else None
        stats = torch.zeros((B, L, 4), **f32) if ln_w is not None else None
        _lib.call("fv_gate_bwd_stream", C.byref(g), _dt(x), _p(x), _p(z), ldx, bs, _p(dy), lddy, dybs, _p(s), _p(conv_w),
                  _p(conv_b), _p(Dskip), _p(ln_w), _p(ln_b), float(eps), _p(stats), _p(dz), _p(e), _p(ds), _p(dD), _p(dlw),
                  _p(dlb), _stream(x))
        return e, ds, dD, dlw, dlb
    tpg = bwd_tiles_per_group(geom, B, D, x.dtype)
    e = torch.empty((B, L, D), device=x.device, dtype=x.dtype)
    ds = torch.empty((tpg, B, geom.Lp, D), device=x.device, dtype=torch.float32)
    dD = torch.zeros((2, D), device=x.device, dtype=torch.float32)
    dlw = torch.zeros(D, device=x.device, dtype=torch.float32) if ln_w is not None else None
    dlb = torch.zeros(D, device=x.device, dtype=torch.float32) if ln_w is not None else None
    g = geom.c_struct(B, D)
    _lib.call("fv_gate_bwd", C.byref(g), _dt(x), _p(x), _p(z), ldx, bs, _p(dy), lddy, dybs, _p(s), _p(conv_w),
              _p(conv_b), _p(Dskip), _p(ln_w), _p(ln_b), float(eps), _p(dz), _p(e), _p(ds), _p(dD), _p(dlw), _p(dlb),
              _stream(x))
    return e, ds, dD, dlw, dlb


def scan_bwd(ds: Tensor, u: Tensor, xdbl: Tensor, geom: Geometry, dt_rank: int, d_state: int, dt_w: Tensor,
             dt_bias: Tensor, A: Tensor, a_is_log: bool = True, pre: Optional[Tensor] = None):
    """K2a-bwd -> (du, ddelta (2, B, Lp, D) act dtype, dBC (2, B*Lp, 2N) act dtype, dA (2, D, N), d_dt_bias (2, D))."""
    _check_cuda(ds, u, xdbl)
    _, B, Lp, D = u.shape
    assert u.is_contiguous() and ds.is_contiguous() and ds.dtype == torch.float32 and xdbl.stride(2) == 1
    g = geom.c_struct(B, D)
    ncol = int(_lib.lib().fv_scan_bwd_planes(C.byref(g)))
    du = torch.empty_like(u)
    ddelta = torch.empty_like(u)
    planes = torch.empty((ncol, 2, B * Lp, 2 * d_state), device=u.device, dtype=torch.float32)
    dA = torch.zeros((2, D, d_state), device=u.device, dtype=torch.float32)
    dbias = torch.zeros((2, D), device=u.device, dtype=torch.float32)
    g = geom.c_struct(B, D)
    if SCAN_BWD_SHORT and _lib.lib().fv_scan_bwd_short_supported(C.byref(g), d_state):
        # dt_proj as a GEMM (fp32, cuBLAS): delta_pre = dt_bias + dt . W_dt^T, (2, B*Lp, D)
        if pre is None:
            pre = torch.baddbmm(dt_bias.float()[:, None, :], xdbl[..., :dt_rank].float(), dt_w.float().transpose(1, 2))
        _lib.call("fv_scan_bwd_short", C.byref(g), _dt(u), ds.shape[0], _p(u), _p(xdbl), xdbl.stride(1), dt_rank, d_state,
                  _p(pre), _p(A), int(a_is_log), _p(ds), _p(du), _p(ddelta), _p(planes), _p(dA), _p(dbias), _stream(u))
    else:
        _lib.call("fv_scan_bwd", C.byref(g), _dt(u), ds.shape[0], _p(u), _p(xdbl), xdbl.stride(1), dt_rank, d_state,
                  _p(dt_w), _p(dt_bias), _p(A), int(a_is_log), _p(ds), _p(du), _p(ddelta), _p(planes), _p(dA), _p(dbias),
                  _stream(u))
    dbc = torch.empty((2, B * Lp, 2 * d_state), device=u.device, dtype=u.dtype)
    _lib.call("fv_reduce_planes", _dt(u), _p(planes), ncol, dbc.numel(), _p(dbc), _stream(u))
    return du, ddelta, dbc, dA, dbias


def gate_bwd_v_supported(geom: Geometry, batch: int, dim: int, dtype: torch.dtype) -> bool:
    if dtype != torch.bfloat16:
        return False
    g = geom.c_struct(batch, dim)
    # the conv backward must be able to produce dD as well (streaming kernel: dim % 64 == 0)
    return bool(_lib.lib().fv_gate_bwd_v_supported(C.byref(g), FV_BF16)) and dim % 64 == 0


def gate_bwd_v(v: Tensor, z: Tensor, dy: Tensor, geom: Geometry, ln_w: Optional[Tensor], ln_b: Optional[Tensor], eps: float,
               dz: Tensor):
    """K2b-bwd from the saved pre-norm value v (B, L, D).  Writes dz in place; returns (e (B, L, D), ds (1, B, Lp, D) fp32,
    dln_w, dln_b).  The D-skip gradients come from ``conv_pool_bwd(..., want_dD=True)``."""
    _check_cuda(v, z, dy)
    B, L, D = v.shape
    assert v.is_contiguous()
    ldz, zbs = _tokmajor(z, "z")
    assert _tokmajor(dz, "dz") == (ldz, zbs)
    lddy, dybs = _tokmajor(dy, "dy")
    f32 = dict(device=v.device, dtype=torch.float32)
    e = torch.empty((B, L, D), device=v.device, dtype=v.dtype)
    ds = torch.empty((1, B, geom.Lp, D), **f32)
    dlw = torch.zeros(D, **f32) if ln_w is not None else None
    dlb = torch.zeros(D, **f32) if ln_w is not None else None
    g = geom.c_struct(B, D)
    _lib.call("fv_gate_bwd_v", C.byref(g), _dt(v), _p(v), _p(z), ldz, zbs, _p(dy), lddy, dybs, _p(ln_w), _p(ln_b), float(eps),
              _p(dz), _p(e), _p(ds), _p(dlw), _p(dlb), _stream(v))
    return e, ds, dlw, dlb


def conv_pool_bwd(x: Tensor, e: Tensor, du: Tensor, geom: Geometry, conv_w: Tensor, conv_b: Optional[Tensor],
                  Dskip: Tensor, scale: float, dx: Tensor, want_dD: bool = False):
    """K1-bwd.  Writes dx in place (the x half of the d(xz) buffer); returns (dconv_w (2, D, 4), dconv_b (2, D)) and,
    with ``want_dD``, also dDskip (2, D) = sum e * xc."""
    _check_cuda(x, e, du)
    B, L, D = x.shape
    ldx, bs = _tokmajor(x, "x")
    assert _tokmajor(dx, "dx") == (ldx, bs) and e.is_contiguous() and du.is_contiguous()
    dcw = torch.zeros((2, D, 4), device=x.device, dtype=torch.float32)
    dcb = torch.zeros((2, D), device=x.device, dtype=torch.float32) if conv_b is not None else None
    dD = torch.zeros((2, D), device=x.device, dtype=torch.float32) if want_dD else None
    g = geom.c_struct(B, D)
    _lib.call("fv_conv_pool_bwd", C.byref(g), _dt(x), _p(x), ldx, bs, _p(e), _p(du), _p(conv_w), _p(conv_b),
              _p(Dskip), float(scale), FV_POOL_MEAN, _p(dx), _p(dcw), _p(dcb), _p(dD), _stream(x))
    return (dcw, dcb, dD) if want_dD else (dcw, dcb)


def add_norm_bwd(dy: Tensor, dres_out: Optional[Tensor], res_out: Tensor, weight: Tensor, eps: float, is_rms: bool,
                 has_bias: bool, x_dtype: torch.dtype, want_dx: bool, want_dres: bool):
    """-> (dx (x_dtype) | None, dresidual_in fp32 | None, dweight, dbias | None)"""
    _check_cuda(dy, res_out, weight)
    shape = res_out.shape
    cols = shape[-1]
    dy2 = dy.reshape(-1, cols)
    if dy2.stride(1) != 1:
        dy2 = dy2.contiguous()
    rows = dy2.shape[0]
    res2 = res_out.reshape(rows, cols)
    assert res2.is_contiguous() and res2.dtype == torch.float32
    if dres_out is not None:
        dres_out = dres_out.reshape(rows, cols).float().contiguous()
    if dy2.dtype != x_dtype:
        dy2 = dy2.to(x_dtype)
    dx = torch.empty((rows, cols), device=dy.device, dtype=x_dtype) if want_dx else None
    dres = torch.empty((rows, cols), device=dy.device, dtype=torch.float32) if want_dres else None
    dw = torch.zeros(cols, device=dy.device, dtype=torch.float32)
    db = torch.zeros(cols, device=dy.device, dtype=torch.float32) if has_bias else None
    _lib.call("fv_add_norm_bwd", _dt(dy2), rows, cols, _p(dy2), dy2.stride(0), _p(dres_out), _p(res2), _p(weight),
              float(eps), int(is_rms), _p(dx), cols, _p(dres), _p(dw), _p(db), _stream(dy))
    return (None if dx is None else dx.reshape(shape), None if dres is None else dres.reshape(shape), dw, db)
