"""Operator API -- mirror of the reference's ``mamba_ssm/ops/selective_scan_interface.py``.

Same names, argument meaning and error behaviour as the reference's public functions
(``selective_scan_fn`` :105-123, ``mamba_inner_fn_no_out_proj`` :1652-1681,
``mamba_inner_fn_no_out_proj_withoutZ`` :1684-1713,
``FastVim_mamba_inner_fn_no_out_proj_withoutZ`` :1716-1753), operating on the reference's
``(batch, dim, seqlen)`` layout, so the reference's op-level tests read the same against this
module.  Everything runs on ``libfastvim_b200.so``; there is no PyTorch fallback.  All functions are
differentiable (``fastvim_b200.autograd``: SelectiveScanFn and friends over the ``fv_*_bwd`` kernels).
"""
from __future__ import annotations

import torch

from . import autograd as fv_autograd
from . import ops


def _prep_bc(M, batch, dim, L, name):
    """-> (batch, groups, N, L) contiguous.  Accepts the reference's shapes
    (selective_scan_interface.py:39-44, 137-138): (dim, N) constant, (batch, N, L), (batch, G, N, L)."""
    if M.dim() == 2:      # constant over time and batch: one group per channel
        return M[None, :, :, None].expand(batch, dim, M.shape[1], L).contiguous()
    if M.dim() == 3:
        M = M[:, None]
    if M.dim() != 4:
        raise ValueError(f"{name} must have 2, 3 or 4 dims, got {M.dim()}")
    return M.contiguous()


def selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                      return_last_state=False):
    """u, delta, z: (batch, dim, L); A: (dim, N) real; B, C: (dim, N) | (batch, N, L) | (batch, G, N, L);
    D, delta_bias: (dim,) fp32.  Returns out (or (out, last_state (batch, dim, N)))."""
    if A.is_complex():
        raise NotImplementedError("complex A is not used by any FastVim model and is not implemented")
    needs_grad = torch.is_grad_enabled() and any(
        t is not None and t.requires_grad for t in (u, delta, A, B, C, D, z, delta_bias))
    if needs_grad:
        return fv_autograd.selective_scan_train(u, delta, A, B, C, D, z, delta_bias, delta_softplus,
                                                return_last_state)
    batch, dim, L = u.shape
    dt = u.dtype
    u_, delta_ = u.contiguous(), delta.to(dt).contiguous()
    Bm, Cm = _prep_bc(B, batch, dim, L, "B").to(dt), _prep_bc(C, batch, dim, L, "C").to(dt)
    out, last = ops.selective_scan_fwd(
        u_, delta_, A.float().contiguous(), Bm, Cm, None if D is None else D.float().contiguous(),
        None if z is None else z.to(dt).contiguous(),
        None if delta_bias is None else delta_bias.float().contiguous(), bool(delta_softplus),
        want_last_state=return_last_state)
    return (out, last) if return_last_state else out


def selective_scan_fn_compressed(u, u_compressed, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                                 return_last_state=False):
    """The 6-tensor form of the reference's own kernel package (``fastvim_kernel/mamba-1p1p1/faster_mamba_ssm/ops/
    selective_scan_interface.py:129-159``, semantics ``selective_scan_ref`` :162-252): scan over the pooled
    ``u_compressed`` (batch, dim, Lc), output repeated ``L // Lc`` times, D skip on the full-resolution ``u``
    (batch, dim, L), optional ``silu(z)`` gate.  Runs on ``fv_selective_scan_fwd/_bwd`` + ``fv_bcast_skip_bdl_fwd``;
    differentiable, including with ``z`` (the reference's backward raises for ``z``, :79-80, and is fp32-only)."""
    batch, dim, L = u.shape
    Lc = u_compressed.shape[2]
    if L % Lc:
        raise ValueError("Compression factor must be integer")                      # reference :191
    cfac = L // Lc
    res = selective_scan_fn(u_compressed, delta, A, B, C, None, None, delta_bias, delta_softplus, return_last_state)
    s, last = res if return_last_state else (res, None)
    out = fv_autograd.BcastSkipFn.apply(s, u.to(s.dtype).contiguous() if D is not None else None, D, Lc, cfac, 1)
    if z is not None:
        out = (out.float() * torch.nn.functional.silu(z.float())).to(out.dtype)
    return (out, last) if return_last_state else out


# --------------------------------------------------------------------------- fused "inner" functions
# Differentiable: every kernel call below is a torch.autograd.Function over a forward / backward kernel pair
# (autograd.CausalConv1dFn, PoolBdlFn, SelectiveScanFn, BcastSkipFn); the x_proj / dt_proj contractions are torch
# matmuls (cuBLAS, as in the reference's backward, selective_scan_interface.py:698-737), differentiated by autograd.
def _inner(x, z, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A, B, C, D, delta_bias,
           B_proj_bias, C_proj_bias, delta_softplus):
    """conv1d(+SiLU) -> x_proj -> dt_proj -> selective scan (u = conv output, D skip, z gate) on (batch, dim, L)
    tensors: the forward of MambaInnerFnNoOutProj (selective_scan_interface.py:208-330)."""
    batch, dim, L = x.shape
    R, N = delta_proj_weight.shape[1], A.shape[-1]
    act = x.dtype
    xc = fv_autograd.CausalConv1dFn.apply(x, conv1d_weight, conv1d_bias, True)                     # :231-233
    # x_proj / dt_proj are GEMMs (cuBLAS through torch), in the activation dtype as under autocast (:221-226)
    x_dbl = torch.nn.functional.linear(xc.transpose(1, 2).reshape(batch * L, dim), x_proj_weight.to(act))   # :237-239
    delta = (delta_proj_weight.to(act) @ x_dbl[:, :R].t()).reshape(dim, batch, L).permute(1, 0, 2).contiguous()  # :240-243
    if B is None:
        B = x_dbl[:, R:R + N]
        if B_proj_bias is not None:
            B = B + B_proj_bias.to(B.dtype)
        B = B.reshape(batch, L, N).transpose(1, 2)[:, None].contiguous()                        # :248-262
    if C is None:
        C = x_dbl[:, -N:]
        if C_proj_bias is not None:
            C = C + C_proj_bias.to(C.dtype)
        C = C.reshape(batch, L, N).transpose(1, 2)[:, None].contiguous()                        # :263-277
    return selective_scan_fn(xc, delta, A, B, C, D, z=z, delta_bias=delta_bias, delta_softplus=delta_softplus)


def mamba_inner_fn_no_out_proj(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A, B=None, C=None,
                               D=None, delta_bias=None, B_proj_bias=None, C_proj_bias=None, delta_softplus=True):
    """xz: (batch, 2*dim, L) -> (batch, dim, L).  Reference :1652-1681."""
    if A.is_complex():
        raise NotImplementedError("complex A is not used by any FastVim model and is not implemented")
    if xz.stride(-1) != 1:
        xz = xz.contiguous()
    x, z = xz.chunk(2, dim=1)
    return _inner(x, z.contiguous(), conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A, B, C, D,
                  delta_bias, B_proj_bias, C_proj_bias, delta_softplus)


def mamba_inner_fn_no_out_proj_withoutZ(x, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A, B=None,
                                        C=None, D=None, delta_bias=None, B_proj_bias=None, C_proj_bias=None,
                                        delta_softplus=True):
    """x: (batch, dim, L) -> (batch, dim, L), no z gate.  Reference :1684-1713."""
    if A.is_complex():
        raise NotImplementedError("complex A is not used by any FastVim model and is not implemented")
    if x.stride(-1) != 1:
        x = x.contiguous()
    return _inner(x, None, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A, B, C, D, delta_bias,
                  B_proj_bias, C_proj_bias, delta_softplus)


def FastVim_mamba_inner_fn_no_out_proj_withoutZ(x, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A,
                                                B=None, C=None, D=None, delta_bias=None, B_proj_bias=None,
                                                C_proj_bias=None, delta_softplus=True, num_of_col=14,
                                                collapse_method="mean", scaling_factor=1, pre_x_shape=None):
    """x: (batch, dim, L) -> (batch, dim, L): conv -> mean over ``num_of_col`` -> x_proj / dt_proj -> scan over the
    pooled sequence -> repeat_interleave + D * conv.  Reference :1716-1753, forward :452-603.  As in the reference
    (:503-508) only ``collapse_method="mean"`` is defined for this function."""
    if A.is_complex():
        raise NotImplementedError("complex A is not used by any FastVim model and is not implemented")
    if collapse_method != "mean":
        raise NotImplementedError("FastVim_mamba_inner_fn_no_out_proj_withoutZ defines collapse_method='mean' only "
                                  "(reference selective_scan_interface.py:503-508)")
    if B is not None or C is not None:
        raise NotImplementedError("input-independent B / C are not used by FastVim and are not implemented here")
    if x.stride(-1) != 1:
        x = x.contiguous()
    batch, dim, L = x.shape
    if pre_x_shape is not None and int(pre_x_shape[-1]) != int(num_of_col):
        raise ValueError(f"pre_x_shape {tuple(pre_x_shape)} does not end with num_of_col={num_of_col}")
    if L % num_of_col:
        raise ValueError(f"seqlen {L} is not a multiple of num_of_col {num_of_col}")
    rows = L // num_of_col
    R, N = delta_proj_weight.shape[1], A.shape[-1]
    act = x.dtype
    xc = fv_autograd.CausalConv1dFn.apply(x, conv1d_weight, conv1d_bias, True)                     # :496-498
    u = fv_autograd.PoolBdlFn.apply(xc, rows, num_of_col, 1, float(scaling_factor))                # :503-508
    x_dbl = torch.nn.functional.linear(u.transpose(1, 2).reshape(batch * rows, dim), x_proj_weight.to(act))  # :512-514
    delta = (delta_proj_weight.to(act) @ x_dbl[:, :R].t()).reshape(dim, batch, rows).permute(1, 0, 2).contiguous()
    Bm = x_dbl[:, R:R + N]
    Cm = x_dbl[:, -N:]
    if B_proj_bias is not None:
        Bm = Bm + B_proj_bias.to(Bm.dtype)
    if C_proj_bias is not None:
        Cm = Cm + C_proj_bias.to(Cm.dtype)
    Bm = Bm.reshape(batch, rows, N).transpose(1, 2)[:, None].contiguous()
    Cm = Cm.reshape(batch, rows, N).transpose(1, 2)[:, None].contiguous()
    s = selective_scan_fn(u, delta, A, Bm, Cm, None, None, delta_bias, delta_softplus)          # :556-566
    return fv_autograd.BcastSkipFn.apply(s, xc, D, rows, num_of_col, 1)                            # :570-571
