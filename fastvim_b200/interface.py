"""Operator API -- mirror of the reference's ``mamba_ssm/ops/selective_scan_interface.py``.

Same names, argument meaning and error behaviour as the reference's public functions
(``selective_scan_fn`` :105-123, ``mamba_inner_fn_no_out_proj`` :1652-1681,
``mamba_inner_fn_no_out_proj_withoutZ`` :1684-1713,
``FastVim_mamba_inner_fn_no_out_proj_withoutZ`` :1716-1753), operating on the reference's
``(batch, dim, seqlen)`` layout, so the reference's op-level tests read the same against this
module.  Everything runs on ``libfastvim_b200.so``; there is no PyTorch fallback.
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
