"""Composed (operator-by-operator) mixer path on the reference's ``(batch, dim, seqlen)`` layout.

The fused token-major kernels (``fv_block_fwd``, K1/K2a/K2b and their ``MixerFn`` backward) cover the plain
``(outer, pool, 1)`` FastVim layouts.  Two reference variants need more general pooling and are run here the way the
reference itself runs them -- one operator at a time -- with every operator on ``libfastvim_b200.so``:

* **FastMaskVim** (``mamba_ssm/modules/mamba_simple_masked_faster.py:167-325``): the sequence is the kept tokens of an
  MAE-masked image; pooling is a scatter-add by ``ids_keep // num_of_col`` with the constant divisor ``num_of_col``
  (``compute_row_means_constantdivide`` :376-416), broadcast is a gather (:261-264).  The b-direction pools and
  re-gathers the flipped sequence with the UN-flipped row ids (:213-215, 297-300) -- reproduced as written
  (SURVEY.md Appendix C.3).
* **Channel-First FastChannelVim training** (``mamba_simple_channel_faster.py:205-420``): ``(rows, cols, tpp)``
  layouts (``inner > 1``), which the ``MixerFn`` backward kernels do not walk.
* **Max-pool training** (``collapse_method="max"``: ``cell_imaging/config/FastChannelVimS_maxpool.yaml``): the fused
  backward kernels implement the mean; the max routes its gradient through ``autograd.PoolMaxBdlFn``.

Kernels: ``fv_causal_conv1d_fwd/_bwd`` (x2 directions), ``fv_pool_bdl_fwd`` / ``fv_bcast_skip_bdl_fwd``,
``fv_selective_scan_fwd/_bwd``; GEMMs are torch matmuls (cuBLAS), LayerNorm / gate / index_add / gather are torch
ops exactly as in the reference module.  Everything is differentiable through ``fastvim_b200.autograd``.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from . import autograd as A


def _direction(u, xc, x_w, dt_w, dt_b, A_log, Dk, R, N, bcast):
    """x_proj -> dt_proj -> pooled selective scan -> broadcast + D skip, for one direction.
    u (B, D, Lp) pooled, xc (B, D, L) conv output; reference mamba_simple_faster.py:312-358."""
    B, D, Lp = u.shape
    act = u.dtype
    x_dbl = F.linear(u.transpose(1, 2).reshape(B * Lp, D), x_w.to(act))                     # (B*Lp, R+2N)
    dt = (dt_w.to(act) @ x_dbl[:, :R].t()).reshape(D, B, Lp).permute(1, 0, 2).contiguous()   # (B, D, Lp)
    Bm = x_dbl[:, R:R + N].reshape(B, Lp, N).transpose(1, 2)
    Cm = x_dbl[:, R + N:].reshape(B, Lp, N).transpose(1, 2)
    s = A.selective_scan_train(u, dt, -torch.exp(A_log.float()), Bm, Cm, None, None, dt_b.float(), True, False)
    return bcast(s, xc, Dk.float())


def mixer_forward_composed(m, hidden_states, act_dtype, *, outer: int, pool: int, inner: int = 1,
                           ids_keep: Optional[torch.Tensor] = None):
    """``m``: a ``fastvim_b200.mixer.Mamba``-like module (reference parameter names).  hidden_states (B, L, d_model)
    in sequence order.  With ``ids_keep`` (B, L) int64 -- ORIGINAL token ids of the kept tokens -- the masked pooling
    of FastMaskVim is used (Lp = num_of_rows, divisor num_of_col); otherwise the (outer, pool, inner) mean pool."""
    if m.collapse_method not in ("mean", "max"):
        raise NotImplementedError(f"fastvim_b200.composed: collapse_method {m.collapse_method!r}")
    if m.collapse_method == "max" and ids_keep is not None:
        raise NotImplementedError("the masked mixer defines collapse_method='mean' only "
                                  "(reference mamba_simple_masked_faster.py:212-216)")
    B, L, _ = hidden_states.shape
    D, R, N = m.d_inner, m.dt_rank, m.d_state
    h = hidden_states.to(act_dtype)
    # (B, 2D, L) with L contiguous, as the reference lays xz out (mamba_simple_faster.py:189-195)
    xz = torch.matmul(m.in_proj.weight.to(act_dtype), h.transpose(1, 2))
    if m.in_proj.bias is not None:
        xz = xz + m.in_proj.bias.to(act_dtype)[:, None]
    x, z = xz[:, :D], xz[:, D:]
    x_flip = x.flip([-1]).contiguous()
    xc = A.CausalConv1dFn.apply(x, m.conv1d.weight, m.conv1d.bias, True)
    xc_b = A.CausalConv1dFn.apply(x_flip, m.conv1d_b.weight, m.conv1d_b.bias, True)

    if ids_keep is None:
        if L != outer * pool * inner:
            raise ValueError(f"sequence length {L} != outer*pool*inner = {outer * pool * inner}")
        scale = float(getattr(m, "scaling_factor", 1))
        if m.collapse_method == "max":     # x.reshape(pre_x_shape).max(dim).values: no scaling factor (:299-305)
            pool_fn = lambda t: A.PoolMaxBdlFn.apply(t, outer, pool, inner)
        else:
            pool_fn = lambda t: A.PoolBdlFn.apply(t, outer, pool, inner, scale)
        bcast = lambda s, t, Dk: A.BcastSkipFn.apply(s, t, Dk, outer, pool, inner)
    else:
        if ids_keep.shape != (B, L):
            raise ValueError(f"ids_keep must be (batch, seqlen) = {(B, L)}, got {tuple(ids_keep.shape)}")
        rows, cols = m.num_of_rows, m.num_of_col
        rid = ids_keep // cols                                                              # :208
        lin = (torch.arange(B, device=rid.device)[:, None] * rows + rid).reshape(-1)        # :385-392

        def pool_fn(t):                                                                     # :394-410 (fp32 accumulators)
            sums = torch.zeros((B * rows, D), device=t.device, dtype=torch.float32)
            sums = sums.index_add(0, lin, t.permute(0, 2, 1).reshape(B * L, D).float())
            return (sums.view(B, rows, D) / cols).permute(0, 2, 1).contiguous().to(t.dtype)

        gidx = rid[:, None, :].expand(-1, D, -1)

        def bcast(s, t, Dk):                                                                # :261-264
            return torch.gather(s, 2, gidx) + (Dk[None, :, None] * t).to(s.dtype)

    out_f = _direction(pool_fn(xc), xc, m.x_proj.weight, m.dt_proj.weight, m.dt_proj.bias, m.A_log, m.D, R, N, bcast)
    out_b = _direction(pool_fn(xc_b), xc_b, m.x_proj_b.weight, m.dt_proj_b.weight, m.dt_proj_b.bias, m.A_b_log, m.D_b,
                       R, N, bcast)
    y = (out_f + out_b.flip([-1])).transpose(1, 2) / 2                                      # (B, L, D)
    if m.use_norm_after_ssm:
        y = F.layer_norm(y.float(), (D,), m.layernorm.weight.float(), m.layernorm.bias.float(), m.layernorm.eps)
    y = (y * F.silu(z.transpose(1, 2))).to(act_dtype)
    out = F.linear(y, m.out_proj.weight.to(act_dtype),
                   None if m.out_proj.bias is None else m.out_proj.bias.to(act_dtype))
    if m.init_layer_scale is not None:
        out = out * m.gamma
    return out
