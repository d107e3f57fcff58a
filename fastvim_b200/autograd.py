"""Training path: ``torch.autograd.Function``s over the CUDA forward / backward kernels.

``MixerFn`` is the whole SSM mixer between (and including) ``in_proj`` and ``out_proj``; it plays the
role of the reference's fused autograd functions (``FastVim_MambaInnerFnNoOutProj_withoutZ``,
``mamba_ssm/ops/selective_scan_interface.py:452-776``) for the live module branch
(``mamba_simple_faster.py:269-453``), in token-major layout.  GEMMs are cuBLAS calls through torch
(as in the reference, ``:698-737``); everything between them runs on ``libfastvim_b200.so``.
``AddNormFn`` is the fused residual-add + norm (``ops/triton/layernorm.py:402-489`` ``LayerNormFn``).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops


class MixerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, in_w, in_b, conv_w, conv_b, x_w, dt_w, dt_b, A_log, Dk, ln_w, ln_b, out_w, out_b,
                geom, scale, eps, d_state, dt_rank):
        """h (B, L, dm) act dtype; in_w (2D, dm), out_w (dm, D), x_w (2, R+2N, D) act dtype;
        conv_w (2, D, 4), conv_b (2, D) | None, dt_w (2, D, R), dt_b (2, D), A_log (2, D, N), Dk (2, D),
        ln_w / ln_b (D) | None: fp32."""
        B, L, _ = h.shape
        D = conv_w.shape[1]
        from . import mixer as _mixer

        xz = _mixer.linear(h, in_w, in_b)
        x, z = xz[..., :D], xz[..., D:]
        if _mixer.FUSED_BLOCK and ops.block_fwd_supported(geom, B, D, xz.dtype, dt_rank, d_state):
            y, u, xdbl, s = ops.block_fwd(x, z, geom, conv_w, conv_b, x_w.contiguous(), dt_w.contiguous(),
                                          dt_b, A_log, Dk, ln_w, ln_b, eps, scale, dt_rank, d_state, a_is_log=True,
                                          save=True)
        else:
            u = ops.conv_pool_fwd(x, geom, conv_w, conv_b, scale, "mean")
            xdbl = torch.bmm(u.view(2, B * geom.Lp, D), x_w.transpose(1, 2))
            s = ops.scan_fwd(u, xdbl, geom, dt_rank, d_state, dt_w, dt_b, A_log, a_is_log=True)
            y = ops.gate_fwd(x, z, s, geom, conv_w, conv_b, Dk, ln_w, ln_b, eps)
        out = _mixer.linear(y, out_w, out_b)
        ctx.save_for_backward(h, in_w, conv_w, conv_b, x_w, dt_w, dt_b, A_log, Dk, ln_w, ln_b, out_w, xz, u, xdbl, s, y)
        ctx.meta = (geom, scale, eps, d_state, dt_rank, in_b is not None, out_b is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        (h, in_w, conv_w, conv_b, x_w, dt_w, dt_b, A_log, Dk, ln_w, ln_b, out_w, xz, u, xdbl, s, y) = ctx.saved_tensors
        geom, scale, eps, N, R, has_in_b, has_out_b = ctx.meta
        B, L, dm = h.shape
        D = conv_w.shape[1]
        Lp = geom.Lp
        dt = xz.dtype
        dout = dout.to(dt).contiguous()
        dout2 = dout.view(B * L, dm)
        # out_proj
        dy = (dout2 @ out_w).view(B, L, D)
        d_out_w = dout2.t() @ y.view(B * L, D)
        d_out_b = dout2.sum(0) if has_out_b else None
        # epilogue
        x, z = xz[..., :D], xz[..., D:]
        dxz = torch.empty_like(xz)
        e, ds, dDk, dln_w, dln_b = ops.gate_bwd(x, z, dy, s, geom, conv_w, conv_b, Dk, ln_w, ln_b, eps, dxz[..., D:])
        # scan
        du, ddelta, dbc, dA_log, d_dt_b = ops.scan_bwd(ds, u, xdbl, geom, R, N, dt_w, dt_b, A_log, True)
        ddelta2 = ddelta.view(2, B * Lp, D)
        ddt = torch.bmm(ddelta2, dt_w.to(dt))                                   # (2, B*Lp, R)
        d_dt_w = torch.bmm(ddelta2.transpose(1, 2), xdbl[..., :R]).float()      # (2, D, R)
        dxdbl = torch.cat([ddt, dbc], dim=-1)                                   # (2, B*Lp, R+2N)
        u2 = u.view(2, B * Lp, D)
        d_x_w = torch.bmm(dxdbl.transpose(1, 2), u2)                            # (2, R+2N, D)
        du_total = torch.baddbmm(du.view(2, B * Lp, D), dxdbl, x_w).view(2, B, Lp, D).contiguous()
        # conv + pool (+ D skip)
        d_conv_w, d_conv_b = ops.conv_pool_bwd(x, e, du_total, geom, conv_w, conv_b, Dk, scale, dxz[..., :D])
        # in_proj
        dxz2 = dxz.view(B * L, 2 * D)
        dh = (dxz2 @ in_w).view(B, L, dm)
        d_in_w = dxz2.t() @ h.view(B * L, dm)
        d_in_b = dxz2.sum(0) if has_in_b else None
        return (dh, d_in_w, d_in_b, d_conv_w, d_conv_b, d_x_w, d_dt_w, d_dt_b, dA_log, dDk, dln_w, dln_b, d_out_w,
                d_out_b, None, None, None, None, None)


def mixer_forward_train(mixer, hidden_states, geom, act_dtype):
    """Differentiable forward of ``fastvim_b200.mixer.Mamba``: parameters are stacked per direction and
    cast with ordinary (differentiable) torch ops, then handed to ``MixerFn``."""
    if mixer.collapse_method != "mean":
        raise NotImplementedError("fastvim_b200: training is implemented for collapse_method='mean' "
                                  "(the reference's fused autograd path ignores 'max' too, "
                                  "selective_scan_interface.py:503-508)")
    f32 = torch.float32
    m = mixer
    conv_w = torch.stack([m.conv1d.weight[:, 0], m.conv1d_b.weight[:, 0]]).to(f32)
    conv_b = None if m.conv1d.bias is None else torch.stack([m.conv1d.bias, m.conv1d_b.bias]).to(f32)
    x_w = torch.stack([m.x_proj.weight, m.x_proj_b.weight]).to(act_dtype)
    dt_w = torch.stack([m.dt_proj.weight, m.dt_proj_b.weight]).to(f32)
    dt_b = torch.stack([m.dt_proj.bias, m.dt_proj_b.bias]).to(f32)
    A_log = torch.stack([m.A_log, m.A_b_log]).to(f32)
    Dk = torch.stack([m.D, m.D_b]).to(f32)
    ln_w = m.layernorm.weight.to(f32) if m.use_norm_after_ssm else None
    ln_b = m.layernorm.bias.to(f32) if m.use_norm_after_ssm else None
    in_b = None if m.in_proj.bias is None else m.in_proj.bias.to(act_dtype)
    out_b = None if m.out_proj.bias is None else m.out_proj.bias.to(act_dtype)
    return MixerFn.apply(hidden_states.to(act_dtype), m.in_proj.weight.to(act_dtype), in_b, conv_w, conv_b, x_w, dt_w,
                         dt_b, A_log, Dk, ln_w, ln_b, m.out_proj.weight.to(act_dtype), out_b, geom,
                         float(m.scaling_factor), m.layernorm.eps if m.use_norm_after_ssm else 1e-5, m.d_state,
                         m.dt_rank)


class AddNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, residual, eps, prenorm, is_rms):
        w = weight.float()
        b = None if bias is None else bias.float()
        res = None if residual is None else residual.float()
        y, res_out, _, _ = ops.add_norm_fwd(x, res, w, b, eps, is_rms, want_residual=True)
        ctx.save_for_backward(res_out, w)
        ctx.meta = (eps, is_rms, bias is not None, x.dtype, residual is not None, prenorm,
                    None if residual is None else residual.dtype, weight.dtype)
        if prenorm:
            return y, res_out
        return y

    @staticmethod
    def backward(ctx, dy, *rest):
        res_out, w = ctx.saved_tensors
        eps, is_rms, has_bias, x_dtype, has_res, prenorm, res_dtype, w_dtype = ctx.meta
        dres_out = rest[0] if (prenorm and rest and rest[0] is not None) else None
        dx, dres, dw, db = ops.add_norm_bwd(dy.contiguous(), dres_out, res_out, w, eps, is_rms, has_bias, x_dtype,
                                            want_dx=True, want_dres=has_res)
        if dres is not None and res_dtype != torch.float32:
            dres = dres.to(res_dtype)
        return dx, dw.to(w_dtype), (None if db is None else db.to(w_dtype)), dres, None, None, None


def add_norm_train(x, weight, bias, residual, eps, prenorm, residual_in_fp32, is_rms):
    out = AddNormFn.apply(x, weight, bias, residual, eps, prenorm, is_rms)
    if not prenorm:
        return out
    y, res_out = out
    if not residual_in_fp32 and residual is None:
        res_out = res_out.to(x.dtype)
    return y, res_out


def selective_scan_train(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state):
    raise NotImplementedError(
        "fastvim_b200: the backward of the (batch, dim, L) operator selective_scan_fn is not built yet; the "
        "model training path goes through fastvim_b200.mixer.Mamba (MixerFn)")
