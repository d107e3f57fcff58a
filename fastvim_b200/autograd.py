"""Training path: ``torch.autograd.Function``s over the CUDA forward / backward kernels.

``MixerFn`` is the whole SSM mixer between (and including) ``in_proj`` and ``out_proj``; it plays the
role of the reference's fused autograd functions (``FastVim_MambaInnerFnNoOutProj_withoutZ``,
``mamba_ssm/ops/selective_scan_interface.py:452-776``) for the live module branch
(``mamba_simple_faster.py:269-453``), in token-major layout.  GEMMs fall back to cuBLAS through torch
(as in the reference, ``:698-737``) only where the tcgen05 GEMM does not apply (fp32 runs, 8-byte row pitches): the dgrad /
wgrad GEMMs of in_proj / out_proj run on ``fv_gemm_bf16`` (csrc/gemm_tc2.cu), everything between them on the other kernels of
``libfastvim_b200.so``.
``AddNormFn`` is the fused residual-add + norm (``ops/triton/layernorm.py:402-489`` ``LayerNormFn``).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

import os

from . import _lib, ops

# streaming gate backward from the saved pre-norm value (csrc/gate_bwd_v.cu); "0" = round-1 recomputing kernel
GATE_BWD_V = os.environ.get("FASTVIM_GATE_BWD_V", "1") != "0"
# x_proj / dt_proj backward GEMMs (N or K = dt_rank, dt_rank + 2 d_state) on the general tcgen05 GEMM where every row pitch is a
# multiple of 16 bytes (FastVim-S/B: dt_rank 24 / 48): four batched launches (both directions each), the two wgrad ones
# accumulating their K splits with TMA reduction stores.  Parity-tested; measured 0.43 ms per FastVim-B step SLOWER than the four
# cuBLAS bmm calls (32.61 vs 32.17 ms, three alternating pairs on one box; eight single launches + plane reductions: +1.25 ms):
# one 128-row tile wide and a few k-blocks deep, these products are latency-bound in a persistent tensor-core kernel.  Opt-in ("1").
TC_SMALL_GEMM = os.environ.get("FASTVIM_TC_SMALL_GEMM", "0") == "1"
# programmatic dependent launch for every kernel of the training forward / backward: measured between -0.6 % and +3.6 % on
# the FastVim-B / -T step, i.e. inside the run-to-run spread of the training line (+-2 %); opt-in until it is measured at N > 1
TRAIN_PDL = os.environ.get("FASTVIM_TRAIN_PDL", "0") == "1"


def _mm_f32(a, b):
    """a @ b for bf16 / fp32 operands with an fp32 RESULT (tensor-core GEMM, fp32 accumulator written out unrounded):
    weight gradients go straight into fp32 master-gradient buffers, no bf16 rounding and no cast pass."""
    if a.dtype == torch.float32:
        return a @ b
    try:
        return torch.mm(a, b, out_dtype=torch.float32)
    except (TypeError, RuntimeError):   # older torch / CPU stand-in runs
        return (a @ b).float()


def _bmm_f32(a, b):
    if a.dtype == torch.float32:
        return torch.bmm(a, b)
    try:
        return torch.bmm(a, b, out_dtype=torch.float32)
    except (TypeError, RuntimeError):
        return torch.bmm(a, b).float()


def _tc():
    from . import mixer as _mixer
    return _mixer.TC_GEMM


def _train_pdl(fn):
    """With ``FASTVIM_TRAIN_PDL=1`` the training forward / backward bodies run with programmatic dependent launch on for
    every kernel (``_lib.pdl_all``): the ~145 launches of a step overlap their prologues with the previous kernel's tail.
    CUDA tensors only (CPU stand-in tests never load the library)."""
    import functools

    @functools.wraps(fn)
    def wrapped(ctx, *args):
        t = next((a for a in args if isinstance(a, torch.Tensor)), None)
        if TRAIN_PDL and t is not None and t.is_cuda:
            with _lib.pdl_all():
                return fn(ctx, *args)
        return fn(ctx, *args)
    return wrapped


def _dgrad(dy2, w):
    """dX (M, K') = dY (M, N') @ W (N', K'): tcgen05 GEMM with W read as an MN-major operand (no transpose copy)."""
    if _tc() and ops.gemm_bf16_ok(dy2, w):
        return ops.gemm_bf16(dy2, w, b_mn=True)
    return dy2 @ w


def _wgrad(dy2, x2):
    """dW (N', K') fp32 = dY (M, N').T @ X (M, K'): both operands MN-major, split-K over the tokens, fp32 planes."""
    if _tc() and ops.gemm_bf16_ok(dy2, x2):
        return ops.gemm_bf16(dy2, x2, a_mn=True, b_mn=True, out_f32=True)
    return _mm_f32(dy2.t(), x2)


class MixerFn(torch.autograd.Function):
    @staticmethod
    @_train_pdl
    def forward(ctx, h, in_w, in_b, conv_w, conv_b, x_w, dt_w, dt_b, A_log, Dk, ln_w, ln_b, out_w, out_b,
                geom, scale, eps, d_state, dt_rank):
        """h (B, L, dm) act dtype; in_w (2D, dm), out_w (dm, D), x_w (2, R+2N, D): MASTER weights (fp32 under autocast) --
        they are cast to the activation dtype here, outside autograd, and their gradients are returned in their own dtype;
        conv_w (2, D, 4), conv_b (2, D) | None, dt_w (2, D, R), dt_b (2, D), A_log (2, D, N), Dk (2, D),
        ln_w / ln_b (D) | None: fp32."""
        B, L, _ = h.shape
        D = conv_w.shape[1]
        from . import mixer as _mixer

        act = h.dtype
        w_dtypes = (in_w.dtype, out_w.dtype, x_w.dtype, None if in_b is None else in_b.dtype,
                    None if out_b is None else out_b.dtype)
        in_w, out_w, x_w = in_w.to(act), out_w.to(act), x_w.to(act).contiguous()
        in_b = None if in_b is None else in_b.to(act)
        out_b = None if out_b is None else out_b.to(act)
        h = h.contiguous()   # saved for the backward's (B*L, dm) views
        xz = _mixer.linear(h, in_w, in_b)
        x, z = xz[..., :D], xz[..., D:]
        v = pre = None
        if _mixer.FUSED_BLOCK and ops.block_fwd_supported(geom, B, D, xz.dtype, dt_rank, d_state):
            # the cluster kernel also saves the pre-norm value v (instead of the scan planes s) for the streaming gate backward
            want_v = GATE_BWD_V and ops.gate_bwd_v_supported(geom, B, D, xz.dtype)
            y, u, xdbl, s, vp = ops.block_fwd(x, z, geom, conv_w, conv_b, x_w, dt_w.contiguous(),
                                              dt_b, -torch.exp(A_log), Dk, ln_w, ln_b, eps, scale, dt_rank, d_state,
                                              a_is_log=False, save=True, save_v=want_v)
            if vp is not None:
                v, pre = vp
        else:
            u = ops.conv_pool_fwd(x, geom, conv_w, conv_b, scale, "mean")
            xdbl = ops.x_proj(u, x_w, _mixer.TC_GEMM)
            s = ops.scan_fwd(u, xdbl, geom, dt_rank, d_state, dt_w, dt_b, A_log, a_is_log=True)
            y = ops.gate_fwd(x, z, s, geom, conv_w, conv_b, Dk, ln_w, ln_b, eps)
        out = _mixer.linear(y, out_w, out_b)
        ctx.save_for_backward(h, in_w, conv_w, conv_b, x_w, dt_w, dt_b, A_log, Dk, ln_w, ln_b, out_w, xz, u, xdbl, s, y, v, pre)
        ctx.meta = (geom, scale, eps, d_state, dt_rank, in_b is not None, out_b is not None, w_dtypes)
        return out

    @staticmethod
    @_train_pdl
    def backward(ctx, dout):
        (h, in_w, conv_w, conv_b, x_w, dt_w, dt_b, A_log, Dk, ln_w, ln_b, out_w, xz, u, xdbl, s, y, v, pre) = ctx.saved_tensors
        geom, scale, eps, N, R, has_in_b, has_out_b, w_dtypes = ctx.meta
        in_w_dt, out_w_dt, x_w_dt, in_b_dt, out_b_dt = w_dtypes
        B, L, dm = h.shape
        D = conv_w.shape[1]
        Lp = geom.Lp
        dt = xz.dtype
        dout = dout.to(dt).contiguous()
        dout2 = dout.view(B * L, dm)
        # out_proj
        dy = _dgrad(dout2, out_w).view(B, L, D)
        d_out_w = _wgrad(dout2, y.view(B * L, D)).to(out_w_dt)
        d_out_b = dout2.sum(0).to(out_b_dt) if has_out_b else None
        # epilogue
        x, z = xz[..., :D], xz[..., D:]
        dxz = torch.empty_like(xz)
        if v is not None:   # streaming: v, z, dy in -> dz, e out; the D-skip gradients come from the conv backward below
            e, ds, dln_w, dln_b = ops.gate_bwd_v(v, z, dy, geom, ln_w, ln_b, eps, dxz[..., D:])
            dDk = None
        else:
            e, ds, dDk, dln_w, dln_b = ops.gate_bwd(x, z, dy, s, geom, conv_w, conv_b, Dk, ln_w, ln_b, eps, dxz[..., D:])
        # scan
        du, ddelta, dbc, dA_log, d_dt_b = ops.scan_bwd(ds, u, xdbl, geom, R, N, dt_w, dt_b, A_log, True, pre=pre)
        ddelta2 = ddelta.view(2, B * Lp, D)
        u2 = u.view(2, B * Lp, D)
        ncols = R + 2 * N
        dt_w_a = dt_w.to(dt)
        if (TC_SMALL_GEMM and _tc() and R % 8 == 0 and xdbl.stride(0) % 8 == 0
                and ops.gemm_bf16_ok(ddelta2[0], u2[0], xdbl[0], x_w[0], dt_w_a[0])):
            # x_proj / dt_proj backward on the general tcgen05 GEMM (FastVim-S/B: every row pitch is a multiple of 16 bytes)
            # four batched launches (both directions each); the wgrad ones accumulate their K splits with TMA reduction stores
            dxdbl = torch.empty((2, B * Lp, ncols), device=u.device, dtype=dt)
            dxdbl[..., R:] = dbc
            ops.gemm_bf16_batched(ddelta2, dt_w_a, b_mn=True, out=dxdbl[..., :R])                               # ddt
            d_dt_w = ops.gemm_bf16_batched(ddelta2, xdbl[..., :R], a_mn=True, b_mn=True, out_f32=True)          # (2, D, R)
            d_x_w = ops.gemm_bf16_batched(dxdbl, u2, a_mn=True, b_mn=True, out_f32=True).to(x_w_dt)             # (2, R+2N, D)
            du_total = ops.gemm_bf16_batched(dxdbl, x_w, b_mn=True)                                             # (2, B*Lp, D)
            du_total = du_total.add_(du.view(2, B * Lp, D)).view(2, B, Lp, D)
        else:
            ddt = torch.bmm(ddelta2, dt_w_a)                                        # (2, B*Lp, R)
            d_dt_w = _bmm_f32(ddelta2.transpose(1, 2), xdbl[..., :R])               # (2, D, R) fp32
            dxdbl = torch.cat([ddt, dbc], dim=-1)                                   # (2, B*Lp, R+2N)
            d_x_w = _bmm_f32(dxdbl.transpose(1, 2), u2).to(x_w_dt)                  # (2, R+2N, D)
            du_total = torch.baddbmm(du.view(2, B * Lp, D), dxdbl, x_w).view(2, B, Lp, D).contiguous()
        # conv + pool (+ D skip)
        if dDk is None:
            d_conv_w, d_conv_b, dDk = ops.conv_pool_bwd(x, e, du_total, geom, conv_w, conv_b, Dk, scale, dxz[..., :D],
                                                        want_dD=True)
        else:
            d_conv_w, d_conv_b = ops.conv_pool_bwd(x, e, du_total, geom, conv_w, conv_b, Dk, scale, dxz[..., :D])
        # in_proj
        dxz2 = dxz.view(B * L, 2 * D)
        dh = _dgrad(dxz2, in_w).view(B, L, dm)
        d_in_w = _wgrad(dxz2, h.view(B * L, dm)).to(in_w_dt)
        d_in_b = dxz2.sum(0).to(in_b_dt) if has_in_b else None
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
    x_w = torch.stack([m.x_proj.weight, m.x_proj_b.weight])     # master dtype: cast inside MixerFn
    dt_w = torch.stack([m.dt_proj.weight, m.dt_proj_b.weight]).to(f32)
    dt_b = torch.stack([m.dt_proj.bias, m.dt_proj_b.bias]).to(f32)
    A_log = torch.stack([m.A_log, m.A_b_log]).to(f32)
    Dk = torch.stack([m.D, m.D_b]).to(f32)
    ln_w = m.layernorm.weight.to(f32) if m.use_norm_after_ssm else None
    ln_b = m.layernorm.bias.to(f32) if m.use_norm_after_ssm else None
    # in_proj / out_proj / x_proj master weights go in as they are: MixerFn casts them to the activation dtype outside
    # autograd and hands back gradients in the master dtype (fp32 GEMM results, no bf16 rounding, no cast pass)
    return MixerFn.apply(hidden_states.to(act_dtype), m.in_proj.weight, m.in_proj.bias, conv_w, conv_b, x_w, dt_w,
                         dt_b, A_log, Dk, ln_w, ln_b, m.out_proj.weight, m.out_proj.bias, geom,
                         float(m.scaling_factor), m.layernorm.eps if m.use_norm_after_ssm else 1e-5, m.d_state,
                         m.dt_rank)


# patch embedding under autograd on the library's kernels; "0" = eager unfold + cuBLAS F.linear
NATIVE_PATCH_TRAIN = os.environ.get("FASTVIM_NATIVE_PATCH_TRAIN", "1") != "0"


class PatchEmbedFn(torch.autograd.Function):
    """Training form of the patch embedding (reference models/fastvim.py:67-103, an ``nn.Conv2d`` with kernel = stride =
    patch under autocast): ``fv_patchify`` (one pass image -> bf16 patches) + the tcgen05 GEMM with the bias added in its
    epilogue.  Backward: ``dW = dY.T @ patches`` on the general tcgen05 GEMM (both operands MN-major, split-K over the
    tokens, fp32), ``db`` = column sums of ``dY``.  The images get no gradient (callers with ``x.requires_grad`` stay on
    the eager path).  ``w`` (E, C, p, p) / ``b`` (E) are the MASTER parameters: they are cast here, outside autograd, on
    every call (a CUDA-graph-replayed optimizer step never bumps ``_version``, so nothing is cached)."""

    @staticmethod
    def forward(ctx, x, w, b, patch, per_channel=False):
        # (B*gh*gw, C*p*p) bf16; per channel (FastChannelVim's shared projection): (B*C*gh*gw, p*p)
        cols = ops.patchify(x, patch, per_channel=per_channel)
        wmat = w.reshape(w.shape[0], -1).to(torch.bfloat16).contiguous()
        # the reference's autocast conv adds the bias rounded to bf16; keep that rounding, in an fp32 container
        b32 = None if b is None else b.to(torch.bfloat16).float().contiguous()
        out = ops.gemm_bf16_tn(cols, wmat, bias=b32)
        ctx.save_for_backward(cols)
        ctx.meta = (tuple(w.shape), w.dtype, None if b is None else b.dtype)
        return out

    @staticmethod
    def backward(ctx, dy):
        (cols,) = ctx.saved_tensors
        wshape, wdt, bdt = ctx.meta
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        dw = db = None
        if ctx.needs_input_grad[1]:
            dw = _wgrad(dy2, cols).reshape(wshape).to(wdt)
        if bdt is not None and ctx.needs_input_grad[2]:
            db = dy2.sum(dim=0, dtype=torch.float32).to(bdt)
        return None, dw, db, None, None


class AddNormFn(torch.autograd.Function):
    @staticmethod
    @_train_pdl
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
    @_train_pdl
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
    if not residual_in_fp32:
        res_out = res_out.to(x.dtype if residual is None else residual.dtype)
    return y, res_out


# --------------------------------------------------------------------------- operator API on (batch, dim, L)
class SelectiveScanFn(torch.autograd.Function):
    """``SelectiveScanFn`` of the reference (``selective_scan_interface.py:12-102``) over fv_selective_scan_fwd/_bwd.
    B, C arrive as (batch, groups, N, L); nothing but the inputs is saved (the backward re-runs the recurrence)."""

    @staticmethod
    def forward(ctx, u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state):
        out, last = ops.selective_scan_fwd(u, delta, A, B, C, D, z, delta_bias, delta_softplus,
                                           want_last_state=return_last_state)
        ctx.save_for_backward(u, delta, A, B, C, D, z, delta_bias)
        ctx.delta_softplus = delta_softplus
        if return_last_state:
            ctx.mark_non_differentiable(last)   # "the gradient of the last state is not considered" (:115-117)
            return out, last
        return out

    @staticmethod
    def backward(ctx, dout, *unused):
        u, delta, A, B, C, D, z, delta_bias = ctx.saved_tensors
        du, ddelta, dA, dB, dC, dD, dz, dbias = ops.selective_scan_bwd(
            dout.to(u.dtype).contiguous(), u, delta, A, B, C, D, z, delta_bias, ctx.delta_softplus)
        return du, ddelta, dA, dB, dC, dD, dz, dbias, None, None


def _bc_4d(M, batch, dim, L, dt):
    """B / C in any of the reference's shapes -> (batch, groups, N, L) contiguous, differentiably."""
    if M.dim() == 2:      # (dim, N): constant over batch and time, one group per channel
        return M.to(dt)[None, :, :, None].expand(batch, dim, M.shape[1], L).contiguous()
    if M.dim() == 3:
        M = M[:, None]
    if M.dim() != 4:
        raise ValueError(f"B / C must have 2, 3 or 4 dims, got {M.dim()}")
    return M.to(dt).contiguous()


def selective_scan_train(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state):
    """Differentiable ``selective_scan_fn``: dtype / layout preparation with ordinary torch ops, then SelectiveScanFn."""
    batch, dim, L = u.shape
    dt = u.dtype
    f32 = torch.float32
    return SelectiveScanFn.apply(
        u.contiguous(), delta.to(dt).contiguous(), A.to(f32).contiguous(), _bc_4d(B, batch, dim, L, dt),
        _bc_4d(C, batch, dim, L, dt), None if D is None else D.to(f32).contiguous(),
        None if z is None else z.to(dt).contiguous(), None if delta_bias is None else delta_bias.to(f32).contiguous(),
        bool(delta_softplus), bool(return_last_state))


class CausalConv1dFn(torch.autograd.Function):
    """Depthwise causal conv + SiLU on (batch, dim, L) (``causal_conv1d_fn``; call sites
    ``selective_scan_interface.py:231-233, 496-498`` forward and ``:751-753`` backward)."""

    @staticmethod
    def forward(ctx, x, weight, bias, silu):
        w = weight.reshape(weight.shape[0], -1)
        ctx.save_for_backward(x, w, bias)
        ctx.meta = (silu, weight.shape, weight.dtype, None if bias is None else bias.dtype)
        return ops.causal_conv1d_fwd(x, w, bias, silu)

    @staticmethod
    def backward(ctx, dout):
        x, w, bias = ctx.saved_tensors
        silu, wshape, wdt, bdt = ctx.meta
        dx, dw, db = ops.causal_conv1d_bwd(x, w, bias, dout.to(x.dtype).contiguous(), silu)
        return dx, dw.reshape(wshape).to(wdt), (None if db is None else db.to(bdt)), None


class PoolBdlFn(torch.autograd.Function):
    """Mean pool (times scaling_factor) over the ``pool`` axis of a (batch, dim, outer*pool*inner) sequence
    (``selective_scan_interface.py:503-508``); backward = broadcast of the pooled gradient (``:700-704``)."""

    @staticmethod
    def forward(ctx, xc, outer, pool, inner, scale):
        ctx.meta = (outer, pool, inner, scale)
        return ops.pool_bdl_fwd(xc, outer, pool, inner, "mean", scale)

    @staticmethod
    def backward(ctx, du):
        outer, pool, inner, scale = ctx.meta
        return ops.bcast_skip_bdl_fwd((du * (scale / pool)).contiguous(), None, None, outer, pool, inner), None, None, None, None


class PoolMaxBdlFn(torch.autograd.Function):
    """Max pool over the ``pool`` axis of a (batch, dim, outer*pool*inner) sequence (the live module branch,
    ``mamba_simple_faster.py:299-305``, ``mamba_simple_channel_faster.py:258-289``: ``x.reshape(...).max(dim).values``).
    Forward on ``fv_pool_bdl_fwd``; backward routes the pooled gradient to the FIRST position that attains the maximum
    (what ``torch.max(dim)`` differentiates to), as an elementwise mask."""

    @staticmethod
    def forward(ctx, xc, outer, pool, inner):
        u = ops.pool_bdl_fwd(xc, outer, pool, inner, "max", 1.0)
        ctx.save_for_backward(xc, u)
        ctx.meta = (outer, pool, inner)
        return u

    @staticmethod
    def backward(ctx, du):
        xc, u = ctx.saved_tensors
        outer, pool, inner = ctx.meta
        B, D, L = xc.shape
        hit = xc.view(B, D, outer, pool, inner) == u.view(B, D, outer, 1, inner)
        first = hit & (hit.cumsum(dim=3) == 1)
        dx = first.to(du.dtype) * du.reshape(B, D, outer, 1, inner)
        return dx.reshape(B, D, L).to(xc.dtype), None, None, None


class BcastSkipFn(torch.autograd.Function):
    """out = repeat_interleave(s) + D * xc (``selective_scan_interface.py:570-571``); backward: ds = sum over the pool
    axis of dout, dxc = D * dout, dD = sum dout * xc (``:636-642``)."""

    @staticmethod
    def forward(ctx, s, xc, Dskip, outer, pool, inner):
        ctx.save_for_backward(xc, Dskip)
        ctx.meta = (outer, pool, inner)
        return ops.bcast_skip_bdl_fwd(s, xc, Dskip, outer, pool, inner)

    @staticmethod
    def backward(ctx, dout):
        xc, Dskip = ctx.saved_tensors
        outer, pool, inner = ctx.meta
        dout = dout.contiguous()
        ds = ops.pool_bdl_fwd(dout, outer, pool, inner, "mean", float(pool))   # mean * pool = sum
        dxc = dD = None
        if Dskip is not None:
            zero_s = torch.zeros_like(ds)
            dxc = ops.bcast_skip_bdl_fwd(zero_s, dout, Dskip, outer, pool, inner)   # D[d] * dout
            dD = ops.rowdot_bdl(dout, xc).to(Dskip.dtype)
        return ds, dxc, dD, None, None, None
