"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the FastVim SSM-block hot path.

A plain PyTorch (CPU, fp32 or fp64, autograd-differentiable) restatement of the
reference's algorithm for the path ``BASELINE.json:north_star`` names.  Every function
cites the reference file:line it follows (paths relative to ``/root/reference``).

Pinning: ``oracle/gen_golden.py`` imports the unmodified reference (through
``oracle/ref_loader.py``) in the build container, checks every function below against
it on seeded inputs and writes the vectors under ``tests/golden/``;
``tests/test_oracle_golden.py`` re-checks the oracle against those vectors on every
run.  The one piece of arithmetic that lives outside the reference tree is the
depthwise causal conv (``causal-conv1d==1.1.3.post1``, un-vendored, reference
``README.md:43``): it is restated from the reference's own PyTorch fallback
(``mamba_ssm/modules/mamba_simple.py:302-303``) -- "conv parity pinned on the
reference's fallback form, not on the third-party kernel".

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The product package
(``fastvim_b200``) never does; it fails loudly without its CUDA library.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- scan
def selective_scan_oracle(u, delta, A, B, C, D=None, z=None, delta_bias=None,
                          delta_softplus=False, return_last_state=False,
                          compute_dtype=torch.float32):
    """Real-valued selective scan.

    Follows ``mamba_ssm/ops/selective_scan_interface.py:126-206`` (selective_scan_ref)
    and the kernel math ``csrc/selective_scan/selective_scan_fwd_kernel.cuh:147-265``:
        delta = softplus(delta + delta_bias)
        h[l]  = exp(delta[l] * A) * h[l-1] + delta[l] * B[l] * u[l],  h[-1] = 0
        y[l]  = sum_n C[l, n] * h[l, n]  (+ D * u)  (* silu(z))
    u, delta, z: (Bt, Dm, L); A: (Dm, N); B, C: (Dm, N) | (Bt, N, L) | (Bt, G, N, L).
    """
    dtype_in = u.dtype
    cd = torch.float64 if u.dtype == torch.float64 else compute_dtype   # fp64 inputs: fp64 gradient reference
    u_, delta_ = u.to(cd), delta.to(cd)
    if delta_bias is not None:
        delta_ = delta_ + delta_bias.to(cd)[..., None]
    if delta_softplus:
        delta_ = F.softplus(delta_)  # threshold 20, as fwd_kernel.cuh:153-156
    Bt, Dm, L = u_.shape
    N = A.shape[1]
    A_ = A.to(cd)

    def expand(M):  # -> (Bt, Dm, N, L)
        M = M.to(cd)
        if M.dim() == 2:
            return M[None, :, :, None].expand(Bt, Dm, N, L)
        if M.dim() == 3:
            return M[:, None].expand(Bt, Dm, N, L)
        G = M.shape[1]
        return M.repeat_interleave(Dm // G, dim=1)

    Bx, Cx = expand(B), expand(C)
    dA = torch.exp(delta_[:, :, None, :] * A_[None, :, :, None])  # (Bt, Dm, N, L)
    dBu = delta_[:, :, None, :] * Bx * u_[:, :, None, :]
    h = u_.new_zeros(Bt, Dm, N)
    ys = []
    for l in range(L):
        h = dA[..., l] * h + dBu[..., l]
        ys.append((h * Cx[..., l]).sum(-1))
    y = torch.stack(ys, dim=2)
    if D is not None:
        y = y + u_ * D.to(cd)[None, :, None]
    if z is not None:
        y = y * F.silu(z.to(cd))
    y = y.to(dtype_in)
    return (y, h) if return_last_state else y


def compressed_scan_oracle(u_full, u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                           return_last_state=False):
    """The 6-tensor "compressed" scan of the reference's own kernel package (``fastvim_kernel/mamba-1p1p1/
    faster_mamba_ssm/ops/selective_scan_interface.py:162-252`` ``selective_scan_ref``): the recurrence runs over the
    pooled ``u`` (Bt, Dm, Lc); its output is repeated ``cfac = L // Lc`` times along the sequence, the D skip uses the
    full-resolution ``u_full`` (Bt, Dm, L), then the optional ``silu(z)`` gate (full resolution)."""
    assert u_full.shape[2] % u.shape[2] == 0, "Compression factor must be integer"      # :191
    cfac = u_full.shape[2] // u.shape[2]
    dtype_in = u.dtype
    cd = torch.float64 if u.dtype == torch.float64 else torch.float32
    res = selective_scan_oracle(u.to(cd), delta.to(cd), A, B, C, None, None, delta_bias, delta_softplus,
                                return_last_state=True, compute_dtype=cd)
    y, last = res
    y = torch.repeat_interleave(y, cfac, dim=2)                                         # :244
    out = y if D is None else y + u_full.to(cd) * D.to(cd)[None, :, None]               # :245
    if z is not None:
        out = out * F.silu(z.to(cd))                                                    # :246-247
    out = out.to(dtype_in)
    return (out, last) if return_last_state else out


# --------------------------------------------------------------------------- conv
def causal_conv1d_oracle(x, weight, bias=None, activation="silu"):
    """Depthwise causal conv over the last axis.  x: (Bt, Dm, L), weight: (Dm, W).

    ``out[b,d,t] = act(bias[d] + sum_k weight[d,k] * x[b,d,t-(W-1)+k])`` with zero left
    padding -- the reference's own PyTorch form ``act(conv1d(x)[..., :seqlen])``
    (``mamba_ssm/modules/mamba_simple.py:302-303``; module definition
    ``mamba_simple_faster.py:89-97``: groups=d_inner, padding=d_conv-1).
    """
    Dm, W = weight.shape
    out = F.conv1d(x, weight[:, None, :].to(x.dtype),
                   None if bias is None else bias.to(x.dtype),
                   padding=W - 1, groups=Dm)[..., : x.shape[-1]]
    return F.silu(out) if activation in ("silu", "swish") else out


# --------------------------------------------------------------------------- pool
def pool_index(L: int, outer: int, pool: int, inner: int) -> Tensor:
    """Sequence position t -> pooled position, for a sequence viewed as
    (outer, pool, inner):  (t // (pool*inner)) * inner + t % inner.
    FastVim: (Hr, Wc, 1) ``mamba_simple_faster.py:287-297, 356``;
    ChannelVim: ``mamba_simple_channel_faster.py:225-256, 325-340``."""
    assert L == outer * pool * inner
    t = torch.arange(L)
    return (t // (pool * inner)) * inner + t % inner


def pool_oracle(xc, outer, pool, inner=1, method="mean", scaling_factor=1.0):
    """(Bt, Dm, L) -> (Bt, Dm, outer*inner); reference ``x.reshape(pre_x_shape).mean(3)``
    (``mamba_simple_faster.py:287-305``)."""
    Bt, Dm, L = xc.shape
    v = xc.reshape(Bt, Dm, outer, pool, inner)
    if method == "mean":
        p = v.mean(dim=3)
        if scaling_factor != 1:
            p = p * scaling_factor
    elif method == "max":
        p = v.max(dim=3).values
    else:
        raise ValueError(method)
    return p.reshape(Bt, Dm, outer * inner)


def broadcast_oracle(s, outer, pool, inner=1):
    """(Bt, Dm, Lp) -> (Bt, Dm, L): ``out.repeat_interleave(num_of_col, dim=2)``
    (``mamba_simple_faster.py:356``), generalised to (outer, pool, inner)."""
    Bt, Dm, Lp = s.shape
    v = s.reshape(Bt, Dm, outer, 1, inner).expand(Bt, Dm, outer, pool, inner)
    return v.reshape(Bt, Dm, outer * pool * inner)


# --------------------------------------------------------------------------- mixer
def mixer_oracle(hidden, p: Dict[str, Tensor], token_size: Sequence[int], *,
                 d_state=16, dt_rank=None, use_norm_after_ssm=True,
                 collapse_method="mean", scaling_factor=1.0, ln_eps=1e-5,
                 return_intermediates=False, layout=None, ids_keep=None):
    """FastVim ``Mamba.forward`` live branch, ``mamba_simple_faster.py:181-457``
    (the branch every shipped config takes, ``use_fast_path=False``, :269-453).

    hidden: (Bt, L, d_model); ``p`` uses the module's parameter names (:78-177):
    in_proj.weight, conv1d.weight/bias, x_proj.weight, dt_proj.weight/bias, A_log, D,
    the ``*_b`` twins, layernorm.weight/bias, out_proj.weight, [in_proj.bias,
    out_proj.bias, gamma].  token_size = (rows, cols) *as the mixer sees it* (already
    swapped on odd layers, ``models/fastvim.py:244-260``).
    """
    Bt, L, _ = hidden.shape
    rows, cols = token_size
    if ids_keep is not None:
        # FastMaskVim (``mamba_simple_masked_faster.py:167-325``): the sequence is the kept tokens; pooling is a
        # scatter-add by ``ids_keep // cols`` divided by the constant ``cols`` (:208-215, 376-416); broadcast is a
        # gather (:261-264).  The b-direction uses the SAME (un-flipped) row ids on the flipped sequence (:297-300).
        assert layout is None and collapse_method == "mean" and scaling_factor == 1
        return _masked_mixer_oracle(hidden, p, rows, cols, ids_keep, d_state, dt_rank, use_norm_after_ssm, ln_eps)
    # ``layout`` = (outer, pool, inner) generalises the pooling to the ChannelVim variants
    # (mamba_simple_channel_faster.py:225-256, 325-340): Channel-First (rows, cols, tpp), Spatial-First
    # (tpp*rows, cols, 1).  Default: FastVim (rows, cols, 1).  Below, ``rows`` = pooled length, ``cols`` = pool.
    outer, pool, inner = layout if layout is not None else (rows, cols, 1)
    assert L == outer * pool * inner
    rows, cols = outer * inner, pool
    D2 = p["in_proj.weight"].shape[0]
    Dm = D2 // 2
    R = dt_rank if dt_rank is not None else p["dt_proj.weight"].shape[1]
    N = d_state

    xz = F.linear(hidden, p["in_proj.weight"], p.get("in_proj.bias")).transpose(1, 2)  # :189-195
    A = -torch.exp(p["A_log"].float().to(hidden.dtype))      # :197
    A_b = -torch.exp(p["A_b_log"].float().to(hidden.dtype))  # :198
    x, z = xz.chunk(2, dim=1)                                # :270
    x_flip = x.flip([-1])                                    # :272
    xc = causal_conv1d_oracle(x, p["conv1d.weight"][:, 0], p.get("conv1d.bias"))            # :274-279
    xc_b = causal_conv1d_oracle(x_flip, p["conv1d_b.weight"][:, 0], p.get("conv1d_b.bias"))  # :280-285
    u = pool_oracle(xc, outer, pool, inner, collapse_method, scaling_factor)      # :287-305
    u_b = pool_oracle(xc_b, outer, pool, inner, collapse_method, scaling_factor)

    inter = {}

    def direction(u_c, xc_full, tag):
        x_dbl = F.linear(u_c.transpose(1, 2).reshape(Bt * rows, Dm), p[f"x_proj{tag}.weight"])  # :321-323
        dt, Bm, Cm = torch.split(x_dbl, [R, N, N], dim=-1)                                       # :324-326
        dt = (p[f"dt_proj{tag}.weight"] @ dt.t()).reshape(Dm, Bt, rows).permute(1, 0, 2)        # :328-334
        Bm = Bm.reshape(Bt, rows, N).transpose(1, 2)                                             # :336
        Cm = Cm.reshape(Bt, rows, N).transpose(1, 2)                                             # :337
        s = selective_scan_oracle(u_c, dt, A if tag == "" else A_b, Bm, Cm, D=None, z=None,
                                  delta_bias=p[f"dt_proj{tag}.bias"].float(), delta_softplus=True,
                                  compute_dtype=hidden.dtype if hidden.dtype == torch.float64
                                  else torch.float32)                                            # :343-354
        y = broadcast_oracle(s, outer, pool, inner)                                              # :356
        y = y + p["D" + tag].float().to(y.dtype)[None, :, None] * xc_full                        # :358
        if return_intermediates:
            inter[f"x_dbl{tag}"], inter[f"scan{tag}"] = x_dbl, s
        return y

    out = direction(u, xc, "")
    out_b = direction(u_b, xc_b, "_b")
    y = (out + out_b.flip([-1])).transpose(1, 2) / 2         # :438
    if use_norm_after_ssm:
        y = F.layer_norm(y, (Dm,), p["layernorm.weight"], p["layernorm.bias"], ln_eps)  # :437
    g = y * F.silu(z.transpose(1, 2))                        # :440 / :449
    o = F.linear(g, p["out_proj.weight"], p.get("out_proj.bias"))  # :442-444
    if "gamma" in p:
        o = o * p["gamma"]                                   # :455-456
    if return_intermediates:
        inter.update(xz=xz, xc=xc, xc_b=xc_b, u=u, u_b=u_b, gated=g)
        return o, inter
    return o


def _masked_mixer_oracle(hidden, p, rows, cols, ids_keep, d_state, dt_rank, use_norm_after_ssm, ln_eps):
    Bt, L, _ = hidden.shape
    Dm = p["in_proj.weight"].shape[0] // 2
    R = dt_rank if dt_rank is not None else p["dt_proj.weight"].shape[1]
    N = d_state
    cd = hidden.dtype if hidden.dtype == torch.float64 else torch.float32
    xz = F.linear(hidden, p["in_proj.weight"], p.get("in_proj.bias")).transpose(1, 2)   # :175-182
    x, z = xz.chunk(2, dim=1)                                                            # :194
    x_flip = x.flip([-1])                                                                # :196
    xc = causal_conv1d_oracle(x, p["conv1d.weight"][:, 0], p.get("conv1d.bias"))         # :198-203
    xc_b = causal_conv1d_oracle(x_flip, p["conv1d_b.weight"][:, 0], p.get("conv1d_b.bias"))  # :204-209
    rid = ids_keep // cols                                                               # :211

    def row_means(t):                                                                    # :376-416
        sums = torch.zeros(Bt, rows, Dm, dtype=cd)
        sums = sums.scatter_add(1, rid[:, :, None].expand(-1, -1, Dm), t.transpose(1, 2).to(cd))
        return (sums / cols).transpose(1, 2).to(t.dtype)

    def direction(t, tag):
        u_c = row_means(t)
        x_dbl = F.linear(u_c.transpose(1, 2).reshape(Bt * rows, Dm), p[f"x_proj{tag}.weight"])  # :232-234
        dt, Bm, Cm = torch.split(x_dbl, [R, N, N], dim=-1)
        dt = (p[f"dt_proj{tag}.weight"] @ dt.t()).reshape(Dm, Bt, rows).permute(1, 0, 2)        # :239-241
        Bm = Bm.reshape(Bt, rows, N).transpose(1, 2)
        Cm = Cm.reshape(Bt, rows, N).transpose(1, 2)
        A = -torch.exp(p["A_log" if tag == "" else "A_b_log"].float().to(hidden.dtype))
        s = selective_scan_oracle(u_c, dt, A, Bm, Cm, D=None, z=None, delta_bias=p[f"dt_proj{tag}.bias"].float(),
                                  delta_softplus=True, compute_dtype=cd)                          # :249-260
        y = torch.gather(s, 2, rid[:, None, :].expand(-1, Dm, -1))                                # :261-263
        return y + p["D" + tag].float().to(y.dtype)[None, :, None] * t                            # :264

    out, out_b = direction(xc, ""), direction(xc_b, "_b")
    y = (out + out_b.flip([-1])).transpose(1, 2) / 2                                     # :306
    if use_norm_after_ssm:
        y = F.layer_norm(y, (Dm,), p["layernorm.weight"], p["layernorm.bias"], ln_eps)   # :305
    o = F.linear(y * F.silu(z.transpose(1, 2)), p["out_proj.weight"], p.get("out_proj.bias"))
    if "gamma" in p:
        o = o * p["gamma"]
    return o


def block_interior_oracle(x, z, rows, cols, conv_w, conv_b, x_w, dt_w, dt_b, A_log, Dk, ln_w, ln_b,
                          eps=1e-5, scaling_factor=1.0):
    """The block interior alone -- ``mamba_simple_faster.py:269-453`` between the in_proj output and
    the out_proj input -- from direction-stacked parameters (the layout the fused CUDA kernel takes):
    x, z (Bt, L, D) token-major; conv_w (2, D, 4), conv_b (2, D), x_w (2, R+2N, D), dt_w (2, D, R),
    dt_b (2, D), A_log (2, D, N), Dk (2, D), ln_w / ln_b (D) or None.  Returns the gated (Bt, L, D).
    Implemented as ``mixer_oracle`` with identity in_proj / out_proj."""
    Dm = x.shape[-1]
    eye2, eye = torch.eye(2 * Dm, dtype=x.dtype), torch.eye(Dm, dtype=x.dtype)
    p = {"in_proj.weight": eye2, "out_proj.weight": eye,
         "conv1d.weight": conv_w[0][:, None, :], "conv1d_b.weight": conv_w[1][:, None, :],
         "x_proj.weight": x_w[0], "x_proj_b.weight": x_w[1],
         "dt_proj.weight": dt_w[0], "dt_proj_b.weight": dt_w[1],
         "dt_proj.bias": dt_b[0], "dt_proj_b.bias": dt_b[1],
         "A_log": A_log[0], "A_b_log": A_log[1], "D": Dk[0], "D_b": Dk[1]}
    if conv_b is not None:
        p["conv1d.bias"], p["conv1d_b.bias"] = conv_b[0], conv_b[1]
    if ln_w is not None:
        p["layernorm.weight"] = ln_w
        p["layernorm.bias"] = ln_b if ln_b is not None else torch.zeros_like(ln_w)
    return mixer_oracle(torch.cat([x, z], dim=-1), p, (rows, cols), d_state=A_log.shape[-1],
                        use_norm_after_ssm=ln_w is not None, scaling_factor=scaling_factor, ln_eps=eps)


def mamba_inner_oracle(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A,
                       B=None, C=None, D=None, delta_bias=None, delta_softplus=True,
                       has_z=True):
    """``mamba_inner_fn_no_out_proj[_withoutZ]`` semantics = ``mamba_inner_ref`` without
    the out_proj (``selective_scan_interface.py:1757-1810`` minus :1808-1810):
    conv(+SiLU) -> x_proj -> dt_proj -> selective_scan(u=conv, D, z).  xz: (Bt, 2Dm, L)
    (or (Bt, Dm, L) when ``has_z`` is False, the _withoutZ form :779-1016)."""
    L = xz.shape[-1]
    R = delta_proj_weight.shape[1]
    N = A.shape[-1]
    if has_z:
        x, z = xz.chunk(2, dim=1)
    else:
        x, z = xz, None
    xc = causal_conv1d_oracle(x, conv1d_weight[:, 0], conv1d_bias)
    Bt, Dm, _ = xc.shape
    x_dbl = F.linear(xc.transpose(1, 2).reshape(Bt * L, Dm), x_proj_weight)
    delta = (delta_proj_weight @ x_dbl[:, :R].t()).reshape(Dm, Bt, L).permute(1, 0, 2)
    if B is None:
        B = x_dbl[:, R:R + N].reshape(Bt, L, N).transpose(1, 2)
    if C is None:
        C = x_dbl[:, -N:].reshape(Bt, L, N).transpose(1, 2)
    return selective_scan_oracle(xc, delta, A, B, C, D, z=z, delta_bias=delta_bias,
                                 delta_softplus=delta_softplus)


def fastvim_inner_oracle(x, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A,
                         D, delta_bias, num_of_col, scaling_factor=1.0):
    """``FastVim_mamba_inner_fn_no_out_proj_withoutZ`` forward,
    ``selective_scan_interface.py:452-603``: conv -> mean-pool over ``num_of_col``
    -> x_proj/dt_proj -> scan on the pooled sequence -> repeat_interleave -> + D*conv."""
    Bt, Dm, L = x.shape
    rows = L // num_of_col
    R, N = delta_proj_weight.shape[1], A.shape[-1]
    xc = causal_conv1d_oracle(x, conv1d_weight[:, 0], conv1d_bias)                 # :496
    u = pool_oracle(xc, rows, num_of_col, 1, "mean", scaling_factor)               # :503-508
    x_dbl = F.linear(u.transpose(1, 2).reshape(Bt * rows, Dm), x_proj_weight)      # :512-514
    delta = (delta_proj_weight @ x_dbl[:, :R].t()).reshape(Dm, Bt, rows).permute(1, 0, 2)  # :515-519
    Bm = x_dbl[:, R:R + N].reshape(Bt, rows, N).transpose(1, 2)
    Cm = x_dbl[:, -N:].reshape(Bt, rows, N).transpose(1, 2)
    s = selective_scan_oracle(u, delta, A, Bm, Cm, None, None, delta_bias, True)   # :558-568
    return broadcast_oracle(s, rows, num_of_col, 1) + D[None, :, None] * xc        # :570-571


# --------------------------------------------------------------------------- norm / block / model
def add_norm_oracle(x, weight, bias=None, residual=None, eps=1e-5, is_rms=True,
                    out_dtype=None):
    """Fused residual-add + RMSNorm/LayerNorm, prenorm form -> (y, residual_out).
    ``mamba_ssm/ops/triton/layernorm.py:18-49`` (refs) / :66-121 (kernel): the sum is
    kept in fp32 (``residual_in_fp32``), y is returned in the activation dtype."""
    out_dtype = out_dtype or x.dtype
    cd = torch.float64 if x.dtype == torch.float64 else torch.float32
    r = x.to(cd) if residual is None else x.to(cd) + residual.to(cd)
    if is_rms:
        y = r * torch.rsqrt(r.square().mean(-1, keepdim=True) + eps) * weight.to(cd)
        if bias is not None:
            y = y + bias.to(cd)
    else:
        y = F.layer_norm(r, r.shape[-1:], weight.to(cd), None if bias is None else bias.to(cd), eps)
    return y.to(out_dtype), r


def rotate_tokens(h, rows, cols):
    """(Bt, rows*cols, C) row-major token order -> column-major, the odd-layer
    transpose of ``models/fastvim.py:192-200`` (inverse: call with (cols, rows))."""
    Bt, M, Cc = h.shape
    return h.reshape(Bt, rows, cols, Cc).transpose(1, 2).reshape(Bt, M, Cc)


def block_oracle(hidden, residual, p, layer_idx, token_size, *, rotate_every_block=True,
                 norm_eps=1e-5, rms_norm=True, **mixer_kw):
    """``Block.forward`` (``models/fastvim.py:146-212``) with drop_path = identity;
    the mixer of an odd layer is built with swapped token_size (:244-260)."""
    hs, residual = add_norm_oracle(hidden, p["norm.weight"], p.get("norm.bias"), residual,
                                   norm_eps, rms_norm)
    rows, cols = token_size
    odd = rotate_every_block and (layer_idx % 2 != 0)
    mp = {k[len("mixer."):]: v for k, v in p.items() if k.startswith("mixer.")}
    if odd:
        hs = rotate_tokens(hs, rows, cols)
        out = mixer_oracle(hs, mp, (cols, rows), **mixer_kw)
        out = rotate_tokens(out, cols, rows)
    else:
        out = mixer_oracle(hs, mp, (rows, cols), **mixer_kw)
    return out, residual


def patch_embed_oracle(images, weight, bias, patch):
    """``PatchEmbed.forward`` (``models/fastvim.py:72-103``), rowwise scan path:
    pad to a multiple of the patch, Conv2d(k=stride=patch), flatten to (Bt, L, C)."""
    H, W = images.shape[-2:]
    ph, pw = (patch - H % patch) % patch, (patch - W % patch) % patch
    if ph or pw:
        images = F.pad(images, (0, pw, 0, ph))
    x = F.conv2d(images, weight, bias, stride=patch)
    return x.flatten(2).transpose(1, 2), (x.shape[-2], x.shape[-1])


def fastvim_oracle(images, sd: Dict[str, Tensor], *, depth, patch=16, final_pool_type="mean",
                   rotate_every_block=True, norm_eps=1e-5, return_features=False, **mixer_kw):
    """``VisionMamba.forward`` (``models/fastvim.py:484-557``) for the shipped FastVim
    configs (rms_norm, residual_in_fp32, abs pos-embed, mean pool, no cls token)."""
    x, token_size = patch_embed_oracle(images, sd["patch_embed.proj.weight"],
                                       sd["patch_embed.proj.bias"], patch)
    if "pos_embed" in sd:
        x = x + sd["pos_embed"]
    hidden, residual = x, None
    for i in range(depth):
        pre = f"layers.{i}."
        p = {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
        hidden, residual = block_oracle(hidden, residual, p, i, token_size,
                                        rotate_every_block=rotate_every_block,
                                        norm_eps=norm_eps, **mixer_kw)
    hs, _ = add_norm_oracle(hidden, sd["norm_f.weight"], sd.get("norm_f.bias"), residual, norm_eps, True)
    if final_pool_type == "mean":
        feat = hs.mean(dim=1)
    elif final_pool_type == "none":
        feat = hs[:, -1, :]
    else:
        feat = hs
    if return_features:
        return feat
    return F.linear(feat, sd["head.weight"], sd["head.bias"])


# --------------------------------------------------------------------------- FastChannelVim model
def channelvim_oracle(images, sd: Dict[str, Tensor], *, depth, scan_order="Channel-First", patch=16,
                      rotate_every_block=True, norm_eps=1e-5, final_pool_type="mean", **mixer_kw):
    """``VisionMamba.forward`` of ``models/channel_wise_tokenization/models_channel_mamba_faster.py:590-683`` in eval mode
    (no hierarchical channel sampling, ``input_channel_order=None``): per-channel patch embedding with one shared
    ``Conv3d(1, E, (1, p, p))`` (:113-121, 180-184) + per-channel embedding, token order Channel-First
    ``(rows, cols, tpp)`` or Spatial-First ``(tpp, rows, cols)`` (:186-199), positional embedding repeated per channel
    (:621-630), ``Block`` (:206-336) with the odd-layer transposition (:304-331) and the mixer's pooled layout
    (``mamba_simple_channel_faster.py:225-256, 325-340``), final add + norm, mean pool, head."""
    Bt, Cn, H, W = images.shape
    w3 = sd["patch_embed.proj.weight"]                         # (E, 1, 1, p, p)
    x = F.conv3d(images[:, None], w3, sd.get("patch_embed.proj.bias"), stride=(1, patch, patch))   # (Bt, E, C, gh, gw)
    gh, gw = x.shape[-2:]
    ce = sd["patch_embed.channel_embed.weight"][:Cn]           # ids = arange(C)
    x = x + ce.t()[None, :, :, None, None]
    pe = sd.get("pos_embed")
    if scan_order == "Channel-First":
        x = x.permute(0, 3, 4, 2, 1).reshape(Bt, gh * gw * Cn, -1)
        if pe is not None:
            x = x + torch.repeat_interleave(pe, Cn, 1)
    else:
        x = x.permute(0, 2, 3, 4, 1).reshape(Bt, Cn * gh * gw, -1)
        if pe is not None:
            x = x + pe.expand(Cn, -1, -1).reshape(1, Cn * gh * gw, -1)
    hidden, residual = x, None
    t0, t1 = gh, gw
    for i in range(depth):
        pre = f"layers.{i}."
        p = {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
        hs, residual = add_norm_oracle(hidden, p["norm.weight"], p.get("norm.bias"), residual, norm_eps, True)
        mp = {k[len("mixer."):]: v for k, v in p.items() if k.startswith("mixer.")}
        odd = rotate_every_block and i % 2 != 0
        M = hs.shape[1]
        rows, cols = (t1, t0) if odd else (t0, t1)
        if odd:
            if scan_order == "Spatial-First":
                hs = hs.reshape(Bt, Cn, t0, t1, -1).transpose(2, 3).reshape(Bt, M, -1)
            else:
                hs = hs.reshape(Bt, t0, t1, Cn, -1).transpose(1, 2).reshape(Bt, M, -1)
        layout = (rows, cols, Cn) if scan_order == "Channel-First" else (Cn * rows, cols, 1)
        out = mixer_oracle(hs, mp, (rows, cols), layout=layout, **mixer_kw)
        if odd:
            if scan_order == "Spatial-First":
                out = out.reshape(Bt, Cn, t1, t0, -1).transpose(2, 3).reshape(Bt, M, -1)
            else:
                out = out.reshape(Bt, t1, t0, Cn, -1).transpose(1, 2).reshape(Bt, M, -1)
        hidden = out
    hs, _ = add_norm_oracle(hidden, sd["norm_f.weight"], sd.get("norm_f.bias"), residual, norm_eps, True)
    feat = hs.mean(dim=1) if final_pool_type == "mean" else hs[:, -1, :]
    return F.linear(feat, sd["head.weight"], sd["head.bias"])


# --------------------------------------------------------------------------- FastMaskVim encoder
def masked_blocks_oracle(hidden, sd: Dict[str, Tensor], ids_keep, token_size, *, depth, rotate_every_block=True,
                         norm_eps=1e-5, **mixer_kw):
    """A stack of ``Block_masked`` (``models/mae/models_mamba_faster_mae_vimdecoder_v2.py:279-402``) + the final
    add + RMSNorm of ``forward_encoder`` (:805-819).  Odd layers map the kept ids through the (h, w) -> (w, h) rotation
    (:320-328), re-sort the tokens by rotated id (:376-386), run the mixer built with the swapped token_size and restore
    the order (:392-396).  hidden (Bt, len_keep, E); ids_keep (Bt, len_keep) sorted original token ids."""
    H, W = token_size
    idx = torch.arange(H * W)
    rot = (idx % W) * H + idx // W
    residual = None
    for i in range(depth):
        pre = f"layers.{i}."
        p = {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
        hs, residual = add_norm_oracle(hidden, p["norm.weight"], p.get("norm.bias"), residual, norm_eps, True)
        mp = {k[len("mixer."):]: v for k, v in p.items() if k.startswith("mixer.")}
        odd = rotate_every_block and i % 2 != 0
        ids = ids_keep
        if odd:
            ids = rot[ids_keep]
            order = torch.argsort(ids)
            ids = torch.gather(ids, 1, order)
            hs = torch.gather(hs, 1, order[..., None].expand(-1, -1, hs.shape[-1]))
        ts = (W, H) if odd else (H, W)
        out = mixer_oracle(hs, mp, ts, ids_keep=ids, **mixer_kw)
        if odd:
            out = torch.gather(out, 1, torch.argsort(order, -1)[..., None].expand(-1, -1, out.shape[-1]))
        hidden = out
    hs, _ = add_norm_oracle(hidden, sd["norm_f.weight"], sd.get("norm_f.bias"), residual, norm_eps, True)
    return hs


def masked_encoder_oracle(images, sd: Dict[str, Tensor], ids_keep, *, depth, patch=16, **kw):
    """``MaskedAutoencoderViM.forward_encoder`` (:776-819) with the kept ids given: patch embedding + positional
    embedding, gather of the kept tokens, ``masked_blocks_oracle``."""
    x, token_size = patch_embed_oracle(images, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], patch)
    x = x + sd["pos_embed"]
    x = torch.gather(x, 1, ids_keep[..., None].expand(-1, -1, x.shape[-1]))
    return masked_blocks_oracle(x, sd, ids_keep, token_size, depth=depth, **kw)


# --------------------------------------------------------------------------- init helper
def random_mixer_params(d_model, *, d_state=16, d_conv=4, expand=2, seed=0,
                        dtype=torch.float32) -> Dict[str, Tensor]:
    """Random parameters with the reference's shapes and init distributions
    (``mamba_simple_faster.py:78-177``); used by tests so both sides load the same dict."""
    g = torch.Generator().manual_seed(seed)
    Dm = expand * d_model
    R = math.ceil(d_model / 16)

    def U(shape, a):
        return (torch.rand(shape, generator=g) * 2 - 1) * a

    p = {"in_proj.weight": U((2 * Dm, d_model), d_model ** -0.5),
         "out_proj.weight": U((d_model, Dm), Dm ** -0.5),
         "layernorm.weight": 1 + 0.1 * U((Dm,), 1.0), "layernorm.bias": 0.1 * U((Dm,), 1.0)}
    for tag in ("", "_b"):
        p[f"conv1d{tag}.weight"] = U((Dm, 1, d_conv), d_conv ** -0.5)
        p[f"conv1d{tag}.bias"] = U((Dm,), d_conv ** -0.5)
        p[f"x_proj{tag}.weight"] = U((R + 2 * d_state, Dm), Dm ** -0.5)
        p[f"dt_proj{tag}.weight"] = U((Dm, R), R ** -0.5)
        dt = torch.exp(torch.rand(Dm, generator=g) * (math.log(0.1) - math.log(1e-3)) + math.log(1e-3)).clamp(min=1e-4)
        p[f"dt_proj{tag}.bias"] = dt + torch.log(-torch.expm1(-dt))
        p["A" + tag + "_log"] = torch.log(torch.arange(1, d_state + 1, dtype=torch.float32)).repeat(Dm, 1) \
            + 0.1 * U((Dm, d_state), 1.0)
        p["D" + tag] = 1 + 0.1 * U((Dm,), 1.0)
    return {k: v.to(dtype) for k, v in p.items()}
