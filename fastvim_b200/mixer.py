"""FastVim SSM mixer -- host-side mirror of the reference ``Mamba`` module.

Same constructor keywords, parameter names/shapes (state-dict compatible) and ``forward``
signature as ``mamba_ssm/modules/mamba_simple_faster.py:27-179, 181-457`` of the reference, so
it drops into the reference's ``models/fastvim.py`` ``Block`` unchanged.  The body between
``in_proj`` and ``out_proj`` runs on the B200 kernels of ``libfastvim_b200.so``:

    in_proj (GEMM) -> K1 conv+SiLU+pool (both directions) -> x_proj (batched GEMM)
      -> K2a bidirectional pooled scan (dt_proj + softplus fused) -> K2b broadcast + D skip +
      LayerNorm + SiLU(z) gate -> out_proj (GEMM)

in token-major layout (no (B, D, L) transposes, no flips, no repeat_interleave).  The extra
``rotated=True`` keyword lets ``fastvim_b200.vision.Block`` fold the odd-layer token rotation
(reference ``models/fastvim.py:192-210``) into kernel addressing instead of two copies; when
the reference's own Block calls this module it has already permuted the tokens and
``rotated`` stays False.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import autograd as fv_autograd
from . import ops
from .ops import Geometry


# Use the one-launch fused block interior (fv_block_fwd) whenever the configuration qualifies; the
# four-launch path (K1 -> x_proj -> K2a -> K2b) covers everything else (fp32, long sequences, channel
# layouts, max pooling).  Module-level switches so tests and tools/kbench.py can compare the two.
FUSED_BLOCK = os.environ.get("FASTVIM_FUSED_BLOCK", "1") != "0"
# in_proj / out_proj on the hand-written tcgen05 / TMEM / TMA GEMM (csrc/gemm_tc.cu) when the shape qualifies
# (bf16, no bias, K % 64 == 0, N % 64 == 0, weight block fits shared memory); cuBLAS through F.linear otherwise.
TC_GEMM = os.environ.get("FASTVIM_TC_GEMM", "1") != "0"
# inference: fold the next block's residual add + RMSNorm into this block's out_proj epilogue (fv_gemm_out_norm) when
# d_model fits one accumulator (FastVim-T); "0" = separate fv_gemm_bf16_tn + fv_add_norm_fwd launches
FUSED_OUT_NORM = os.environ.get("FASTVIM_FUSED_OUT_NORM", "1") != "0"
# channel layouts (inner > 1): fv_conv_pool_w_fwd + fv_gate_w_fwd instead of the generic K1 / K2b; "0" = generic kernels
GROUP_KERNELS = os.environ.get("FASTVIM_GROUP_KERNELS", "1") != "0"
# inference: fv_block_fwd_signal + fv_gemm_out_norm_flow (the out_proj GEMM starts on finished images while the block kernel's
# second round is still running); "0" = the plain pair
FLOW = os.environ.get("FASTVIM_FLOW", "1") != "0"


def linear(x, w, b):
    """y = x @ w.T (+ b).  x (..., K), w (N, K)."""
    if (TC_GEMM and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.is_cuda and x.stride(-1) == 1
            and ops.gemm_supported(x.numel() // x.shape[-1], w.shape[0], w.shape[1])):
        x2 = x.reshape(-1, x.shape[-1])
        if x2.stride(0) % 8 == 0 and x2.data_ptr() % 16 == 0 and w.is_contiguous():
            return ops.gemm_bf16_tn(x2, w, bias=b).view(*x.shape[:-1], w.shape[0])
    return F.linear(x, w, b)


class Mamba(nn.Module):
    def __init__(self, d_model, d_state=16, d_conv=4, expand=2, dt_rank="auto", dt_min=0.001,
                 dt_max=0.1, dt_init="random", dt_scale=1.0, dt_init_floor=1e-4, conv_bias=True,
                 bias=False, use_fast_path=False, layer_idx=None, device=None, dtype=None,
                 init_layer_scale=None, scanpath_type="rowwise", token_size=None,
                 use_norm_after_ssm=True, use_our_selective_scan=False, collapse_method="mean",
                 scaling_factor=1):
        factory_kwargs = {"device": device, "dtype": dtype}
        super().__init__()
        if d_conv != 4:
            raise NotImplementedError("fastvim_b200 kernels implement d_conv=4 (every FastVim config)")
        self.d_model, self.d_state, self.d_conv, self.expand = d_model, d_state, d_conv, expand
        self.d_inner = int(expand * d_model)
        self.dt_rank = math.ceil(d_model / 16) if dt_rank == "auto" else dt_rank
        # The reference's two branches compute the same function EXCEPT use_fast_path=True with collapse 'mean' and no
        # post-SSM norm, where its fused branch drops the silu(z) gate (mamba_simple_faster.py:262-267).
        # No shipped config sets use_fast_path; this mirror implements the live branch and refuses that one combination.
        if use_fast_path and not use_norm_after_ssm and collapse_method == "mean":
            raise NotImplementedError(
                "fastvim_b200 implements the live branch (silu(z) gate kept); the reference's use_fast_path=True + "
                "use_norm_after_ssm=False branch omits the gate and is not reproduced (INTEGRATION.md)")
        self.use_fast_path = use_fast_path
        self.layer_idx = layer_idx
        self.use_our_selective_scan = use_our_selective_scan
        self.num_of_rows, self.num_of_col = token_size[0], token_size[1]
        self.scanpath_type = scanpath_type
        self.scaling_factor = scaling_factor
        self.collapse_method = collapse_method
        self.init_layer_scale = init_layer_scale
        if init_layer_scale is not None:
            self.gamma = nn.Parameter(init_layer_scale * torch.ones(d_model), requires_grad=True)
        self.in_proj = nn.Linear(d_model, self.d_inner * 2, bias=bias, **factory_kwargs)
        self.use_norm_after_ssm = use_norm_after_ssm
        if use_norm_after_ssm:
            self.layernorm = nn.LayerNorm(self.d_inner, **factory_kwargs)
        self.activation = "silu"
        self.act = nn.SiLU()

        def make_dir(special_dt_init):
            conv = nn.Conv1d(self.d_inner, self.d_inner, bias=conv_bias, kernel_size=d_conv,
                             groups=self.d_inner, padding=d_conv - 1, **factory_kwargs)
            x_proj = nn.Linear(self.d_inner, self.dt_rank + d_state * 2, bias=False, **factory_kwargs)
            dt_proj = nn.Linear(self.dt_rank, self.d_inner, bias=True, **factory_kwargs)
            if special_dt_init:
                # reference init, mamba_simple_faster.py:111-130.  The reference applies it to the
                # forward direction only; dt_proj_b keeps nn.Linear's default init (:166-171).
                dt_init_std = self.dt_rank ** -0.5 * dt_scale
                if dt_init == "constant":
                    nn.init.constant_(dt_proj.weight, dt_init_std)
                elif dt_init == "random":
                    nn.init.uniform_(dt_proj.weight, -dt_init_std, dt_init_std)
                else:
                    raise NotImplementedError
                dt = torch.exp(torch.rand(self.d_inner, **factory_kwargs) * (math.log(dt_max) - math.log(dt_min))
                               + math.log(dt_min)).clamp(min=dt_init_floor)
                inv_dt = dt + torch.log(-torch.expm1(-dt))
                with torch.no_grad():
                    dt_proj.bias.copy_(inv_dt)
                dt_proj.bias._no_reinit = True
            A_log = torch.log(torch.arange(1, d_state + 1, dtype=torch.float32, device=device)).repeat(self.d_inner, 1)
            A_log = nn.Parameter(A_log.contiguous())
            A_log._no_weight_decay = True
            Dp = nn.Parameter(torch.ones(self.d_inner, device=device))
            Dp._no_weight_decay = True
            return conv, x_proj, dt_proj, A_log, Dp

        # creation order follows the reference so that seeded inits line up (:89-171)
        self.conv1d, self.x_proj, self.dt_proj, self.A_log, self.D = make_dir(True)
        self.conv1d_b, self.x_proj_b, self.dt_proj_b, self.A_b_log, self.D_b = make_dir(False)
        self.out_proj = nn.Linear(self.d_inner, d_model, bias=bias, **factory_kwargs)
        self.pre_x_shape = (-1, self.d_inner, self.num_of_rows, self.num_of_col)
        self._pack_cache = {}

    # ------------------------------------------------------------------ parameter packing
    def _packed(self, act_dtype):
        """Direction-stacked, kernel-ready copies of the small parameters, cached until any
        parameter changes (``_version`` bump) -- inference pays the packing once."""
        plist = list(self.parameters())
        key = (act_dtype,) + tuple((p.data_ptr(), p._version) for p in plist if p is not None)
        hit = self._pack_cache.get("k")
        if hit == key:
            return self._pack_cache["v"]
        with torch.no_grad():
            f32 = torch.float32
            pk = {
                "conv_w": torch.stack([self.conv1d.weight[:, 0], self.conv1d_b.weight[:, 0]]).to(f32).contiguous(),
                "conv_b": None if self.conv1d.bias is None else
                torch.stack([self.conv1d.bias, self.conv1d_b.bias]).to(f32).contiguous(),
                "x_w_t": torch.stack([self.x_proj.weight.t(), self.x_proj_b.weight.t()]).to(act_dtype).contiguous(),
                "x_w": torch.stack([self.x_proj.weight, self.x_proj_b.weight]).to(act_dtype).contiguous(),
                "dt_w": torch.stack([self.dt_proj.weight, self.dt_proj_b.weight]).to(f32).contiguous(),
                "dt_b": torch.stack([self.dt_proj.bias, self.dt_proj_b.bias]).to(f32).contiguous(),
                "A_log": torch.stack([self.A_log, self.A_b_log]).to(f32).contiguous(),
                # A = -exp(A_log) (reference :197-198), evaluated once per parameter version instead of per image and thread
                "A_neg": (-torch.exp(torch.stack([self.A_log, self.A_b_log]).to(f32))).contiguous(),
                "D": torch.stack([self.D, self.D_b]).to(f32).contiguous(),
                "in_w": self.in_proj.weight.to(act_dtype).contiguous(),
                "in_b": None if self.in_proj.bias is None else self.in_proj.bias.to(act_dtype),
                "out_w": self.out_proj.weight.to(act_dtype).contiguous(),
                "out_b": None if self.out_proj.bias is None else self.out_proj.bias.to(act_dtype),
                "ln_w": self.layernorm.weight.to(f32).contiguous() if self.use_norm_after_ssm else None,
                "ln_b": self.layernorm.bias.to(f32).contiguous() if self.use_norm_after_ssm else None,
            }
            if act_dtype == torch.bfloat16 and pk["x_w"].is_cuda:
                pk["x_w_packed"] = ops.block_pack_xproj(pk["x_w"])   # fragment order for the fused block kernel
        self._pack_cache = {"k": key, "v": pk}
        return pk

    def invalidate(self):
        """Drop the packed parameter copies.  ``_packed`` keys on ``Parameter._version``, which optimizer steps replayed
        from a CUDA graph never bump: call this after graph-replayed training and before an eval forward."""
        self._pack_cache = {}

    def geometry(self, rotated: bool = False) -> Geometry:
        return Geometry.grid(self.num_of_rows, self.num_of_col, rotated)

    # ------------------------------------------------------------------ forward
    def forward(self, hidden_states, inference_params=None, rotated: bool = False):
        """hidden_states (B, L, d_model) -> (B, L, d_model)   [reference :181-457]."""
        if inference_params is not None:
            raise NotImplementedError("autoregressive decode is outside the FastVim vision path")
        act_dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else hidden_states.dtype
        geom = self.geometry(rotated)
        needs_grad = torch.is_grad_enabled() and (
            hidden_states.requires_grad or any(p.requires_grad for p in self.parameters()))
        if needs_grad and self.collapse_method == "max":
            # max pooling under autograd (the reference's live branch differentiates x.reshape(...).max(3).values,
            # mamba_simple_faster.py:299-305): operator-by-operator path; it returns the gamma-scaled output itself.
            # That path walks the sequence in memory order, so a geometry-folded odd layer (rotated=True: memory is the
            # (cols, rows) row-major grid) is permuted physically here, as the reference's Block does (models/fastvim.py:192-210).
            from . import composed

            B, L, dm = hidden_states.shape
            rows, cols = self.num_of_rows, self.num_of_col
            h = hidden_states
            if rotated:
                h = h.reshape(B, cols, rows, dm).transpose(1, 2).reshape(B, L, dm)
            out = composed.mixer_forward_composed(self, h, act_dtype, outer=rows, pool=cols)
            if rotated:
                out = out.reshape(B, rows, cols, dm).transpose(1, 2).reshape(B, L, dm)
            return out
        if needs_grad:
            out = fv_autograd.mixer_forward_train(self, hidden_states, geom, act_dtype)
        else:
            out = self._forward_inference(hidden_states.to(act_dtype), geom, act_dtype)
        if self.init_layer_scale is not None:
            out = out * self.gamma
        return out

    def out_norm_fusable(self, h, act_dtype) -> bool:
        """Can this mixer's out_proj carry the next residual add + RMSNorm in its epilogue (``fv_gemm_out_norm``)?"""
        return (FUSED_OUT_NORM and TC_GEMM and act_dtype == torch.bfloat16 and h.is_cuda and self.out_proj.bias is None
                and self.init_layer_scale is None and not self.training
                and ops.gemm_out_norm_supported(h.numel() // h.shape[-1], self.d_model, self.d_inner))

    def forward_out_norm(self, hidden_states, rotated, residual, norm_w, eps, want_residual=True, flow=None):
        """Inference only: ``(rmsnorm(residual + mixer(h)) * norm_w, residual + mixer(h))`` with the add + norm folded into the
        out_proj GEMM epilogue.  The caller checks ``out_norm_fusable`` first."""
        geom = self.geometry(rotated)
        act_dtype = torch.bfloat16
        return self._forward_inference(hidden_states.to(act_dtype), geom, act_dtype,
                                       out_norm=(residual, norm_w, eps, want_residual), flow=flow)

    def _linear_out(self, y, pk, out_norm, flow=None):
        if out_norm is None:
            return linear(y, pk["out_w"], pk["out_b"])
        residual, norm_w, eps, want_residual = out_norm
        return ops.gemm_out_norm(y, pk["out_w"], residual, norm_w, eps, want_residual, flow=flow)

    def _forward_inference(self, h, geom, act_dtype, out_norm=None, flow=None):
        pk = self._packed(act_dtype)
        B, L, _ = h.shape
        D, R, N = self.d_inner, self.dt_rank, self.d_state
        xz = linear(h, pk["in_w"], pk["in_b"])                          # (B, L, 2D) token-major  [a2]
        x, z = xz[..., :D], xz[..., D:]
        eps = self.layernorm.eps if self.use_norm_after_ssm else 1e-5
        if (FUSED_BLOCK and self.collapse_method == "mean"
                and ops.block_fwd_supported(geom, B, D, xz.dtype, R, N)):
            # one launch for [a3-a9]: the image's x stays resident in shared memory (csrc/block_fwd.cu)
            # flow = (sync, launch_index): the block kernel publishes images as they complete and the out_proj GEMM, launched
            # programmatically dependent, consumes them while the second round of images is still being computed
            use_flow = (flow is not None and out_norm is not None and FLOW
                        and ops.block_fwd_signal_supported(geom, B, D, R, N))
            y = ops.block_fwd(x, z, geom, pk["conv_w"], pk["conv_b"], pk["x_w"], pk["dt_w"], pk["dt_b"],
                              pk["A_neg"], pk["D"], pk["ln_w"], pk["ln_b"], eps, float(self.scaling_factor), R, N,
                              a_is_log=False, xproj_w_packed=pk.get("x_w_packed"), signal=flow if use_flow else None)
            return self._linear_out(y, pk, out_norm, (flow[0], L, flow[1]) if use_flow else None)   # [a10] (+ next [a11])
        if GROUP_KERNELS and ops.conv_pool_w_supported(geom, B, D, xz.dtype):
            # channel layouts (inner > 1): staged conv + pool that also writes the D-skip term, streaming gate
            u, w = ops.conv_pool_w_fwd(x, geom, pk["conv_w"], pk["conv_b"], pk["D"], float(self.scaling_factor),
                                       self.collapse_method)
            xdbl = ops.x_proj(u, pk["x_w"], TC_GEMM)
            s = ops.scan_fwd(u, xdbl, geom, R, N, pk["dt_w"], pk["dt_b"], pk["A_log"], a_is_log=True)
            y = ops.gate_w_fwd(w, z, s, geom, pk["ln_w"], pk["ln_b"], eps)
            return self._linear_out(y, pk, out_norm)
        u = ops.conv_pool_fwd(x, geom, pk["conv_w"], pk["conv_b"], float(self.scaling_factor),
                              self.collapse_method)                      # (2, B, Lp, D)         [a3-a5]
        xdbl = ops.x_proj(u, pk["x_w"], TC_GEMM)                         # (2, B*Lp, R+2N)        [a6]
        s = ops.scan_fwd(u, xdbl, geom, R, N, pk["dt_w"], pk["dt_b"], pk["A_log"], a_is_log=True)  # [a6-a7]
        y = ops.gate_fwd(x, z, s, geom, pk["conv_w"], pk["conv_b"], pk["D"], pk["ln_w"], pk["ln_b"],
                         self.layernorm.eps if self.use_norm_after_ssm else 1e-5)             # [a8-a9]
        return self._linear_out(y, pk, out_norm)                         # [a10]
