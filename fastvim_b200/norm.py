"""Fused residual-add + RMSNorm/LayerNorm -- mirror of the reference's
``mamba_ssm/ops/triton/layernorm.py`` public API (``RMSNorm`` :515-536, ``rms_norm_fn`` :507-512,
``layer_norm_fn`` :492-504) on the CUDA kernel ``fv_add_norm_fwd`` (no Triton)."""
from __future__ import annotations

import torch

from . import autograd as fv_autograd
from . import ops


def _norm(x, weight, bias, residual, eps, prenorm, residual_in_fp32, is_rms):
    needs_grad = torch.is_grad_enabled() and any(
        t is not None and t.requires_grad for t in (x, weight, bias, residual))
    if needs_grad:
        return fv_autograd.add_norm_train(x, weight, bias, residual, eps, prenorm, residual_in_fp32, is_rms)
    res_dtype = None if residual is None else residual.dtype
    if residual is not None and residual.dtype != torch.float32:
        residual = residual.float()
    w = weight if weight.dtype == torch.float32 else weight.float()
    b = bias if bias is None or bias.dtype == torch.float32 else bias.float()
    y, res_out, _, _ = ops.add_norm_fwd(x, residual, w, b, eps, is_rms, want_residual=prenorm)
    if not prenorm:
        return y
    if not residual_in_fp32:
        # reference (layernorm.py:_layer_norm_fwd): residual_out keeps residual.dtype, or x.dtype without a residual,
        # unless an fp32 residual stream was requested
        res_out = res_out.to(res_dtype if res_dtype is not None else x.dtype)
    return y, res_out


def layer_norm_fn(x, weight, bias, residual=None, eps=1e-6, prenorm=False, residual_in_fp32=False,
                  is_rms_norm=False):
    return _norm(x, weight, bias, residual, eps, prenorm, residual_in_fp32, is_rms_norm)


def rms_norm_fn(x, weight, bias, residual=None, prenorm=False, residual_in_fp32=False, eps=1e-6):
    return _norm(x, weight, bias, residual, eps, prenorm, residual_in_fp32, True)


class RMSNorm(torch.nn.Module):
    def __init__(self, hidden_size, eps=1e-5, device=None, dtype=None):
        super().__init__()
        self.eps = eps
        self.weight = torch.nn.Parameter(torch.empty(hidden_size, device=device, dtype=dtype))
        self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        torch.nn.init.ones_(self.weight)

    def forward(self, x, residual=None, prenorm=False, residual_in_fp32=False):
        return rms_norm_fn(x, self.weight, self.bias, residual=residual, eps=self.eps, prenorm=prenorm,
                           residual_in_fp32=residual_in_fp32)
