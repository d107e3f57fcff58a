"""Training path of the mixer (autograd.Function over the CUDA forward/backward kernels)."""
from __future__ import annotations


def mixer_forward_train(mixer, hidden_states, geom, act_dtype):
    raise NotImplementedError(
        "fastvim_b200: the backward kernels of the mixer are not built yet; run under "
        "torch.no_grad() / inference_mode (there is deliberately no PyTorch fallback)")


def add_norm_train(x, weight, bias, residual, eps, prenorm, residual_in_fp32, is_rms):
    raise NotImplementedError(
        "fastvim_b200: the backward kernel of the fused add+norm is not built yet; run under "
        "torch.no_grad() (there is deliberately no PyTorch fallback)")


def selective_scan_train(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state):
    raise NotImplementedError(
        "fastvim_b200: selective_scan backward kernel is not built yet; run under torch.no_grad()")
