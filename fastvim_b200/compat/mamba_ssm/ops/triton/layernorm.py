"""``mamba_ssm.ops.triton.layernorm`` (reference ops/triton/layernorm.py:402-512) -> the CUDA add+norm
kernels of fastvim_b200 (no Triton)."""
from fastvim_b200.norm import RMSNorm, layer_norm_fn, rms_norm_fn  # noqa: F401
