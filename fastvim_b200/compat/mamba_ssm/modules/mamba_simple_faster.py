"""``mamba_ssm.modules.mamba_simple_faster`` (reference mamba_ssm/modules/mamba_simple_faster.py:27-457)
-> the B200 mixer.  Same class name, constructor keywords, parameter names and forward signature."""
from fastvim_b200.mixer import Mamba  # noqa: F401
from fastvim_b200.norm import RMSNorm, layer_norm_fn, rms_norm_fn  # noqa: F401
