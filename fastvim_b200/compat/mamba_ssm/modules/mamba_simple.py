"""Import shim: the reference's plain (un-pooled) bidirectional Vim mixer ``mamba_ssm.modules.mamba_simple.Mamba``."""
from fastvim_b200.mixer_plain import Mamba  # noqa: F401
