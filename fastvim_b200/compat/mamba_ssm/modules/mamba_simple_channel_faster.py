"""``mamba_ssm.modules.mamba_simple_channel_faster`` (reference :24-420) -> the B200 FastChannelVim mixer."""
from fastvim_b200.mixer_channel import Mamba  # noqa: F401
from fastvim_b200.norm import RMSNorm, layer_norm_fn, rms_norm_fn  # noqa: F401
