"""``mamba_ssm.modules.mamba_simple_channel_faster_2dcompress`` (reference :24-425) -> the B200 2dcompress mixer."""
from fastvim_b200.mixer_channel_2dcompress import Mamba  # noqa: F401
from fastvim_b200.norm import RMSNorm, layer_norm_fn, rms_norm_fn  # noqa: F401
