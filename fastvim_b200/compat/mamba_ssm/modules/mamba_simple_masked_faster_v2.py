"""``mamba_ssm.modules.mamba_simple_masked_faster_v2`` (reference :21-335) -> the B200 FastMaskVim encoder mixer."""
from fastvim_b200.mixer_masked import Mamba_masked  # noqa: F401
