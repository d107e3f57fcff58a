"""``mamba_ssm.modules.mamba_simple_masked_faster`` (reference :21-325) -> the B200 FastMaskVim encoder mixer."""
from fastvim_b200.mixer_masked import Mamba_masked  # noqa: F401
