"""``mamba_ssm.ops.selective_scan_interface`` (reference :105-123, :1652-1753) -> fastvim_b200.interface."""
from fastvim_b200.interface import (  # noqa: F401
    FastVim_mamba_inner_fn_no_out_proj_withoutZ,
    mamba_inner_fn_no_out_proj,
    mamba_inner_fn_no_out_proj_withoutZ,
    selective_scan_fn,
)
