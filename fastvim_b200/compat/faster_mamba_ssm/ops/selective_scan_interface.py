"""``faster_mamba_ssm.ops.selective_scan_interface`` (reference fastvim_kernel/mamba-1p1p1/faster_mamba_ssm/ops/
selective_scan_interface.py:129-159): the 6-tensor compressed ``selective_scan_fn(u, u_compressed, delta, A, B, C, ...)``."""
from fastvim_b200.interface import selective_scan_fn_compressed as selective_scan_fn  # noqa: F401
