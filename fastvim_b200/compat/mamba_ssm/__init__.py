"""Drop-in ``mamba_ssm`` namespace backed by fastvim_b200 (reference: mamba-1p1p1/mamba_ssm/__init__.py)."""
__version__ = "1.1.1+fastvim_b200"

from fastvim_b200.interface import mamba_inner_fn_no_out_proj, selective_scan_fn  # noqa: F401
from fastvim_b200.mixer import Mamba  # noqa: F401
