"""Import shim for the reference's own kernel package ``faster_mamba_ssm`` (fastvim_kernel/mamba-1p1p1/faster_mamba_ssm)."""
