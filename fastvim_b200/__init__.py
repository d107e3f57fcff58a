"""fastvim_b200 -- B200-native (sm_100a) implementation of the FastVim SSM-block hot path.

Host side: PyTorch (memory, streams, distributed) calling ``libfastvim_b200.so`` through a C ABI
(``include/fastvim_b200.h``).  Public surface mirrors the reference's operator / module API:

    fastvim_b200.interface   selective_scan_fn, mamba_inner_fn_no_out_proj*, FastVim_mamba_inner_fn_*
    fastvim_b200.mixer       Mamba           (mamba_ssm.modules.mamba_simple_faster.Mamba)
    fastvim_b200.norm        RMSNorm, rms_norm_fn, layer_norm_fn  (mamba_ssm.ops.triton.layernorm)
    fastvim_b200.vision      PatchEmbed, Block, create_block, VisionMamba, factories (models/fastvim.py)
    fastvim_b200.compat      install(): registers the above under the reference's module paths
"""
__version__ = "0.1.0"
