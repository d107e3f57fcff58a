"""Plain bidirectional Vim mixer -- host-side mirror of the reference ``mamba_ssm.modules.mamba_simple.Mamba``
(``mamba-1p1p1/mamba_ssm/modules/mamba_simple.py:40-407``), the reference's BASELINE block (no pooling): the FastMaskVim
MAE decoder is a stack of these (``models/mae/models_mamba_faster_mae_vimdecoder_v2.py``), so the reference's MAE model
file only imports over the compat shims when this class exists.

Same constructor keywords, parameter names and ``forward(hidden_states, inference_params=None)`` as the reference.  The
computation is the FastVim mixer with a pooling window of ONE token -- conv + SiLU, scan over all L tokens in both
directions, D skip, direction average, LayerNorm, SiLU(z) gate (reference :219-262) -- so it runs on the same kernels
through ``fastvim_b200.mixer.Mamba`` with the geometry ``(outer = L, pool = 1)`` (four-launch path; forward and backward).
The reference's no-norm fast-path branch gates each direction with z inside the scan (:264-299), which equals the
gated average computed here."""
from __future__ import annotations

from .mixer import Mamba as _PooledMamba
from .ops import Geometry


class Mamba(_PooledMamba):
    def __init__(self, d_model, d_state=16, d_conv=4, expand=2, dt_rank="auto", dt_min=0.001, dt_max=0.1,
                 dt_init="random", dt_scale=1.0, dt_init_floor=1e-4, conv_bias=True, bias=False, use_fast_path=True,
                 layer_idx=None, device=None, dtype=None, init_layer_scale=None, use_norm_after_ssm=True):
        # use_fast_path only selects between two equivalent formulations in the reference; nothing to switch here
        super().__init__(d_model, d_state=d_state, d_conv=d_conv, expand=expand, dt_rank=dt_rank, dt_min=dt_min,
                         dt_max=dt_max, dt_init=dt_init, dt_scale=dt_scale, dt_init_floor=dt_init_floor,
                         conv_bias=conv_bias, bias=bias, use_fast_path=False, layer_idx=layer_idx, device=device,
                         dtype=dtype, init_layer_scale=init_layer_scale, token_size=[1, 1],
                         use_norm_after_ssm=use_norm_after_ssm, collapse_method="mean", scaling_factor=1)
        self.use_fast_path = use_fast_path
        self._seqlen = 1

    def geometry(self, rotated: bool = False) -> Geometry:
        return Geometry.grid(self._seqlen, 1, False)

    def forward(self, hidden_states, inference_params=None):
        self._seqlen = int(hidden_states.shape[1])
        return super().forward(hidden_states, inference_params=inference_params, rotated=False)
