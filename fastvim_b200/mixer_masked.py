"""FastMaskVim encoder mixer -- host-side mirror of the reference ``Mamba_masked``
(``mamba_ssm/modules/mamba_simple_masked_faster.py:21-325`` and ``..._v2.py``; used by ``Block_masked``,
``models/mae/models_mamba_faster_mae_vimdecoder_v2.py:376-396``).

Same constructor keywords, parameter names (state-dict compatible with ``mixer.Mamba``) and
``forward(hidden_states, ids_keep, inference_params=None)`` signature.  ``hidden_states`` holds the kept tokens of an
MAE-masked image, ``ids_keep`` their original token ids; pooling scatters by ``ids_keep // num_of_col`` with the
constant divisor ``num_of_col``.  The body runs operator by operator on ``libfastvim_b200.so``
(``fastvim_b200.composed``): the masked path only exists for MAE pre-training, i.e. always under autograd.
"""
from __future__ import annotations

import torch

from . import composed
from .mixer import Mamba as _SpatialMamba


class Mamba_masked(_SpatialMamba):
    def __init__(self, d_model, d_state=16, d_conv=4, expand=2, dt_rank="auto", dt_min=0.001, dt_max=0.1,
                 dt_init="random", dt_scale=1.0, dt_init_floor=1e-4, conv_bias=True, bias=False, use_fast_path=False,
                 layer_idx=None, device=None, dtype=None, init_layer_scale=None, scanpath_type="rowwise",
                 token_size=None, use_norm_after_ssm=True, collapse_method="mean"):
        super().__init__(d_model, d_state=d_state, d_conv=d_conv, expand=expand, dt_rank=dt_rank, dt_min=dt_min,
                         dt_max=dt_max, dt_init=dt_init, dt_scale=dt_scale, dt_init_floor=dt_init_floor,
                         conv_bias=conv_bias, bias=bias, use_fast_path=use_fast_path, layer_idx=layer_idx, device=device,
                         dtype=dtype, init_layer_scale=init_layer_scale, scanpath_type=scanpath_type,
                         token_size=token_size, use_norm_after_ssm=use_norm_after_ssm, collapse_method=collapse_method,
                         scaling_factor=1)

    def forward(self, hidden_states, ids_keep, inference_params=None):
        """hidden_states (B, len_keep, d_model), ids_keep (B, len_keep) int64 -> (B, len_keep, d_model)."""
        if inference_params is not None:
            raise NotImplementedError("autoregressive decode is outside the FastVim vision path")
        if self.collapse_method != "mean":
            raise NotImplementedError("Mamba_masked defines collapse_method='mean' only "
                                      "(reference mamba_simple_masked_faster.py:212-216)")
        act_dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else hidden_states.dtype
        return composed.mixer_forward_composed(self, hidden_states, act_dtype, outer=self.num_of_rows,
                                               pool=self.num_of_col, ids_keep=ids_keep)
