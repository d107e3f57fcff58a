"""FastChannelVim SSM mixer -- host-side mirror of the reference
``mamba_ssm/modules/mamba_simple_channel_faster.py`` ``Mamba`` (:24-175 ctor, :176-420 forward).

Tokens are (spatial patch, image channel) pairs: ``L = rows * cols * tokens_per_patch``.  The scan runs over the
sequence pooled along the patch columns, in one of two orders (:225-256, :325-340):

* ``Channel-First``  sequence = (row, col, channel): pooled sequence (row, channel)       -> layout (rows, cols, tpp)
* ``Spatial-First``  sequence = (channel, row, col): pooled sequence (channel, row)       -> layout (tpp*rows, cols, 1)

Both are instances of the (outer, pool, inner) geometry every kernel of ``libfastvim_b200.so`` walks
(``include/fastvim_b200.h``: ``fv_geom``), so this module only picks the geometry and reuses ``mixer.Mamba``.
"""
from __future__ import annotations

import torch

from .mixer import Mamba as _SpatialMamba
from .ops import Geometry


class Mamba(_SpatialMamba):
    def __init__(self, d_model, d_state=16, d_conv=4, expand=2, dt_rank="auto", dt_min=0.001, dt_max=0.1,
                 dt_init="random", dt_scale=1.0, dt_init_floor=1e-4, conv_bias=True, bias=False, use_fast_path=False,
                 layer_idx=None, device=None, dtype=None, init_layer_scale=None, scanpath_type="rowwise",
                 token_size=None, use_norm_after_ssm=True, use_our_selective_scan=False, scan_order="Channel-First",
                 collapse_method="mean"):
        if scan_order not in ("Channel-First", "Spatial-First"):
            raise ValueError(f"scan_order must be 'Channel-First' or 'Spatial-First', got {scan_order!r}")
        # the reference asserts even grids "since we do compress and expand" (:66-71)
        assert token_size[0] % 2 == 0 and token_size[1] % 2 == 0, "num_of_rows / num_of_col need to be even"
        super().__init__(d_model, d_state=d_state, d_conv=d_conv, expand=expand, dt_rank=dt_rank, dt_min=dt_min,
                         dt_max=dt_max, dt_init=dt_init, dt_scale=dt_scale, dt_init_floor=dt_init_floor,
                         conv_bias=conv_bias, bias=bias, use_fast_path=use_fast_path, layer_idx=layer_idx, device=device,
                         dtype=dtype, init_layer_scale=init_layer_scale, scanpath_type=scanpath_type,
                         token_size=token_size, use_norm_after_ssm=use_norm_after_ssm,
                         use_our_selective_scan=use_our_selective_scan, collapse_method=collapse_method,
                         scaling_factor=1)
        self.scan_order = scan_order

    def channel_geometry(self, tokens_per_patch: int) -> Geometry:
        rows, cols, tpp = self.num_of_rows, self.num_of_col, int(tokens_per_patch)
        if self.scan_order == "Channel-First":
            # t = (r*cols + c)*tpp + ch lives at memory row t: strides (cols*tpp, tpp, 1)
            return Geometry(rows, cols, tpp, cols * tpp, tpp, 1)
        return Geometry(tpp * rows, cols, 1, cols, 1, 0)

    def forward(self, hidden_states, tokens_per_patch, inference_params=None):
        """hidden_states (B, rows*cols*tokens_per_patch, d_model) -> same shape   [reference :176-420]."""
        if inference_params is not None:
            raise NotImplementedError("autoregressive decode is outside the FastVim vision path")
        geom = self.channel_geometry(tokens_per_patch)
        if hidden_states.shape[1] != geom.L:
            raise ValueError(f"sequence length {hidden_states.shape[1]} != rows*cols*tokens_per_patch = {geom.L}")
        act_dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else hidden_states.dtype
        needs_grad = torch.is_grad_enabled() and (
            hidden_states.requires_grad or any(p.requires_grad for p in self.parameters()))
        if needs_grad:
            from . import autograd as fv_autograd

            if geom.inner != 1 or self.collapse_method == "max":
                # Channel-First ((rows, cols, tpp) pooling) and max pooling: not walked by the fused MixerFn backward kernels;
                # run operator by operator (conv / pool / scan / broadcast kernels, each with its backward)
                from . import composed

                return composed.mixer_forward_composed(self, hidden_states, act_dtype, outer=geom.outer,
                                                       pool=geom.pool, inner=geom.inner)
            out = fv_autograd.mixer_forward_train(self, hidden_states, geom, act_dtype)
        else:
            out = self._forward_inference(hidden_states.to(act_dtype), geom, act_dtype)
        if self.init_layer_scale is not None:
            out = out * self.gamma
        return out
