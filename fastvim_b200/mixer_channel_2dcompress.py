"""FastChannelVim "2D compress" mixer -- host-side mirror of the reference
``mamba_ssm/modules/mamba_simple_channel_faster_2dcompress.py`` ``Mamba`` (ctor :24-175, forward :176-425).

Channel-First token order ``(row, col, channel)``; which axis the tokens are pooled over depends on the layer
(:224-255, broadcast :329-343):

* every third layer (``(layer_idx + 1) % 3 == 0``) -- "channelwise scan": mean over ALL ``rows * cols`` patches, the
  scan runs over the ``tokens_per_patch`` channel positions        -> geometry ``(1, rows*cols, tpp)``
* otherwise -- row scan with columns and channels pooled together   -> geometry ``(rows, cols*tpp, 1)``

Both are ``(outer, pool, inner)`` geometries of the kernels in ``libfastvim_b200.so``; this module only picks them.
The reference defines the variant for ``scan_order="Channel-First"`` only (Spatial-First prints "not implemented yet").
"""
from __future__ import annotations

from .mixer_channel import Mamba as _ChannelMamba
from .ops import Geometry


class Mamba(_ChannelMamba):
    def __init__(self, d_model, d_state=16, d_conv=4, expand=2, dt_rank="auto", dt_min=0.001, dt_max=0.1,
                 dt_init="random", dt_scale=1.0, dt_init_floor=1e-4, conv_bias=True, bias=False, use_fast_path=False,
                 layer_idx=None, device=None, dtype=None, init_layer_scale=None, token_size=None,
                 use_norm_after_ssm=True, use_our_selective_scan=False, scan_order="Channel-First",
                 collapse_method="mean"):
        if scan_order != "Channel-First":
            raise NotImplementedError("the 2dcompress mixer is defined for scan_order='Channel-First' only "
                                      "(reference mamba_simple_channel_faster_2dcompress.py:224, 259, 329)")
        if layer_idx is None:
            raise ValueError("the 2dcompress mixer needs layer_idx (it selects the pooled axis)")
        super().__init__(d_model, d_state=d_state, d_conv=d_conv, expand=expand, dt_rank=dt_rank, dt_min=dt_min,
                         dt_max=dt_max, dt_init=dt_init, dt_scale=dt_scale, dt_init_floor=dt_init_floor,
                         conv_bias=conv_bias, bias=bias, use_fast_path=use_fast_path, layer_idx=layer_idx, device=device,
                         dtype=dtype, init_layer_scale=init_layer_scale, token_size=token_size,
                         use_norm_after_ssm=use_norm_after_ssm, use_our_selective_scan=use_our_selective_scan,
                         scan_order=scan_order, collapse_method=collapse_method)

    def channel_geometry(self, tokens_per_patch: int) -> Geometry:
        rows, cols, tpp = self.num_of_rows, self.num_of_col, int(tokens_per_patch)
        if (self.layer_idx + 1) % 3 == 0:      # channelwise scan: pool over every patch, sequence = channels
            return Geometry(1, rows * cols, tpp, rows * cols * tpp, tpp, 1)
        return Geometry(rows, cols * tpp, 1, cols * tpp, 1, 0)
