"""CPU: argument validation of the host-side mirrors and of the C ABI (status + message, nothing launched).

The reference raises through TORCH_CHECK / Python asserts (selective_scan.cpp:235-278, mamba_simple_channel_faster.py:66-71,
selective_scan_interface.py:503-508); the mirrors keep the same conditions."""
import ctypes

import pytest
import torch

from fastvim_b200 import _lib, interface
from fastvim_b200.mixer import Mamba
from fastvim_b200.mixer_channel import Mamba as ChannelMamba
from fastvim_b200.mixer_channel_2dcompress import Mamba as Channel2dMamba
from fastvim_b200.mixer_masked import Mamba_masked
from fastvim_b200.ops import Geometry


def test_mixer_constructor_conditions():
    with pytest.raises(NotImplementedError):
        Mamba(32, d_conv=3, token_size=[4, 6], layer_idx=0)
    with pytest.raises(AssertionError):          # "num_of_rows / num_of_col need to be even" (reference :66-71)
        ChannelMamba(32, token_size=[3, 6], layer_idx=0)
    with pytest.raises(ValueError):
        ChannelMamba(32, token_size=[4, 6], layer_idx=0, scan_order="Diagonal")
    with pytest.raises(NotImplementedError):     # reference prints "not implemented yet" for Spatial-First
        Channel2dMamba(32, token_size=[4, 6], layer_idx=0, scan_order="Spatial-First")
    with pytest.raises(ValueError):
        Channel2dMamba(32, token_size=[4, 6], layer_idx=None)
    m = Mamba(32, token_size=[4, 6], layer_idx=0)
    assert m.dt_rank == 2 and m.d_inner == 64 and m.A_log.shape == (64, 16)
    assert getattr(m.A_log, "_no_weight_decay") and getattr(m.D, "_no_weight_decay") and m.dt_proj.bias._no_reinit


def test_channel_geometries():
    m = ChannelMamba(32, token_size=[4, 6], layer_idx=0, scan_order="Channel-First")
    g = m.channel_geometry(3)
    assert (g.outer, g.pool, g.inner, g.L, g.Lp) == (4, 6, 3, 72, 12)
    assert (g.stride_outer, g.stride_pool, g.stride_inner) == (18, 3, 1)          # memory order == sequence order
    g = ChannelMamba(32, token_size=[4, 6], layer_idx=0, scan_order="Spatial-First").channel_geometry(3)
    assert (g.outer, g.pool, g.inner, g.Lp) == (12, 6, 1, 12)
    g = Geometry.grid(4, 6, rotated=True)
    assert (g.stride_outer, g.stride_pool) == (1, 4)                               # column-major walk of a 6 x 4 grid
    with pytest.raises(ValueError):
        m(torch.zeros(1, 70, 32), 3)                                               # 70 != rows * cols * tpp


def test_operator_api_argument_errors():
    u = torch.zeros(1, 4, 8)
    A = torch.zeros(4, 2)
    with pytest.raises(NotImplementedError):
        interface.selective_scan_fn(u, u, torch.zeros(4, 2, dtype=torch.complex64), A, A)
    fn = interface.FastVim_mamba_inner_fn_no_out_proj_withoutZ
    args = (u, torch.zeros(4, 1, 4), None, torch.zeros(5, 4), torch.zeros(4, 1), A)
    with pytest.raises(NotImplementedError):                                       # only 'mean' is defined (:503-508)
        fn(*args, collapse_method="max")
    with pytest.raises(ValueError):
        fn(*args, num_of_col=3)                                                    # 8 % 3 != 0
    with pytest.raises(ValueError):
        fn(*args, num_of_col=4, pre_x_shape=(-1, 4, 4, 2))
    with pytest.raises(_lib.FastVimLibraryError):                                  # no CPU path
        interface.selective_scan_fn(u, u, A, torch.zeros(1, 2, 8), torch.zeros(1, 2, 8))


def test_masked_mixer_conditions():
    m = Mamba_masked(32, token_size=[4, 6], layer_idx=0, collapse_method="max")
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 5, 32), torch.zeros(1, 5, dtype=torch.long))
    m = Mamba_masked(32, token_size=[4, 6], layer_idx=0)
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 5, 32), torch.zeros(1, 5, dtype=torch.long), inference_params=object())


def test_c_abi_reports_bad_arguments_without_launching():
    l = _lib.lib()
    g = _lib.fv_geom(1, 64, 2, 2, 1, 2, 1, 0)
    null = None
    # operator-API backward: missing buffers / unsupported d_state
    rc = l.fv_selective_scan_bwd(0, 1, 4, 8, 16, 1, null, null, null, null, null, null, null, null, 0, null, null, null, null,
                                 null, null, null, null, null, null, 0, null)
    assert rc != 0 and b"null pointer" in l.fv_last_error()
    assert l.fv_selective_scan_bwd_workspace_bytes(2, 4, 128, 16) == 0             # one chunk: no checkpoints
    assert l.fv_selective_scan_bwd_workspace_bytes(2, 4, 129, 16) == 2 * 4 * 2 * 16 * 4
    # short scan backward: Lp <= 16 and d_state 16 only
    assert l.fv_scan_bwd_short_supported(ctypes.byref(g), 16) == 1
    assert l.fv_scan_bwd_short_supported(ctypes.byref(g), 8) == 0
    g_long = _lib.fv_geom(1, 64, 17, 2, 1, 2, 1, 0)
    assert l.fv_scan_bwd_short_supported(ctypes.byref(g_long), 16) == 0
    assert l.fv_scan_bwd_planes(ctypes.byref(g)) == 2 and l.fv_scan_bwd_planes(ctypes.byref(g_long)) == 1
    # streaming gate backward: plain geometry, dim % 64 == 0, even strides
    assert l.fv_gate_bwd_stream_supported(ctypes.byref(g), 128, 64) == 1
    assert l.fv_gate_bwd_stream_supported(ctypes.byref(g), 127, 64) == 0
    g_ch = _lib.fv_geom(1, 64, 2, 2, 3, 6, 3, 1)
    assert l.fv_gate_bwd_stream_supported(ctypes.byref(g_ch), 128, 64) == 0
    g_odd = _lib.fv_geom(1, 96, 2, 2, 1, 2, 1, 0)
    assert l.fv_gate_bwd_stream_supported(ctypes.byref(g_odd), 192, 96) == 0
    rc = l.fv_causal_conv1d_bwd(0, 1, 4, 8, null, 32, 8, null, null, 1, null, null, 32, 8, null, null, null)
    assert rc != 0 and b"null pointer" in l.fv_last_error()


def test_random_masking_properties_cpu():
    """random_masking (reference models_mamba_faster_mae_vimdecoder_v2.py:740-774): kept ids sorted, mask / ids_restore
    consistent, x_masked is the gather of the kept tokens."""
    from fastvim_b200.vision_masked import Block_masked, random_masking

    torch.manual_seed(0)
    x = torch.randn(4, 24, 5)
    xm, mask, ids_restore, ids_keep = random_masking(x, 0.75)
    assert xm.shape == (4, 6, 5) and ids_keep.shape == (4, 6)
    assert torch.equal(ids_keep, ids_keep.sort(dim=1).values)
    assert torch.equal(xm, torch.gather(x, 1, ids_keep[..., None].expand(-1, -1, 5)))
    assert torch.equal(mask.sum(1), torch.full((4,), 18.0))
    kept = torch.zeros(4, 24).scatter_(1, ids_keep, 1.0)
    assert torch.equal(mask, 1.0 - kept)                          # 0 = keep, 1 = remove, in original token order
    assert torch.equal(torch.gather(ids_restore, 1, ids_keep), torch.arange(6)[None].expand(4, -1))
    # odd-layer id rotation: rotating a 4 x 6 grid twice (with swapped sizes) is the identity
    r1, r2 = Block_masked.compute_rotate_indices(4, 6), Block_masked.compute_rotate_indices(6, 4)
    assert torch.equal(r2[r1], torch.arange(24))
