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


def test_round2_entry_points_validate_arguments_without_launching():
    """The round-2 entry points (general tcgen05 GEMM, out_proj + norm epilogue, dataflow pair, streaming K1 / K2b) reject bad
    arguments with a status code and a message before anything touches the device (runs without a GPU)."""
    l = _lib.lib()
    null = None
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    # general GEMM
    assert l.fv_gemm_bf16(128, 64, 64, 0, null, 64, 0, null, 64, _lib.FV_BF16, null, 64, 1, null) != 0
    assert b"null pointer" in l.fv_last_error()
    assert l.fv_gemm_bf16(128, 64, 64, 0, p, 60, 0, p, 64, _lib.FV_BF16, p, 64, 1, null) != 0      # lda % 8
    assert b"16-byte" in l.fv_last_error()
    assert l.fv_gemm_bf16(128, 64, 64, 0, p, 64, 0, p, 64, _lib.FV_BF16, p, 64, 2, null) != 0      # split-K needs fp32 planes
    assert b"split-K" in l.fv_last_error()
    assert l.fv_gemm_bf16(128, 64, 64, 0, p, 64, 0, p, 64, 7, p, 64, 1, null) != 0                 # unknown output type
    assert l.fv_gemm_bf16(128, 64, 128, 0, p, 128, 0, p, 128, _lib.FV_F32, p, 64, 3, null) != 0    # 3 splits > 2 k-blocks
    assert b"splits" in l.fv_last_error()
    assert l.fv_gemm_bf16_batched(0, 128, 64, 64, 0, p, 64, 8192, 0, p, 64, 4096, _lib.FV_BF16, p, 64, 8192, 1, null) != 0
    assert l.fv_gemm_bf16_batched(2, 128, 64, 64, 0, p, 64, 8191, 0, p, 64, 4096, _lib.FV_BF16, p, 64, 8192, 1, null) != 0
    assert l.fv_gemm_bf16_batched(2, 128, 64, 128, 0, p, 128, 16384, 0, p, 128, 8192, _lib.FV_F32, p, 64, 8192, 2, null) != 0   # split-K: ACC only
    # split heuristic: fills the SMs, never leaves an empty split
    for Mo, No, K in [(768, 192, 50176), (3072, 768, 25088), (44, 1536, 1792), (128, 64, 64)]:
        s = l.fv_gemm_bf16_splits(Mo, No, K)
        kb = (K + 63) // 64
        assert 1 <= s <= 64 and ((kb + s - 1) // s) * (s - 1) < kb
    # out_proj + add + RMSNorm epilogue: the row must fit one accumulator
    assert l.fv_gemm_out_norm_supported(50176, 192, 384) == 1
    assert l.fv_gemm_out_norm_supported(50176, 384, 768) == 0
    assert l.fv_gemm_out_norm_supported(50176, 192, 100) == 0
    assert l.fv_gemm_out_norm(128, 192, 384, null, 384, null, 384, null, 192, null, null, 1e-5, null, 192, null) != 0
    assert b"null pointer" in l.fv_last_error()
    assert l.fv_gemm_out_norm_flow(392, 192, 384, p, 384, p, 384, p, 192, null, p, 1e-5, p, 192, null, 196, 0, null) != 0
    assert b"sync" in l.fv_last_error()
    # dataflow form of the block kernel: one-CTA-per-image configurations only, flags + positive epoch required
    g_t = _lib.fv_geom(4, 384, 14, 14, 1, 14, 1, 0)
    g_b = _lib.fv_geom(4, 1536, 14, 14, 1, 14, 1, 0)
    assert l.fv_block_fwd_signal_supported(ctypes.byref(g_t), _lib.FV_BF16, 12, 16) == 1
    assert l.fv_block_fwd_signal_supported(ctypes.byref(g_b), _lib.FV_BF16, 48, 16) == 0          # cluster kernel: no flags
    assert l.fv_block_fwd_signal(ctypes.byref(g_t), _lib.FV_BF16, p, p, 768, 150528, p, null, p, null, p, p, p, 0, 12, 16, p,
                                 null, null, 1e-5, 1.0, p, 384, 75264, null, 1, null) != 0
    assert b"done_flags" in l.fv_last_error()
    # streaming K1 / K2b: bf16; channel layouts need their inner slots to fit
    g_ch = _lib.fv_geom(2, 768, 14, 14, 8, 112, 8, 1)
    assert l.fv_conv_pool_w_supported(ctypes.byref(g_ch), _lib.FV_BF16) == 1
    assert l.fv_conv_pool_w_supported(ctypes.byref(g_ch), _lib.FV_F32) == 0
    assert l.fv_conv_pool_w_supported(ctypes.byref(g_t), _lib.FV_BF16) == 1                       # plain geometry: staged kernels
    g_wide = _lib.fv_geom(2, 4096, 14, 14, 8, 112, 8, 1)
    assert l.fv_conv_pool_w_supported(ctypes.byref(g_wide), _lib.FV_BF16) == 0
    assert l.fv_conv_pool_w_fwd(ctypes.byref(g_ch), _lib.FV_BF16, p, 768, 1204224, p, null, 1.0, 0, null, p, p, null) != 0
    assert b"Dskip" in l.fv_last_error()
    assert l.fv_gate_w_fwd(ctypes.byref(g_ch), _lib.FV_F32, p, p, 1536, 2408448, p, null, null, 1e-5, p, 768, 1204224, null) != 0
    assert b"bf16" in l.fv_last_error()
    # PDL switches return the previous setting
    prev = l.fv_set_pdl(0)
    assert l.fv_set_pdl(prev) == 0
    assert l.fv_set_pdl_all(1) in (0, 1) and l.fv_set_pdl_all(0) == 1


def test_x_proj_and_gemm_helpers_fall_back_on_cpu():
    """Host logic of the round-2 wrappers that must not touch the library for CPU tensors (stand-in based tests rely on it)."""
    from fastvim_b200 import ops

    u = torch.randn(2, 3, 5, 16)
    xw = torch.randn(2, 7, 16)
    got = ops.x_proj(u, xw)
    assert torch.allclose(got, torch.bmm(u.view(2, 15, 16), xw.transpose(1, 2)))
    assert not ops.gemm_bf16_ok(torch.zeros(4, 8, dtype=torch.bfloat16))            # CPU tensor
    assert not ops.conv_pool_w_supported(Geometry(4, 6, 3, 18, 3, 1), 2, 64, torch.float32)
