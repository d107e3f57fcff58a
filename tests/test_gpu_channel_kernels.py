"""Channel-layout kernels (inner > 1): fv_conv_pool_w_fwd / fv_gate_w_fwd against the generic K1 / K2b kernels and a direct
torch evaluation of the reference formulas (mamba_simple_channel_faster.py:258-289, 325-340, 400-420); the chunk-parallel
pooled scan at dt_rank 24 against the one-thread-per-chain kernel."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _inputs(Bt, geom, D, seed=0):
    torch.manual_seed(seed)
    L = geom.L
    xz = torch.randn(Bt, L, 2 * D).bfloat16().cuda()
    cw, cb = (torch.randn(2, D, 4) * 0.5).cuda(), (torch.randn(2, D) * 0.5).cuda()
    Dk = (1.0 + 0.2 * torch.randn(2, D)).cuda()
    return xz, cw, cb, Dk


def _conv_dirs(x, cw, cb):
    """x (B, L, D) fp32 in SEQUENCE order -> silu(conv_f), silu(conv_b) in sequence order (conv_b = conv on the flipped
    sequence, flipped back)."""
    xt = x.transpose(1, 2)
    L = x.shape[1]
    f = F.silu(F.conv1d(F.pad(xt, (3, 0)), cw[0][:, None, :], cb[0], groups=x.shape[2]))
    bwd = F.silu(F.conv1d(F.pad(xt.flip(-1), (3, 0)), cw[1][:, None, :], cb[1], groups=x.shape[2])).flip(-1)
    return f.transpose(1, 2)[:, :L], bwd.transpose(1, 2)[:, :L]


@pytest.mark.parametrize("rows,cols,tpp,D,Bt", [(14, 14, 8, 768, 3), (4, 6, 3, 64, 2), (6, 4, 2, 384, 5), (2, 9, 4, 136, 2)])
@pytest.mark.parametrize("mode", ["mean", "max"])
def test_conv_pool_w_vs_generic_and_torch(rows, cols, tpp, D, Bt, mode):
    from fastvim_b200 import ops

    geom = ops.Geometry(rows, cols, tpp, cols * tpp, tpp, 1)      # Channel-First: memory order == sequence order
    assert ops.conv_pool_w_supported(geom, Bt, D, torch.bfloat16)
    assert not ops.conv_pool_w_supported(geom, Bt, D, torch.float32)                                # bf16 only
    xz, cw, cb, Dk = _inputs(Bt, geom, D)
    x = xz[..., :D]
    u, w = ops.conv_pool_w_fwd(x, geom, cw, cb, Dk, 1.5, mode)
    u_ref = ops.conv_pool_fwd(x, geom, cw, cb, 1.5, mode)
    scale = u_ref.float().abs().max().item()
    assert (u.float() - u_ref.float()).abs().max().item() / scale < 1e-2
    xf, xb = _conv_dirs(x.float().cpu(), cw.cpu(), cb.cpu())
    w_want = 0.5 * (Dk[0].cpu() * xf + Dk[1].cpu() * xb)
    assert (w.float().cpu() - w_want).abs().max().item() / w_want.abs().max().item() < 1e-2
    # pooled output against torch as well: mean / max over the `cols` axis of the (rows, cols, tpp) sequence
    pooled = []
    for t in (xf, xb):
        t5 = t.view(Bt, rows, cols, tpp, D)
        pooled.append((t5.mean(2) * 1.5 if mode == "mean" else t5.max(2).values).reshape(Bt, rows * tpp, D))
    u_want = torch.stack(pooled)
    assert (u.float().cpu() - u_want).abs().max().item() / u_want.abs().max().item() < 1e-2


@pytest.mark.parametrize("norm", [True, False])
def test_gate_w_path_matches_generic_path(norm):
    """conv_pool_w -> scan -> gate_w against conv_pool -> scan -> gate (which re-evaluates the convolutions) at the
    FastChannelVim-S JUMP-CP geometry."""
    from fastvim_b200 import ops

    rows, cols, tpp, D, Bt, R, N = 14, 14, 8, 768, 2, 24, 16
    geom = ops.Geometry(rows, cols, tpp, cols * tpp, tpp, 1)
    xz, cw, cb, Dk = _inputs(Bt, geom, D, seed=1)
    x, z = xz[..., :D], xz[..., D:]
    lw = (1.0 + 0.2 * torch.randn(D)).cuda() if norm else None
    lb = (0.2 * torch.randn(D)).cuda() if norm else None
    xw = (torch.randn(2, R + 2 * N, D) * D ** -0.5).bfloat16().cuda()
    dtw = (torch.randn(2, D, R) * R ** -0.5).cuda()
    dtb = (torch.rand(2, D) * 4.0 - 5.0).cuda()
    A_log = torch.log(torch.arange(1, N + 1).float()).repeat(2, D, 1).cuda()
    u, w = ops.conv_pool_w_fwd(x, geom, cw, cb, Dk)
    xdbl = ops.x_proj(u, xw)
    s = ops.scan_fwd(u, xdbl, geom, R, N, dtw, dtb, A_log, a_is_log=True)
    y = ops.gate_w_fwd(w, z, s, geom, lw, lb, 1e-5)
    y_ref = ops.gate_fwd(x, z, s, geom, cw, cb, Dk, lw, lb, 1e-5)
    err = (y.float() - y_ref.float()).abs().max().item() / y_ref.float().abs().max().item()
    assert err < 2e-2, err


def test_scan_fwd_chunked_rank24_matches_plain_kernel():
    """FastChannelVim-S: 112 pooled rows, dt_rank 24.  An 8-image launch takes the chunk-parallel kernel (too few chains to fill
    the SMs), a 128-image launch the one-thread-per-chain kernel: same inputs, same result up to summation order."""
    from fastvim_b200 import ops

    torch.manual_seed(2)
    D, R, N, Lp, Bs = 768, 24, 16, 112, 8
    geom = ops.Geometry(14, 14, 8, 112, 8, 1)
    u = (torch.randn(2, 128, Lp, D) * 0.5).bfloat16().cuda()
    xdbl = (torch.randn(2, 128 * Lp, R + 2 * N) * 0.5).bfloat16().cuda()
    dtw = (torch.randn(2, D, R) * R ** -0.5).cuda()
    dtb = (torch.rand(2, D) * 4.0 - 5.0).cuda()
    A_log = torch.log(torch.arange(1, N + 1).float()).repeat(2, D, 1).cuda()
    s_plain = ops.scan_fwd(u, xdbl, geom, R, N, dtw, dtb, A_log, a_is_log=True)[:, :Bs]
    us = u[:, :Bs].contiguous()
    xs = xdbl.view(2, 128, Lp, -1)[:, :Bs].reshape(2, Bs * Lp, -1).contiguous()
    s_chunk = ops.scan_fwd(us, xs, geom, R, N, dtw, dtb, A_log, a_is_log=True)
    err = (s_chunk - s_plain).abs().max().item() / s_plain.abs().max().item()
    assert err < 1e-4, err


@pytest.mark.parametrize("rows,cols,D,Bt,rot", [(14, 14, 384, 4, False), (14, 14, 384, 3, True), (16, 64, 384, 1, False),
                                                (128, 128, 384, 1, False), (6, 10, 72, 2, False)])
def test_plain_geometry_w_path_matches_generic_path(rows, cols, D, Bt, rot):
    """inner == 1 (FastVim; 2048^2 takes the cluster conv + pool): conv_pool_w -> gate_w against conv_pool -> gate."""
    from fastvim_b200 import ops

    geom = ops.Geometry.grid(rows, cols, rot)
    assert ops.conv_pool_w_supported(geom, Bt, D, torch.bfloat16)
    xz, cw, cb, Dk = _inputs(Bt, geom, D, seed=3)
    x, z = xz[..., :D], xz[..., D:]
    u, w = ops.conv_pool_w_fwd(x, geom, cw, cb, Dk, 2.0)
    u_ref = ops.conv_pool_fwd(x, geom, cw, cb, 2.0)
    assert torch.equal(u, u_ref)                         # same kernel, one more output
    torch.manual_seed(4)
    s = torch.randn(2, Bt, geom.Lp, D).cuda()
    lw, lb = (1.0 + 0.2 * torch.randn(D)).cuda(), (0.2 * torch.randn(D)).cuda()
    y = ops.gate_w_fwd(w, z, s, geom, lw, lb, 1e-5)
    y_ref = ops.gate_fwd(x, z, s, geom, cw, cb, Dk, lw, lb, 1e-5)
    err = (y.float() - y_ref.float()).abs().max().item() / y_ref.float().abs().max().item()
    assert err < 2e-2, err
