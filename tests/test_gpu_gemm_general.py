"""General tcgen05 GEMM (csrc/gemm_tc2.cu: MN-major operands, split-K fp32 planes, ragged shapes) against fp64 matmuls
of the same bf16 operands -- the backward GEMMs the reference leaves to cuBLAS (selective_scan_interface.py:698-737)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(rows, cols, scale):
    return (torch.randn(rows, cols) * scale).bfloat16().cuda()


CASES = [  # (Mo, No, K, a_mn, b_mn, out_f32)
    (1000, 768, 1536, False, False, False),    # TN, ragged M
    (392, 44, 384, False, False, False),       # x_proj: ragged N (bf16 output with a padded row pitch)
    (1792, 80, 1536, False, False, True),      # x_proj, fp32 out
    (3000, 768, 3072, False, True, False),     # dgrad in_proj (FastVim-B): dX = dY . W
    (2000, 384, 192, False, True, False),      # dgrad out_proj (FastVim-T)
    (768, 192, 6000, True, True, True),        # wgrad in_proj (FastVim-T), split-K, ragged K
    (3072, 768, 2500, True, True, True),       # wgrad in_proj (FastVim-B)
    (44, 1536, 1792, True, True, True),        # wgrad x_proj: 44 output rows
    (1536, 12, 1792, True, True, True),        # wgrad dt_proj: 12 output columns... (No = 12 -> fp32 pitch 48 B)
    (200, 136, 72, True, False, False),        # A transposed only, everything ragged
]


@pytest.mark.parametrize("Mo,No,K,a_mn,b_mn,out_f32", CASES)
def test_gemm_general_vs_fp64(Mo, No, K, a_mn, b_mn, out_f32):
    from fastvim_b200 import ops

    torch.manual_seed(Mo + No + K)
    A = _mk(Mo, K, 0.5)
    B = _mk(No, K, K ** -0.5)
    want = A.double() @ B.double().t()
    a_in = A.t().contiguous() if a_mn else A
    b_in = B.t().contiguous() if b_mn else B
    # 16-byte row pitches: pad the stored matrices where the logical width is not a multiple of 8
    def pad(t):
        if t.shape[1] % 8 == 0:
            return t
        buf = torch.zeros(t.shape[0], (t.shape[1] + 7) // 8 * 8, dtype=t.dtype, device=t.device)
        buf[:, :t.shape[1]] = t
        return buf[:, :t.shape[1]]
    a_in, b_in = pad(a_in), pad(b_in)
    out = None
    if not out_f32 and No % 8:
        out = torch.full((Mo, (No + 7) // 8 * 8), 7.0, dtype=torch.bfloat16, device="cuda")[:, :No]
    c = ops.gemm_bf16(a_in, b_in, a_mn=a_mn, b_mn=b_mn, out_f32=out_f32, out=out)
    assert c.shape == (Mo, No) and c.dtype == (torch.float32 if out_f32 else torch.bfloat16)
    err = (c.double() - want).abs().max().item() / want.abs().max().item()
    assert err < (2e-5 if out_f32 else 4e-3), err
    if out is not None:    # padding columns up to the 16-byte boundary are either untouched or zero (TMA clips in 16-byte units)
        padc = out._base[:, No:]
        assert torch.all((padc == 7.0) | (padc == 0.0))


@pytest.mark.parametrize("acc", [True, False])
def test_gemm_general_split_counts_agree(acc, monkeypatch):
    """Every split count gives the same fp32 result up to summation order, in both split-K forms: partial planes +
    fv_reduce_planes (deterministic) and TMA reduction stores into one zeroed plane (FV_F32_ACC)."""
    from fastvim_b200 import ops

    monkeypatch.setattr(ops, "WGRAD_ACC", acc)
    monkeypatch.setattr(ops, "WGRAD_ACC_MAX_SPLITS", 64)
    torch.manual_seed(3)
    dy, x = _mk(4096, 256, 0.5), _mk(4096, 192, 0.1)
    want = dy.double().t() @ x.double()
    for splits in (1, 2, 7, 16):
        c = ops.gemm_bf16(dy, x, a_mn=True, b_mn=True, out_f32=True, splits=splits)
        err = (c.double() - want).abs().max().item() / want.abs().max().item()
        assert err < 2e-5, (splits, err)


@pytest.mark.parametrize("M,N,K", [(1000, 192, 384), (392, 64, 128), (50176, 192, 384), (300, 256, 768), (129, 128, 1536)])
def test_gemm_out_norm_vs_fp64(M, N, K):
    """out_proj + residual add + RMSNorm in the GEMM epilogue (fv_gemm_out_norm) against fp64: the new residual is
    res + A W^T from the fp32 accumulator (never rounded to bf16), y its RMS normalisation (layernorm.py:66-121 semantics)."""
    from fastvim_b200 import ops

    torch.manual_seed(M + N)
    a = (torch.randn(M, K) * 0.5).bfloat16().cuda()
    w = (torch.randn(N, K) * K ** -0.5).bfloat16().cuda()
    res = torch.randn(M, N).cuda()
    nw = (1.0 + 0.2 * torch.randn(N)).cuda()
    eps = 1e-5
    assert ops.gemm_out_norm_supported(M, N, K) and not ops.gemm_out_norm_supported(M, 384, K)
    want_res = res.double() + a.double() @ w.double().t()
    want_y = want_res * torch.rsqrt(want_res.pow(2).mean(-1, keepdim=True) + eps) * nw.double()
    res_in = res.clone()
    y, r = ops.gemm_out_norm(a, w, res_in, nw, eps)
    assert torch.equal(res_in, res)                       # not in place unless asked
    assert y.dtype == torch.bfloat16 and r.dtype == torch.float32
    assert (r.double() - want_res).abs().max().item() / want_res.abs().max().item() < 2e-5
    assert (y.double() - want_y).abs().max().item() / want_y.abs().max().item() < 4e-3
    # final-norm form (no residual out) and the in-place form give the same numbers
    y2, r2 = ops.gemm_out_norm(a, w, res_in, nw, eps, want_residual=False)
    assert r2 is None and torch.equal(y2, y)
    y3, r3 = ops.gemm_out_norm(a, w, res_in, nw, eps, inplace=True)
    assert torch.equal(y3, y) and torch.equal(r3, r) and r3.data_ptr() == res_in.data_ptr()
    # against the two-launch path it replaces: same y up to the bf16 rounding of the GEMM output that path carries
    c = ops.gemm_bf16_tn(a, w)
    y_two, r_two, _, _ = ops.add_norm_fwd(c, res, nw, None, eps, True, want_residual=True)
    assert (y.float() - y_two.float()).abs().max().item() / want_y.abs().max().item() < 1e-2
    assert (r - r_two).abs().max().item() / want_res.abs().max().item() < 4e-3


def test_fastvim_tiny_fused_out_norm_matches_unfused(monkeypatch):
    """FastVim-T inference: the model with add + RMSNorm folded into the out_proj epilogue against the same model with
    separate launches (and both against each other's launch counts)."""
    from fastvim_b200 import _lib, mixer as M
    from fastvim_b200.vision import fastvim_tiny

    torch.manual_seed(0)
    model = fastvim_tiny(num_classes=10).cuda().eval()
    img = torch.randn(2, 3, 224, 224).cuda()
    outs, launches = {}, {}
    for fused in (True, False):
        monkeypatch.setattr(M, "FUSED_OUT_NORM", fused)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            model(img)
            _lib.reset_launch_count()
            outs[fused] = model(img).float()
            launches[fused] = _lib.launch_count()
    assert launches[False] - launches[True] == 24         # one add_norm launch per block folded away
    err = (outs[True] - outs[False]).abs().max().item() / outs[False].abs().max().item()
    assert err < 2e-2, err


@pytest.mark.parametrize("batch", [3, 256, 300])
def test_block_to_out_proj_dataflow_is_bit_identical(batch, monkeypatch):
    """fv_block_fwd_signal + fv_gemm_out_norm_flow (the out_proj GEMM consumes images as the block kernel publishes them,
    tiles drawn from a device counter) against the plain pair: same arithmetic per tile -> bit-identical logits.  Batch 256
    and 300 run the block kernel in two / three rounds on 148 SMs, so the two kernels really overlap."""
    from fastvim_b200 import mixer as M
    from fastvim_b200.vision import fastvim_tiny

    torch.manual_seed(0)
    model = fastvim_tiny(num_classes=10).cuda().eval()
    img = torch.randn(batch, 3, 224, 224).cuda()
    outs = {}
    for flow in (True, False, True):
        monkeypatch.setattr(M, "FLOW", flow)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            o = model(img).float()
        if flow in outs:
            assert torch.equal(o, outs[flow])           # repeatable: the dynamic tile order does not leak into the values
        outs[flow] = o
    assert torch.isfinite(outs[True]).all()
    assert torch.equal(outs[True], outs[False])


@pytest.mark.parametrize("Bt,Lp,D,ncols", [(2, 14, 384, 44), (1, 128, 384, 44), (3, 112, 768, 56), (256, 14, 1536, 80)])
def test_x_proj_batched_gemm_vs_fp64(Bt, Lp, D, ncols):
    """Both directions of x_proj in ONE launch of the general tcgen05 GEMM (fv_gemm_bf16_batched, ragged N, padded row
    pitch) against fp64 (mamba_simple_faster.py:321-323, 377-379)."""
    from fastvim_b200 import _lib, ops

    torch.manual_seed(Bt + Lp)
    u = (torch.randn(2, Bt, Lp, D) * 0.5).bfloat16().cuda()
    xw = (torch.randn(2, ncols, D) * D ** -0.5).bfloat16().cuda()
    _lib.reset_launch_count()
    xdbl = ops.x_proj(u, xw)
    assert _lib.launch_count() == 1
    want = torch.bmm(u.view(2, Bt * Lp, D).double(), xw.double().transpose(1, 2))
    assert xdbl.shape == (2, Bt * Lp, ncols)
    err = (xdbl.double() - want).abs().max().item() / want.abs().max().item()
    assert err < 4e-3, err
    assert torch.equal(xdbl, ops.x_proj(u, xw))          # deterministic


@pytest.mark.parametrize("embed,batch,xdt", [(192, 8, torch.float32), (768, 3, torch.float32), (192, 5, torch.bfloat16)])
def test_patch_embed_training_on_library_kernels(embed, batch, xdt, monkeypatch):
    """PatchEmbed under autograd (reference models/fastvim.py:67-103, an nn.Conv2d under autocast): fv_patchify + tcgen05
    GEMM forward, weight gradient on the general tcgen05 GEMM -- against an fp64 convolution of the same bf16-rounded
    operands, and against the eager unfold + F.linear path it replaces."""
    import torch.nn.functional as F

    from fastvim_b200 import _lib
    from fastvim_b200 import autograd as fv_autograd
    from fastvim_b200.vision import PatchEmbed

    torch.manual_seed(embed + batch)
    pe = PatchEmbed(img_size=224, patch_size=16, in_chans=3, embed_dim=embed).cuda()
    x = torch.randn(batch, 3, 224, 224, device="cuda").to(xdt)
    g = torch.randn(batch, 196, embed, device="cuda")

    names = []
    real = _lib.call
    monkeypatch.setattr(_lib, "call", lambda name, *a: (names.append(name), real(name, *a))[1])

    def run():
        pe.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = pe(x)
        (out.float() * g).sum().backward()
        return out.detach(), pe.proj.weight.grad.clone(), pe.proj.bias.grad.clone()

    out, dw, db = run()
    assert "fv_patchify" in names and "fv_gemm_bf16_tn" in names and "fv_gemm_bf16" in names, names
    assert out.dtype == torch.bfloat16 and dw.dtype == torch.float32 and dw.shape == pe.proj.weight.shape

    xd = x.bfloat16().double()
    wd = pe.proj.weight.detach().bfloat16().double().requires_grad_()
    bd = pe.proj.bias.detach().bfloat16().double().requires_grad_()
    want = F.conv2d(xd, wd, bd, stride=16).flatten(2).transpose(1, 2)
    assert (out.double() - want).abs().max() <= 2 ** -7 * want.abs().max()          # one bf16 rounding of the output
    (want * g.bfloat16().double()).sum().backward()                                  # the upstream gradient arrives in bf16
    assert (dw.double() - wd.grad).abs().max() <= 1e-4 * wd.grad.abs().max()
    assert (db.double() - bd.grad).abs().max() <= 1e-4 * bd.grad.abs().max()

    monkeypatch.setattr(fv_autograd, "NATIVE_PATCH_TRAIN", False)
    names.clear()
    out_e, dw_e, db_e = run()
    assert "fv_patchify" not in names
    assert (out.float() - out_e.float()).abs().max() <= 2 ** -7 * out_e.float().abs().max()
    assert (dw - dw_e).abs().max() <= 2e-2 * dw_e.abs().max()      # the eager path rounds dW through bf16 (autocast F.linear)
    assert (db - db_e).abs().max() <= 2e-2 * db_e.abs().max()


def test_patch_embed_per_channel_training_on_library_kernels(monkeypatch):
    """FastChannelVim's shared per-channel projection under autograd (reference channelvim PatchEmbedPerChannel, an
    nn.Conv3d with a (1, p, p) kernel): native path vs the eager unfold + F.linear path on the same inputs, and vs fp64."""
    from fastvim_b200 import _lib
    from fastvim_b200 import autograd as fv_autograd
    from fastvim_b200.vision_channel import PatchEmbedPerChannel

    torch.manual_seed(5)
    pe = PatchEmbedPerChannel(img_size=224, patch_size=16, stride=16, in_chans=5, embed_dim=384, hcs=False).cuda()
    x = torch.randn(2, 5, 224, 224, device="cuda")
    names = []
    real = _lib.call
    monkeypatch.setattr(_lib, "call", lambda name, *a: (names.append(name), real(name, *a))[1])

    def run():
        pe.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = pe(x)[0]
        g = torch.linspace(-1, 1, out.numel(), device="cuda").reshape(out.shape)
        (out.float() * g).sum().backward()
        return out.detach(), pe.proj.weight.grad.clone(), pe.proj.bias.grad.clone(), pe.channel_embed.weight.grad.clone(), g

    out, dw, db, dce, g = run()
    assert "fv_patchify" in names and "fv_gemm_bf16_tn" in names and "fv_gemm_bf16" in names, names
    assert dw.shape == pe.proj.weight.shape and dw.dtype == torch.float32

    # fp64 statement of the projection gradient: tokens are (b, gh, gw, c) in Channel-First order
    cols = x.bfloat16().double().reshape(2, 5, 14, 16, 14, 16).permute(0, 2, 4, 1, 3, 5).reshape(-1, 256)
    want_dw = g.bfloat16().double().reshape(-1, 384).t() @ cols
    assert (dw.double().reshape(384, 256) - want_dw).abs().max() <= 1e-4 * want_dw.abs().max()
    want_db = g.bfloat16().double().reshape(-1, 384).sum(0)
    assert (db.double() - want_db).abs().max() <= 1e-4 * want_db.abs().max()

    monkeypatch.setattr(fv_autograd, "NATIVE_PATCH_TRAIN", False)
    names.clear()
    out_e, dw_e, db_e, dce_e, _ = run()
    assert "fv_patchify" not in names
    assert (out.float() - out_e.float()).abs().max() <= 2 ** -6 * out_e.float().abs().max()
    assert (dw - dw_e).abs().max() <= 2e-2 * dw_e.abs().max()
    assert (db - db_e).abs().max() <= 2e-2 * db_e.abs().max()
    assert (dce - dce_e).abs().max() <= 2e-2 * dce_e.abs().max()
