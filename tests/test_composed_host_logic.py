"""CPU: the host logic of the operator-by-operator mixer path (fastvim_b200.composed + fastvim_b200.autograd), with the
kernel wrappers replaced by oracle stand-ins of identical signature (tests/cpu_standins.py), against vectors produced by
the reference's own modules: FastMaskVim mixer (v1 / v2), Channel-First FastChannelVim mixer, 2dcompress channelwise
layer, and max pooling under autograd (FastVim and both FastChannelVim scan orders).  Forward and every gradient."""
import pytest
import torch

import cpu_standins
from util import assert_close, load_golden


def _check(m, g, call):
    m.load_state_dict(g["params"], strict=True)
    m.train()
    h = g["hidden"].clone().requires_grad_()
    out = call(m, h)
    out.backward(g["dout"])
    assert_close(out, g["out"], 5e-5, "out")
    assert_close(h.grad, g["dhidden"], 5e-5, "dhidden")
    got = dict(m.named_parameters())
    for k, want in g["grads"].items():
        assert got[k].grad is not None, k
        assert_close(got[k].grad, want, 5e-5, "d" + k)


@pytest.mark.parametrize("name", ["mmixer_d32_4x6_keep10", "mmixer_d32_6x4_keep24_full", "mmixer_v2_d48_14x14_keep49_nonorm"])
def test_masked_mixer_host_logic(name, monkeypatch):
    from fastvim_b200.mixer_masked import Mamba_masked

    cpu_standins.install(monkeypatch)
    g = load_golden(name)
    m = Mamba_masked(g["params"]["in_proj.weight"].shape[1], token_size=list(g["token_size"]), layer_idx=0,
                     use_norm_after_ssm=g["use_norm_after_ssm"])
    _check(m, g, lambda mod, h: mod(h, g["ids_keep"]))


@pytest.mark.parametrize("name", ["cmixer_d32_4x6_t3_channel_first_grads", "cmixer_max_d32_4x6_t3_channel_first_grads",
                                  "cmixer_max_d32_6x4_t2_spatial_first_grads"])
def test_channel_mixer_host_logic(name, monkeypatch):
    from fastvim_b200.mixer_channel import Mamba

    cpu_standins.install(monkeypatch)
    g = load_golden(name)
    m = Mamba(32, token_size=list(g["token_size"]), layer_idx=0, scan_order=g["scan_order"],
              collapse_method="max" if "_max_" in name else "mean")
    _check(m, g, lambda mod, h: mod(h, g["tokens_per_patch"]))


def test_2dcompress_channelwise_layer_host_logic(monkeypatch):
    from fastvim_b200.mixer_channel_2dcompress import Mamba

    cpu_standins.install(monkeypatch)
    g = load_golden("cmixer2d_d32_4x6_t3_layer2_channels")
    m = Mamba(32, token_size=list(g["token_size"]), layer_idx=g["layer_idx"], scan_order="Channel-First")
    _check(m, g, lambda mod, h: mod(h, g["tokens_per_patch"]))


def test_fastvim_mixer_max_pool_training_host_logic(monkeypatch):
    from fastvim_b200.mixer import Mamba

    cpu_standins.install(monkeypatch)
    g = load_golden("mixer_max_d32_4x6_grads")
    m = Mamba(32, token_size=list(g["token_size"]), layer_idx=0, collapse_method="max")
    _check(m, g, lambda mod, h: mod(h))
    # geometry-folded odd layer (rotated=True: memory holds the (cols, rows) row-major grid): same result as rotating the
    # tokens physically around an un-rotated call, which is what the reference's Block does (models/fastvim.py:192-210)
    rows, cols = g["token_size"]
    h_seq = g["hidden"]
    h_mem = h_seq.reshape(2, rows, cols, -1).transpose(1, 2).reshape(2, rows * cols, -1).clone().requires_grad_()
    out_mem = m(h_mem, rotated=True)
    out_seq = out_mem.reshape(2, cols, rows, -1).transpose(1, 2).reshape(2, rows * cols, -1)
    assert_close(out_seq, g["out"], 5e-5, "rotated out")
    out_seq.backward(g["dout"])
    dh_seq = h_mem.grad.reshape(2, cols, rows, -1).transpose(1, 2).reshape(2, rows * cols, -1)
    assert_close(dh_seq, g["dhidden"], 5e-5, "rotated dhidden")


def test_pool_max_backward_matches_torch_max(monkeypatch):
    from fastvim_b200.autograd import PoolMaxBdlFn

    cpu_standins.install(monkeypatch)
    torch.manual_seed(0)
    for outer, pool, inner in ((4, 6, 1), (3, 5, 2), (1, 7, 3)):
        x = torch.randn(2, 3, outer * pool * inner)
        x[0, 0, :2 * inner] = 1.5                                     # a tie inside the first pooled group
        du = torch.randn(2, 3, outer * inner)
        xa = x.clone().requires_grad_()
        PoolMaxBdlFn.apply(xa, outer, pool, inner).backward(du)
        xb = x.clone().requires_grad_()
        xb.view(2, 3, outer, pool, inner).max(dim=3).values.reshape(2, 3, -1).backward(du)
        # identical wherever the maximum is unique; with ties the gradient goes to exactly one position
        assert torch.equal((xa.grad != 0).sum(), (xb.grad != 0).sum())
        assert torch.allclose(xa.grad.view(2, 3, outer, pool, inner).sum(3), xb.grad.view(2, 3, outer, pool, inner).sum(3))
        mask = torch.ones_like(x, dtype=torch.bool)
        mask[0, 0, :pool * inner] = False
        assert torch.equal(xa.grad[mask], xb.grad[mask])


@pytest.mark.parametrize("name", ["cscan_L128_c8_D", "cscan_L254_c2_noD", "cscan_L196_c14_D_z"])
def test_compressed_scan_autograd_host_logic(name, monkeypatch):
    """selective_scan_fn_compressed under autograd (SelectiveScanFn + BcastSkipFn + the z gate): every gradient against the
    reference's own selective_scan_ref, kernels replaced by stand-ins."""
    from fastvim_b200.interface import selective_scan_fn_compressed

    cpu_standins.install(monkeypatch)
    g = load_golden(name)
    lv = {k: (v.clone().requires_grad_() if v is not None else None) for k, v in g["inputs"].items()}
    out, st = selective_scan_fn_compressed(lv["u"], lv["u_compressed"], lv["delta"], lv["A"], lv["B"], lv["C"], lv["D"],
                                           z=lv["z"], delta_bias=lv["delta_bias"], delta_softplus=True, return_last_state=True)
    assert not st.requires_grad
    out.backward(g["dout"])
    assert_close(out, g["out"], 2e-5, "out")
    assert_close(st, g["last_state"], 2e-5, "last_state")
    for k, want in g["grads"].items():
        assert lv[k].grad is not None and lv[k].grad.shape == want.shape, k
        assert_close(lv[k].grad, want, 2e-5, "d" + k)


@pytest.mark.parametrize("has_z", [True, False])
def test_mamba_inner_fn_autograd_host_logic(has_z, monkeypatch):
    """mamba_inner_fn_no_out_proj[_withoutZ] and FastVim_mamba_inner_fn_no_out_proj_withoutZ under autograd against fp64
    autograd through the oracle's restatement of the reference functions."""
    import fastvim_oracle as O
    from fastvim_b200 import interface as I

    cpu_standins.install(monkeypatch)
    torch.manual_seed(0)
    Bt, Dm, rows, cols, N, R = 2, 16, 3, 5, 8, 3
    L = rows * cols
    xz = torch.randn(Bt, 2 * Dm if has_z else Dm, L)
    cw, cb = torch.randn(Dm, 1, 4) * 0.5, torch.randn(Dm) * 0.5
    xw, dw = torch.randn(R + 2 * N, Dm) * Dm ** -0.5, torch.randn(Dm, R) * R ** -0.5
    A, Dp, dbias = -0.5 * torch.rand(Dm, N) - 0.05, torch.randn(Dm), 0.5 * torch.rand(Dm) - 2.0
    dout = torch.randn(Bt, Dm, L)
    tensors = [xz, cw, cb, xw, dw, A, Dp, dbias]

    def grads(fn, conv):
        lv = [conv(t).requires_grad_() for t in tensors]
        out = fn(*lv)
        out.backward(conv(dout))
        return out.detach(), [v.grad for v in lv]

    ours = I.mamba_inner_fn_no_out_proj if has_z else I.mamba_inner_fn_no_out_proj_withoutZ
    pairs = [(lambda a, b, c, d, e, f, g_, h: ours(a, b, c, d, e, f, None, None, g_, h, delta_softplus=True),
              lambda a, b, c, d, e, f, g_, h: O.mamba_inner_oracle(a, b, c, d, e, f, None, None, g_, h, True, has_z=has_z))]
    if not has_z:
        pairs.append((lambda a, b, c, d, e, f, g_, h: I.FastVim_mamba_inner_fn_no_out_proj_withoutZ(
                          a, b, c, d, e, f, None, None, g_, h, None, None, True, cols, "mean", 0.5, (-1, Dm, rows, cols)),
                      lambda a, b, c, d, e, f, g_, h: O.fastvim_inner_oracle(a, b, c, d, e, f, g_, h, cols, 0.5)))
    for f_ours, f_oracle in pairs:
        out_g, gg = grads(f_ours, lambda t: t.clone())
        out_w, gw = grads(f_oracle, lambda t: t.double().clone())
        assert_close(out_g, out_w, 2e-5, "out")
        for i, (a, b) in enumerate(zip(gg, gw)):
            assert a is not None and a.shape == b.shape, i
            assert_close(a, b, 5e-5, f"grad {i}")


@pytest.mark.parametrize("per_channel", [False, True])
def test_patch_embed_fn_host_logic(per_channel, monkeypatch):
    """autograd.PatchEmbedFn with the two kernel wrappers replaced by torch stand-ins of the same contract (bf16 patches;
    fp32-accumulated GEMM with the fp32 bias added before the bf16 rounding): output, weight and bias gradients against
    autograd through the reference's own statement, an nn.Conv2d / shared-filter Conv3d over bf16-rounded operands
    (models/fastvim.py:67-103, channelvim PatchEmbedPerChannel)."""
    import torch.nn.functional as F

    from fastvim_b200 import autograd as A
    from fastvim_b200 import ops

    def patchify(img, patch, per_channel=False):
        B, C, H, W = img.shape
        gh, gw = H // patch, W // patch
        t = img.reshape(B, C, gh, patch, gw, patch)
        t = t.permute(0, 1, 2, 4, 3, 5).reshape(-1, patch * patch) if per_channel else \
            t.permute(0, 2, 4, 1, 3, 5).reshape(B * gh * gw, C * patch * patch)
        return t.to(torch.bfloat16)

    def gemm_bf16_tn(a, w, out=None, bias=None):
        assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and (bias is None or bias.dtype == torch.float32)
        acc = a.float() @ w.float().t()
        return (acc if bias is None else acc + bias).to(torch.bfloat16)

    monkeypatch.setattr(ops, "patchify", patchify)
    monkeypatch.setattr(ops, "gemm_bf16_tn", gemm_bf16_tn)
    torch.manual_seed(3)
    B, C, p, E = 2, 3, 4, 16
    x = torch.randn(B, C, 8, 12)
    w = (torch.randn(E, 1 if per_channel else C, p, p) * 0.2).requires_grad_()
    b = torch.randn(E).requires_grad_()
    out = A.PatchEmbedFn.apply(x, w, b, p, per_channel)
    ntok = B * (C if per_channel else 1) * 2 * 3
    assert out.shape == (ntok, E) and out.dtype == torch.bfloat16
    g = torch.randn(ntok, E).to(torch.bfloat16)
    out.backward(g)
    assert w.grad.shape == w.shape and w.grad.dtype == torch.float32 and b.grad.dtype == torch.float32

    xr = x.to(torch.bfloat16).double()
    wr = w.detach().to(torch.bfloat16).double().requires_grad_()
    br = b.detach().to(torch.bfloat16).double().requires_grad_()
    if per_channel:   # the same filter on every channel; tokens in (b, c, gh, gw) order
        want = F.conv2d(xr.reshape(B * C, 1, 8, 12), wr, br, stride=p).flatten(2).transpose(1, 2).reshape(ntok, E)
    else:
        want = F.conv2d(xr, wr, br, stride=p).flatten(2).transpose(1, 2).reshape(ntok, E)
    assert (out.double() - want).abs().max() <= 2 ** -8 * want.abs().max() + 1e-6
    want.backward(g.double())
    # on CPU tensors _wgrad falls back to a bf16 matmul whose RESULT is rounded to bf16 (the tcgen05 / cuBLAS forms write the
    # fp32 accumulator; tests/test_gpu_gemm_general.py holds them to 1e-4)
    assert (w.grad.double() - wr.grad).abs().max() <= 2 ** -7 * wr.grad.abs().max()
    assert_close(b.grad, br.grad.float(), 1e-5, "db")
    # no bias; frozen projection
    w2 = w.detach().clone().requires_grad_()
    A.PatchEmbedFn.apply(x, w2, None, p, per_channel).backward(g)
    assert torch.equal(w2.grad, w.grad)
