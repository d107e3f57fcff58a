"""TEST INFRASTRUCTURE ONLY -- pins the oracle to the reference and writes tests/golden.

Run in the build container (needs /root/reference; never on the GPU box):

    python oracle/gen_golden.py

It imports the UNMODIFIED reference Python (``oracle/ref_loader.py``), runs it on CPU
on seeded inputs, asserts that ``oracle/fastvim_oracle.py`` reproduces it, and commits
small input/output/gradient vectors as ``tests/golden/*.pt`` plus a ``manifest.json``
recording the oracle-vs-reference error of every case (including the full FastVim-T
224x224 model, whose 7 M-parameter state dict is too large to commit).

Generators mirror the reference's own tests:
``mamba-1p1p1/tests/ops/test_selective_scan.py:61-122`` (seed 0, batch 2, dim 4,
dstate 8, A=-0.5*rand, delta=0.5*rand, delta_bias=0.5*rand, B,C,u,z=randn).
"""
from __future__ import annotations

import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fastvim_oracle as O  # noqa: E402
from ref_loader import build_reference_fastvim, load_reference  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(GOLD, exist_ok=True)
manifest = {}


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def save(name, obj, **meta):
    path = os.path.join(GOLD, name + ".pt")
    torch.save(obj, path)
    manifest[name] = dict(meta, bytes=os.path.getsize(path))
    print(f"  {name}: {meta}  ({os.path.getsize(path)} bytes)")


def scan_case(ref, seqlen, groups, has_D, has_z, has_bias, softplus, varB=True, varC=True):
    torch.random.manual_seed(0)
    bs, dim, ds = 2, 4, 8
    A = (-0.5 * torch.rand(dim, ds)).requires_grad_()
    shp = (bs, ds, seqlen) if groups == 1 else (bs, groups, ds, seqlen)
    B = torch.randn(*(shp if varB else (dim, ds)), requires_grad=True)
    C = torch.randn(*(shp if varC else (dim, ds)), requires_grad=True)
    D = torch.randn(dim, requires_grad=True) if has_D else None
    z = torch.randn(bs, dim, seqlen, requires_grad=True) if has_z else None
    db = (0.5 * torch.rand(dim)).requires_grad_() if has_bias else None
    u = torch.randn(bs, dim, seqlen, requires_grad=True)
    delta = (0.5 * torch.rand(bs, dim, seqlen)).requires_grad_()
    ins = dict(u=u, delta=delta, A=A, B=B, C=C, D=D, z=z, delta_bias=db)

    def run(fn):
        leaves = {k: (v.detach().clone().requires_grad_() if v is not None else None) for k, v in ins.items()}
        out, st = fn(leaves["u"], leaves["delta"], leaves["A"], leaves["B"], leaves["C"], leaves["D"],
                     z=leaves["z"], delta_bias=leaves["delta_bias"], delta_softplus=softplus,
                     return_last_state=True)
        torch.manual_seed(1)
        g = torch.randn_like(out)
        out.backward(g)
        return out.detach(), st.detach(), {k: v.grad for k, v in leaves.items() if v is not None}, g

    out_r, st_r, gr_r, g = run(ref.ssi.selective_scan_ref)
    out_o, st_o, gr_o, _ = run(O.selective_scan_oracle)
    errs = {"out": relerr(out_o, out_r), "state": relerr(st_o, st_r)}
    errs.update({"d" + k: relerr(gr_o[k], gr_r[k]) for k in gr_r})
    assert max(errs.values()) < 2e-5, errs
    return dict(inputs={k: (v.detach() if v is not None else None) for k, v in ins.items()},
                delta_softplus=softplus, dout=g, out=out_r, last_state=st_r, grads=gr_r), errs


def main():
    ref = load_reference()
    torch.set_num_threads(os.cpu_count() or 1)

    print("[1] selective_scan_ref  (selective_scan_interface.py:126-206)")
    for seqlen in (14, 128, 300):
        for groups in (1, 2):
            case, errs = scan_case(ref, seqlen, groups, True, True, True, True)
            save(f"scan_L{seqlen}_g{groups}_full", case, oracle_vs_ref=max(errs.values()))
    case, errs = scan_case(ref, 64, 1, False, False, False, False)
    save("scan_L64_plain", case, oracle_vs_ref=max(errs.values()))
    case, errs = scan_case(ref, 64, 1, True, False, True, True, varB=False, varC=False)
    save("scan_L64_constBC", case, oracle_vs_ref=max(errs.values()))

    print("[2] causal conv fallback form (mamba_simple.py:302-303) -- oracle == shim by construction")
    torch.manual_seed(0)
    x = torch.randn(2, 6, 23)
    w, b = torch.randn(6, 4), torch.randn(6)
    import causal_conv1d
    y_r = causal_conv1d.causal_conv1d_fn(x, w, b, activation="silu")
    assert relerr(O.causal_conv1d_oracle(x, w, b), y_r) < 1e-6
    # independent scalar restatement of the formula, to pin the tap order
    y_s = torch.zeros_like(x)
    for t in range(23):
        acc = b.clone()[None].repeat(2, 1)
        for k in range(4):
            if t - 3 + k >= 0:
                acc = acc + w[:, k] * x[:, :, t - 3 + k]
        y_s[:, :, t] = acc * torch.sigmoid(acc)
    assert relerr(y_s, y_r) < 1e-6
    save("conv_W4", dict(x=x, w=w, b=b, out=y_r), oracle_vs_ref=relerr(O.causal_conv1d_oracle(x, w, b), y_r))

    print("[3] FastVim Mamba mixer live branch (mamba_simple_faster.py:181-457)")
    for name, d_model, ts, norm, sf in (("mixer_d32_4x6", 32, (4, 6), True, 1.0),
                                        ("mixer_d32_6x4_nonorm_sf", 32, (6, 4), False, 0.25),
                                        ("mixer_d48_14x14", 48, (14, 14), True, 1.0)):
        torch.manual_seed(0)
        m = ref.msf.Mamba(d_model, token_size=list(ts), layer_idx=0, use_norm_after_ssm=norm,
                          scaling_factor=sf)
        with torch.no_grad():  # de-trivialise the init so every parameter matters
            for k, v in m.named_parameters():
                if k in ("D", "D_b", "layernorm.weight"):
                    v.add_(0.1 * torch.randn_like(v))
                elif k in ("A_log", "A_b_log", "layernorm.bias"):
                    v.add_(0.1 * torch.randn_like(v))
        h = torch.randn(2, ts[0] * ts[1], d_model, requires_grad=True)
        out = m(h)
        torch.manual_seed(1)
        g = torch.randn_like(out)
        out.backward(g)
        params = {k: v.detach().clone() for k, v in m.named_parameters()}
        grads = {k: v.grad.detach().clone() for k, v in m.named_parameters()}
        p2 = {k: v.clone().requires_grad_() for k, v in params.items()}
        h2 = h.detach().clone().requires_grad_()
        out_o = O.mixer_oracle(h2, p2, ts, use_norm_after_ssm=norm, scaling_factor=sf)
        out_o.backward(g)
        errs = {"out": relerr(out_o, out), "dh": relerr(h2.grad, h.grad)}
        errs.update({"d" + k: relerr(p2[k].grad, grads[k]) for k in grads})
        assert max(errs.values()) < 5e-5, errs
        save(name, dict(params=params, hidden=h.detach(), token_size=ts, use_norm_after_ssm=norm,
                        scaling_factor=sf, dout=g, out=out.detach(), dhidden=h.grad, grads=grads),
             oracle_vs_ref=max(errs.values()))

    print("[3b] FastChannelVim mixer (mamba_simple_channel_faster.py:176-420), both scan orders")
    for name, d_model, ts, tpp, order in (("cmixer_d32_4x6_t3_channel_first", 32, (4, 6), 3, "Channel-First"),
                                          ("cmixer_d32_6x4_t2_spatial_first", 32, (6, 4), 2, "Spatial-First")):
        torch.manual_seed(0)
        m = ref.mscf.Mamba(d_model, token_size=list(ts), layer_idx=0, scan_order=order)
        with torch.no_grad():
            for k, v in m.named_parameters():
                if k in ("D", "D_b", "layernorm.weight", "A_log", "A_b_log", "layernorm.bias"):
                    v.add_(0.1 * torch.randn_like(v))
        h = torch.randn(2, ts[0] * ts[1] * tpp, d_model)
        with torch.no_grad():
            out = m(h, tpp)
        params = {k: v.detach().clone() for k, v in m.named_parameters()}
        layout = (ts[0], ts[1], tpp) if order == "Channel-First" else (tpp * ts[0], ts[1], 1)
        out_o = O.mixer_oracle(h, params, ts, layout=layout)
        e = relerr(out_o, out)
        assert e < 5e-5, (name, e)
        save(name, dict(params=params, hidden=h, token_size=ts, tokens_per_patch=tpp, scan_order=order, out=out),
             oracle_vs_ref=e)

    print("[4] mamba_inner_ref  (selective_scan_interface.py:1757-1810, without out_proj)")
    torch.manual_seed(0)
    Bt, Dm, L, N, R = 2, 16, 40, 8, 3
    xz = torch.randn(Bt, 2 * Dm, L)
    cw, cb = torch.randn(Dm, 1, 4) * 0.5, torch.randn(Dm)
    xw, dw = torch.randn(R + 2 * N, Dm) * 0.3, torch.randn(Dm, R) * 0.3
    A, Dp, dbias = -0.5 * torch.rand(Dm, N), torch.randn(Dm), 0.5 * torch.rand(Dm)
    eye = torch.eye(Dm)
    out_r = ref.ssi.mamba_inner_ref(xz, cw, cb, xw, dw, eye, None, A, None, None, Dp, dbias,
                                    delta_softplus=True)  # out_proj = identity
    out_o = O.mamba_inner_oracle(xz, cw, cb, xw, dw, A, None, None, Dp, dbias, True).transpose(1, 2)
    e = relerr(out_o, out_r)
    assert e < 2e-5, e
    save("mamba_inner", dict(xz=xz, conv_w=cw, conv_b=cb, x_proj_w=xw, dt_proj_w=dw, A=A, D=Dp,
                             delta_bias=dbias, out=out_r.transpose(1, 2).contiguous()), oracle_vs_ref=e)

    print("[5] fused add + RMSNorm refs (ops/triton/layernorm.py:18-49)")
    torch.manual_seed(0)
    x, res, w = torch.randn(3, 10, 24), torch.randn(3, 10, 24), torch.rand(24) + 0.5
    y_r, r_r = ref.ln.rms_norm_ref(x, w, None, residual=res, eps=1e-5, prenorm=True, upcast=True)
    y_o, r_o = O.add_norm_oracle(x, w, None, res, 1e-5, True)
    e = max(relerr(y_o, y_r), relerr(r_o, r_r))
    assert e < 1e-6
    save("rmsnorm", dict(x=x, residual=res, weight=w, eps=1e-5, y=y_r, residual_out=r_r), oracle_vs_ref=e)

    print("[6] small VisionMamba end to end (models/fastvim.py:484-557), 64x96 image -> 4x6 tokens")
    torch.manual_seed(0)
    model = build_reference_fastvim(ref, embed_dim=32, depth=4, img_size=(64, 96), num_classes=10)
    with torch.no_grad():
        for k, v in model.named_parameters():
            if k.endswith(("mixer.D", "mixer.D_b", "norm.weight", "layernorm.weight", "layernorm.bias",
                           "A_log", "A_b_log", "head.bias", "norm_f.weight")):
                v.add_(0.1 * torch.randn_like(v))
    imgs = torch.randn(2, 3, 64, 96)
    logits = model(imgs)
    torch.manual_seed(1)
    g = torch.randn_like(logits)
    logits.backward(g)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    gr = {k: v.grad.detach().clone() for k, v in model.named_parameters()}
    sd2 = {k: v.clone().requires_grad_() for k, v in sd.items()}
    lo = O.fastvim_oracle(imgs, sd2, depth=4)
    lo.backward(g)
    errs = {"logits": relerr(lo, logits)}
    errs.update({"d" + k: relerr(sd2[k].grad, gr[k]) for k in gr})
    worst = max(errs, key=errs.get)
    assert errs[worst] < 2e-4, (worst, errs[worst])
    save("fastvim_small", dict(state_dict=sd, images=imgs, depth=4, dlogits=g, logits=logits.detach(), grads=gr),
         oracle_vs_ref=errs[worst], worst=worst)

    print("[7] full FastVim-T 224x224 batch 2 (BASELINE.json configs[0]); state dict not committed")
    model = build_reference_fastvim(ref, embed_dim=192, depth=24, img_size=224)
    torch.manual_seed(0)
    imgs = torch.randn(2, 3, 224, 224)
    with torch.no_grad():
        logits = model(imgs)
        lo = O.fastvim_oracle(imgs, model.state_dict(), depth=24)
    e = relerr(lo, logits)
    assert e < 1e-4, e
    manifest["fastvim_tiny_224_full"] = dict(oracle_vs_ref=e, note="checked at generation time only")
    print("  fastvim_tiny_224_full:", e)

    print("[8] rectangular / rotated full-size sanity: FastVim-T 2048^2-shape scaled to 256x512")
    model = build_reference_fastvim(ref, embed_dim=64, depth=3, img_size=(256, 512), num_classes=5)
    imgs = torch.randn(1, 3, 256, 512)
    with torch.no_grad():
        logits = model(imgs)
        lo = O.fastvim_oracle(imgs, model.state_dict(), depth=3)
    e = relerr(lo, logits)
    assert e < 1e-4, e
    manifest["fastvim_rect_16x32"] = dict(oracle_vs_ref=e, note="checked at generation time only")

    with open(os.path.join(GOLD, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("golden vectors written to", GOLD)


if __name__ == "__main__":
    main()
