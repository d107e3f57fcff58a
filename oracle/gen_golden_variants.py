"""TEST INFRASTRUCTURE ONLY -- golden vectors for the FastMaskVim and Channel-First FastChannelVim mixers.

Run in the build container (needs /root/reference):  python oracle/gen_golden_variants.py

Imports the UNMODIFIED reference modules ``mamba_simple_masked_faster{,_v2}.Mamba_masked`` and
``mamba_simple_channel_faster.Mamba`` (through ``oracle/ref_loader.py``), runs them forward AND backward on CPU on
seeded inputs, asserts that ``oracle/fastvim_oracle.py`` reproduces outputs and every gradient, and writes
``tests/golden/mmixer_*.pt`` / ``cmixer_*_grads.pt`` (kept separate from gen_golden.py so the earlier vectors are
not rewritten).
"""
from __future__ import annotations

import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fastvim_oracle as O  # noqa: E402
from ref_loader import load_reference  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def detrivialise(m):
    with torch.no_grad():
        for k, v in m.named_parameters():
            if k in ("D", "D_b", "layernorm.weight", "A_log", "A_b_log", "layernorm.bias"):
                v.add_(0.1 * torch.randn_like(v))


def run_case(module, oracle_fn, h, extra):
    h = h.clone().requires_grad_()
    out = module(h, extra)
    torch.manual_seed(1)
    g = torch.randn_like(out)
    out.backward(g)
    params = {k: v.detach().clone() for k, v in module.named_parameters()}
    grads = {k: v.grad.detach().clone() for k, v in module.named_parameters()}
    p2 = {k: v.clone().requires_grad_() for k, v in params.items()}
    h2 = h.detach().clone().requires_grad_()
    out_o = oracle_fn(h2, p2)
    out_o.backward(g)
    errs = {"out": relerr(out_o, out), "dh": relerr(h2.grad, h.grad)}
    errs.update({"d" + k: relerr(p2[k].grad, grads[k]) for k in grads})
    worst = max(errs, key=errs.get)
    assert errs[worst] < 5e-5, (worst, errs[worst])
    return dict(params=params, hidden=h.detach(), dout=g, out=out.detach(), dhidden=h.grad.detach(), grads=grads), errs[worst]


def main():
    ref = load_reference()
    manifest_path = os.path.join(GOLD, "manifest.json")
    manifest = json.load(open(manifest_path))

    print("[m] FastMaskVim mixer (mamba_simple_masked_faster.py:167-325 and _v2)")
    for name, mod, d_model, ts, keep, norm in (("mmixer_d32_4x6_keep10", ref.msmf, 32, (4, 6), 10, True),
                                               ("mmixer_d32_6x4_keep24_full", ref.msmf, 32, (6, 4), 24, True),
                                               ("mmixer_v2_d48_14x14_keep49_nonorm", ref.msmf2, 48, (14, 14), 49, False)):
        torch.manual_seed(0)
        m = mod.Mamba_masked(d_model, token_size=list(ts), layer_idx=0, use_norm_after_ssm=norm)
        detrivialise(m)
        Bt, Ltot = 2, ts[0] * ts[1]
        # kept ids sorted ascending, as random_masking produces them (models/mae/..._v2.py:740-774)
        ids = torch.stack([torch.randperm(Ltot)[:keep].sort().values for _ in range(Bt)])
        h = torch.randn(Bt, keep, d_model)
        case, e = run_case(m, lambda hh, pp: O.mixer_oracle(hh, pp, ts, use_norm_after_ssm=norm, ids_keep=ids), h, ids)
        case.update(token_size=ts, ids_keep=ids, use_norm_after_ssm=norm)
        path = os.path.join(GOLD, name + ".pt")
        torch.save(case, path)
        manifest[name] = dict(oracle_vs_ref=e, bytes=os.path.getsize(path))
        print(f"  {name}: oracle_vs_ref {e:.2e}")

    print("[c] FastChannelVim mixer with gradients (mamba_simple_channel_faster.py:176-420)")
    for name, d_model, ts, tpp, order in (("cmixer_d32_4x6_t3_channel_first_grads", 32, (4, 6), 3, "Channel-First"),
                                          ("cmixer_d32_6x4_t2_spatial_first_grads", 32, (6, 4), 2, "Spatial-First")):
        torch.manual_seed(0)
        m = ref.mscf.Mamba(d_model, token_size=list(ts), layer_idx=0, scan_order=order)
        detrivialise(m)
        h = torch.randn(2, ts[0] * ts[1] * tpp, d_model)
        layout = (ts[0], ts[1], tpp) if order == "Channel-First" else (tpp * ts[0], ts[1], 1)
        case, e = run_case(m, lambda hh, pp: O.mixer_oracle(hh, pp, ts, layout=layout), h, tpp)
        case.update(token_size=ts, tokens_per_patch=tpp, scan_order=order)
        path = os.path.join(GOLD, name + ".pt")
        torch.save(case, path)
        manifest[name] = dict(oracle_vs_ref=e, bytes=os.path.getsize(path))
        print(f"  {name}: oracle_vs_ref {e:.2e}")

    with open(manifest_path, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__" and "--channel-model" not in sys.argv and "--2dcompress" not in sys.argv and "--masked-blocks" not in sys.argv and "--compressed" not in sys.argv and "--maxpool" not in sys.argv:
    main()


def gen_channel_model():
    """FastChannelVim model wrapper (models/channel_wise_tokenization/models_channel_mamba_faster.py:458-683):
    the reference's own VisionMamba on CPU, eval mode (HCS off), 3-channel 32x64 images -> 2 x 4 patches x 3 channels
    = 24 tokens, 4 blocks (two of them with the odd-layer token transposition), both scan orders; logits and the
    gradient of every parameter."""
    import importlib

    load_reference()
    cm = importlib.import_module("models.channel_wise_tokenization.models_channel_mamba_faster")
    manifest_path = os.path.join(GOLD, "manifest.json")
    manifest = json.load(open(manifest_path))
    for name, order in (("channelvim_small_cf", "Channel-First"), ("channelvim_small_sf", "Spatial-First")):
        torch.manual_seed(0)
        model = cm.VisionMamba(img_size=(32, 64), patch_size=16, stride=16, depth=4, embed_dim=32, channels=3,
                               num_classes=7, rms_norm=True, residual_in_fp32=True, fused_add_norm=False,
                               final_pool_type="mean", if_abs_pos_embed=True, drop_path_rate=0.0, scan_order=order,
                               hcs=False).eval()
        with torch.no_grad():
            for k, v in model.named_parameters():
                if k.endswith(("mixer.D", "mixer.D_b", "norm.weight", "layernorm.weight", "layernorm.bias", "A_log",
                               "A_b_log", "head.bias", "norm_f.weight")):
                    v.add_(0.1 * torch.randn_like(v))
        imgs = torch.randn(2, 3, 32, 64)
        logits = model(imgs)
        torch.manual_seed(1)
        g = torch.randn_like(logits)
        logits.backward(g)
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        gr = {k: v.grad.detach().clone() for k, v in model.named_parameters()}
        path = os.path.join(GOLD, name + ".pt")
        torch.save(dict(state_dict=sd, images=imgs, scan_order=order, dlogits=g, logits=logits.detach(), grads=gr,
                        kwargs=dict(img_size=(32, 64), depth=4, embed_dim=32, channels=3, num_classes=7)), path)
        manifest[name] = dict(bytes=os.path.getsize(path), note="reference model output (no oracle restatement: the "
                              "CUDA model is compared with these vectors directly)")
        print(f"  {name}: logits {tuple(logits.shape)}, {len(gr)} gradients, {os.path.getsize(path)} bytes")
    with open(manifest_path, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__" and "--channel-model" in sys.argv:
    gen_channel_model()


def gen_2dcompress():
    """FastChannelVim 2dcompress mixer (mamba_simple_channel_faster_2dcompress.py:176-425): both layer kinds, fwd + grads."""
    import importlib

    load_reference()
    mod = importlib.import_module("mamba_ssm.modules.mamba_simple_channel_faster_2dcompress")
    manifest_path = os.path.join(GOLD, "manifest.json")
    manifest = json.load(open(manifest_path))
    for name, layer_idx in (("cmixer2d_d32_4x6_t3_layer0_rows", 0), ("cmixer2d_d32_4x6_t3_layer2_channels", 2)):
        torch.manual_seed(0)
        ts, tpp, d_model = (4, 6), 3, 32
        m = mod.Mamba(d_model, token_size=list(ts), layer_idx=layer_idx, scan_order="Channel-First")
        detrivialise(m)
        h = torch.randn(2, ts[0] * ts[1] * tpp, d_model)
        layout = (1, ts[0] * ts[1], tpp) if (layer_idx + 1) % 3 == 0 else (ts[0], ts[1] * tpp, 1)
        case, e = run_case(m, lambda hh, pp: O.mixer_oracle(hh, pp, ts, layout=layout), h, tpp)
        case.update(token_size=ts, tokens_per_patch=tpp, layer_idx=layer_idx, layout=layout)
        path = os.path.join(GOLD, name + ".pt")
        torch.save(case, path)
        manifest[name] = dict(oracle_vs_ref=e, bytes=os.path.getsize(path))
        print(f"  {name}: oracle_vs_ref {e:.2e}")
    with open(manifest_path, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__" and "--2dcompress" in sys.argv:
    gen_2dcompress()


def gen_masked_blocks():
    """FastMaskVim encoder blocks (models/mae/models_mamba_faster_mae_vimdecoder_v2.py:279-466): a stack of three
    reference ``Block_masked`` (layer 1 takes the odd-layer id rotation) + the final add + RMSNorm, on 10 kept tokens of
    a 4 x 6 grid; outputs and every gradient."""
    import importlib

    ref = load_reference()
    mm = importlib.import_module("models.mae.models_mamba_faster_mae_vimdecoder_v2")
    manifest_path = os.path.join(GOLD, "manifest.json")
    manifest = json.load(open(manifest_path))
    torch.manual_seed(0)
    ts, d_model, keep, Bt, depth = (4, 6), 32, 10, 2, 3
    layers = torch.nn.ModuleList([mm.create_block_masked(d_model, rms_norm=True, residual_in_fp32=True, fused_add_norm=False,
                                                         layer_idx=i, token_size=ts) for i in range(depth)])
    norm_f = ref.ln.RMSNorm(d_model, eps=1e-5)
    with torch.no_grad():
        for k, v in list(layers.named_parameters()) + list(norm_f.named_parameters()):
            if k.endswith(("mixer.D", "mixer.D_b", "norm.weight", "layernorm.weight", "layernorm.bias", "A_log", "A_b_log",
                           "weight")) and v.dim() == 1 or k.endswith(("A_log", "A_b_log")):
                v.add_(0.1 * torch.randn_like(v))
    ids = torch.stack([torch.randperm(ts[0] * ts[1])[:keep].sort().values for _ in range(Bt)])
    h = torch.randn(Bt, keep, d_model, requires_grad=True)
    hidden, residual = h, None
    for layer in layers:
        hidden, residual = layer(hidden, residual, ids.clone())
    out = ref.ln.rms_norm_ref(hidden, norm_f.weight, None, residual=residual, eps=1e-5, prenorm=False, upcast=True)
    torch.manual_seed(1)
    g = torch.randn_like(out)
    out.backward(g)
    sd = {"layers." + k: v.detach().clone() for k, v in layers.state_dict().items()}
    sd["norm_f.weight"] = norm_f.weight.detach().clone()
    grads = {"layers." + k: v.grad.detach().clone() for k, v in layers.named_parameters()}
    grads["norm_f.weight"] = norm_f.weight.grad.detach().clone()
    path = os.path.join(GOLD, "mblocks_d32_4x6_keep10.pt")
    torch.save(dict(state_dict=sd, hidden=h.detach(), ids_keep=ids, token_size=ts, depth=depth, dout=g, out=out.detach(),
                    dhidden=h.grad.detach(), grads=grads), path)
    manifest["mblocks_d32_4x6_keep10"] = dict(bytes=os.path.getsize(path), note="reference Block_masked stack output")
    print("  mblocks_d32_4x6_keep10:", tuple(out.shape), len(grads), "gradients")
    with open(manifest_path, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__" and "--masked-blocks" in sys.argv:
    gen_masked_blocks()


def gen_compressed_scan():
    """The 6-tensor compressed scan of the reference's own kernel package (fastvim_kernel/mamba-1p1p1/faster_mamba_ssm/
    ops/selective_scan_interface.py:162-252 selective_scan_ref), generator as in its test
    (fastvim_kernel/mamba-1p1p1/tests/test_compressed_scan.py: seed 0, dstate 8, A = -0.5 rand, delta = 0.5 rand,
    delta_bias = 0.5 rand), forward + autograd gradients on CPU."""
    import contextlib
    import importlib
    import io
    import types

    sys.modules.setdefault("faster_selective_scan_cuda", types.ModuleType("faster_selective_scan_cuda"))
    kroot = "/root/reference/fastvim_kernel/mamba-1p1p1"
    if kroot not in sys.path:
        sys.path.insert(0, kroot)
    fssi = importlib.import_module("faster_mamba_ssm.ops.selective_scan_interface")
    manifest_path = os.path.join(GOLD, "manifest.json")
    manifest = json.load(open(manifest_path))
    for name, (bs, dim, L, cfac, has_D, has_z) in (("cscan_L128_c8_D", (2, 4, 128, 8, True, False)),
                                                   ("cscan_L254_c2_noD", (1, 1, 254, 2, False, False)),
                                                   ("cscan_L196_c14_D_z", (2, 6, 196, 14, True, True))):
        torch.random.manual_seed(0)
        ds, Lc = 8, L // cfac
        ins = dict(A=-0.5 * torch.rand(dim, ds), B=torch.randn(bs, ds, Lc), C=torch.randn(bs, ds, Lc),
                   D=torch.randn(dim) if has_D else None, z=torch.randn(bs, dim, L) if has_z else None,
                   delta_bias=0.5 * torch.rand(dim), u=torch.randn(bs, dim, L), delta=0.5 * torch.rand(bs, dim, Lc))
        ins["u_compressed"] = ins["u"].reshape(bs, dim, Lc, cfac).mean(dim=3)        # pooled over consecutive positions

        def run(fn):
            lv = {k: (v.detach().clone().requires_grad_() if v is not None else None) for k, v in ins.items()}
            with contextlib.redirect_stdout(io.StringIO()):                          # the reference prints shapes
                out, st = fn(lv["u"], lv["u_compressed"], lv["delta"], lv["A"], lv["B"], lv["C"], lv["D"], z=lv["z"],
                             delta_bias=lv["delta_bias"], delta_softplus=True, return_last_state=True)
            torch.manual_seed(1)
            g = torch.randn_like(out)
            out.backward(g)
            return out.detach(), st.detach(), {k: v.grad for k, v in lv.items() if v is not None and v.grad is not None}, g

        out_r, st_r, gr_r, g = run(fssi.selective_scan_ref)
        out_o, st_o, gr_o, _ = run(O.compressed_scan_oracle)
        errs = {"out": relerr(out_o, out_r), "state": relerr(st_o, st_r)}
        errs.update({"d" + k: relerr(gr_o[k], gr_r[k]) for k in gr_r})
        assert max(errs.values()) < 2e-5, errs
        path = os.path.join(GOLD, name + ".pt")
        torch.save(dict(inputs={k: v for k, v in ins.items()}, cfac=cfac, dout=g, out=out_r, last_state=st_r, grads=gr_r), path)
        manifest[name] = dict(oracle_vs_ref=max(errs.values()), bytes=os.path.getsize(path))
        print(f"  {name}: oracle_vs_ref {max(errs.values()):.2e}")
    with open(manifest_path, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__" and "--compressed" in sys.argv:
    gen_compressed_scan()


def gen_maxpool():
    """collapse_method="max" under autograd (the live module branch differentiates x.reshape(...).max(dim).values:
    mamba_simple_faster.py:299-305, mamba_simple_channel_faster.py:258-289; shipped config
    cell_imaging/config/FastChannelVimS_maxpool.yaml): forward + every gradient from the reference's own modules."""
    ref = load_reference()
    manifest_path = os.path.join(GOLD, "manifest.json")
    manifest = json.load(open(manifest_path))
    cases = (("cmixer_max_d32_4x6_t3_channel_first_grads", "channel", (4, 6), 3, "Channel-First"),
             ("cmixer_max_d32_6x4_t2_spatial_first_grads", "channel", (6, 4), 2, "Spatial-First"),
             ("mixer_max_d32_4x6_grads", "plain", (4, 6), 1, None))
    for name, kind, ts, tpp, order in cases:
        torch.manual_seed(0)
        if kind == "channel":
            m = ref.mscf.Mamba(32, token_size=list(ts), layer_idx=0, scan_order=order, collapse_method="max")
            layout = (ts[0], ts[1], tpp) if order == "Channel-First" else (tpp * ts[0], ts[1], 1)
            extra = tpp
        else:
            m = ref.msf.Mamba(32, token_size=list(ts), layer_idx=0, collapse_method="max")
            layout, extra = (ts[0], ts[1], 1), None
        detrivialise(m)
        h = torch.randn(2, ts[0] * ts[1] * tpp, 32)
        case, e = run_case(_Wrap(m, kind == "channel"), lambda hh, pp: O.mixer_oracle(hh, pp, ts, layout=layout, collapse_method="max"),
                           h, extra)
        case.update(token_size=ts, tokens_per_patch=tpp, scan_order=order, layout=layout, kind=kind)
        path = os.path.join(GOLD, name + ".pt")
        torch.save(case, path)
        manifest[name] = dict(oracle_vs_ref=e, bytes=os.path.getsize(path))
        print(f"  {name}: oracle_vs_ref {e:.2e}")
    with open(manifest_path, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


class _Wrap:
    """Calls module(h, extra) or module(h) and exposes named_parameters, for run_case."""

    def __init__(self, m, takes_extra):
        self.m, self.takes_extra = m, takes_extra

    def __call__(self, h, extra):
        return self.m(h, extra) if self.takes_extra else self.m(h)

    def named_parameters(self):
        return self.m.named_parameters()


if __name__ == "__main__" and "--maxpool" in sys.argv:
    gen_maxpool()
