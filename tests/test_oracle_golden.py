"""CPU: the oracle must reproduce the golden vectors generated from the UNMODIFIED reference
(oracle/gen_golden.py).  This is what pins the oracle; the GPU parity tests then compare the
CUDA path against the oracle."""
import json
import os

import pytest
import torch

import fastvim_oracle as O
from util import GOLDEN, assert_close, load_golden, relerr

SCAN_CASES = ["scan_L14_g1_full", "scan_L14_g2_full", "scan_L128_g1_full", "scan_L128_g2_full",
              "scan_L300_g1_full", "scan_L300_g2_full", "scan_L64_plain", "scan_L64_constBC"]


@pytest.mark.parametrize("name", SCAN_CASES)
def test_scan_matches_reference_vectors(name):
    g = load_golden(name)
    ins = {k: (v.clone().requires_grad_() if v is not None else None) for k, v in g["inputs"].items()}
    out, st = O.selective_scan_oracle(ins["u"], ins["delta"], ins["A"], ins["B"], ins["C"], ins["D"], z=ins["z"],
                                      delta_bias=ins["delta_bias"], delta_softplus=g["delta_softplus"],
                                      return_last_state=True)
    out.backward(g["dout"])
    assert relerr(out, g["out"]) < 1e-5
    assert relerr(st, g["last_state"]) < 1e-5
    for k, gr in g["grads"].items():
        assert relerr(ins[k].grad, gr) < 2e-5, k


def test_conv_matches_reference_vector():
    g = load_golden("conv_W4")
    assert relerr(O.causal_conv1d_oracle(g["x"], g["w"], g["b"]), g["out"]) < 1e-6


@pytest.mark.parametrize("name", ["mixer_d32_4x6", "mixer_d32_6x4_nonorm_sf", "mixer_d48_14x14"])
def test_mixer_matches_reference_vectors(name):
    g = load_golden(name)
    p = {k: v.clone().requires_grad_() for k, v in g["params"].items()}
    h = g["hidden"].clone().requires_grad_()
    out = O.mixer_oracle(h, p, g["token_size"], use_norm_after_ssm=g["use_norm_after_ssm"],
                         scaling_factor=g["scaling_factor"])
    out.backward(g["dout"])
    assert relerr(out, g["out"]) < 2e-5
    assert relerr(h.grad, g["dhidden"]) < 5e-5
    for k, gr in g["grads"].items():
        assert relerr(p[k].grad, gr) < 5e-5, k


def test_mamba_inner_matches_reference_vector():
    g = load_golden("mamba_inner")
    out = O.mamba_inner_oracle(g["xz"], g["conv_w"], g["conv_b"], g["x_proj_w"], g["dt_proj_w"], g["A"], None, None,
                               g["D"], g["delta_bias"], True)
    assert relerr(out, g["out"]) < 1e-5


def test_rmsnorm_matches_reference_vector():
    g = load_golden("rmsnorm")
    y, r = O.add_norm_oracle(g["x"], g["weight"], None, g["residual"], g["eps"], True)
    assert relerr(y, g["y"]) < 1e-6 and relerr(r, g["residual_out"]) < 1e-6


def test_small_model_matches_reference_vectors():
    g = load_golden("fastvim_small")
    sd = {k: v.clone().requires_grad_() for k, v in g["state_dict"].items()}
    logits = O.fastvim_oracle(g["images"], sd, depth=g["depth"])
    logits.backward(g["dlogits"])
    assert relerr(logits, g["logits"]) < 1e-5
    for k, gr in g["grads"].items():
        assert relerr(sd[k].grad, gr) < 2e-4, k


def test_manifest_records_full_model_pin():
    m = json.load(open(os.path.join(GOLDEN, "manifest.json")))
    assert m["fastvim_tiny_224_full"]["oracle_vs_ref"] < 1e-4


def test_rotation_is_a_pure_index_map():
    """models/fastvim.py:192-210: rotate then inverse-rotate is the identity, bit-exact."""
    h = torch.randn(2, 6 * 4, 5)
    r = O.rotate_tokens(h, 6, 4)
    assert torch.equal(O.rotate_tokens(r, 4, 6), h)
    # element (row i, col j) moves to position j*rows + i
    assert torch.equal(r[:, 2 * 6 + 3], h[:, 3 * 4 + 2])


def test_pool_index_matches_reshape_mean():
    L, outer, pool, inner = 24, 3, 4, 2
    idx = O.pool_index(L, outer, pool, inner)
    x = torch.randn(1, 1, L)
    ref = O.pool_oracle(x, outer, pool, inner)
    acc = torch.zeros(outer * inner).index_add_(0, idx, x[0, 0]) / pool
    assert torch.allclose(acc, ref[0, 0], atol=1e-6)
    assert torch.equal(O.broadcast_oracle(ref, outer, pool, inner)[0, 0], ref[0, 0][idx])


@pytest.mark.parametrize("name", ["cmixer_d32_4x6_t3_channel_first", "cmixer_d32_6x4_t2_spatial_first"])
def test_channel_mixer_oracle_matches_reference_vector(name):
    """(outer, pool, inner) generalisation of the mixer oracle against the reference's own FastChannelVim mixer
    (mamba_simple_channel_faster.py:176-420), both scan orders."""
    g = load_golden(name)
    ts, tpp = g["token_size"], g["tokens_per_patch"]
    layout = (ts[0], ts[1], tpp) if g["scan_order"] == "Channel-First" else (tpp * ts[0], ts[1], 1)
    out = O.mixer_oracle(g["hidden"], g["params"], ts, layout=layout)
    assert relerr(out, g["out"]) < 2e-5


MASKED = ["mmixer_d32_4x6_keep10", "mmixer_d32_6x4_keep24_full", "mmixer_v2_d48_14x14_keep49_nonorm"]


def _oracle_grads(g, **kw):
    p = {k: v.clone().requires_grad_() for k, v in g["params"].items()}
    h = g["hidden"].clone().requires_grad_()
    out = O.mixer_oracle(h, p, g["token_size"], **kw)
    out.backward(g["dout"])
    return out.detach(), h.grad, {k: v.grad for k, v in p.items()}


@pytest.mark.parametrize("name", MASKED)
def test_masked_mixer_oracle_matches_reference_vectors(name):
    """FastMaskVim: oracle forward + every gradient against the reference's own Mamba_masked (v1 and v2 modules)."""
    g = load_golden(name)
    out, dh, grads = _oracle_grads(g, use_norm_after_ssm=g["use_norm_after_ssm"], ids_keep=g["ids_keep"])
    assert_close(out, g["out"], 5e-5, "out")
    assert_close(dh, g["dhidden"], 5e-5, "dhidden")
    for k, want in g["grads"].items():
        assert_close(grads[k], want, 5e-5, "d" + k)


def test_masked_mixer_with_all_tokens_kept_equals_fastvim_mixer():
    """the masked definition coincides with the FastVim mixer when nothing is masked."""
    g = load_golden("mmixer_d32_6x4_keep24_full")
    assert torch.equal(g["ids_keep"], torch.arange(24)[None].expand(2, -1))
    plain = O.mixer_oracle(g["hidden"], g["params"], g["token_size"])
    assert_close(plain, g["out"], 5e-5, "masked(all kept) vs plain")


@pytest.mark.parametrize("name", ["cmixer_d32_4x6_t3_channel_first_grads", "cmixer_d32_6x4_t2_spatial_first_grads"])
def test_channel_mixer_oracle_gradients_match_reference_vectors(name):
    g = load_golden(name)
    tpp, ts = g["tokens_per_patch"], g["token_size"]
    layout = (ts[0], ts[1], tpp) if g["scan_order"] == "Channel-First" else (tpp * ts[0], ts[1], 1)
    out, dh, grads = _oracle_grads(g, layout=layout)
    assert_close(out, g["out"], 5e-5, "out")
    assert_close(dh, g["dhidden"], 5e-5, "dhidden")
    for k, want in g["grads"].items():
        assert_close(grads[k], want, 5e-5, "d" + k)


@pytest.mark.parametrize("name", ["channelvim_small_cf", "channelvim_small_sf"])
def test_channel_model_state_dict_is_reference_compatible(name):
    """The FastChannelVim wrapper keeps the reference's parameter names and shapes (strict load of a state dict written by
    the reference's own VisionMamba); construction needs no GPU."""
    from fastvim_b200.vision_channel import VisionMamba

    g = load_golden(name)
    m = VisionMamba(**g["kwargs"], rms_norm=True, residual_in_fp32=True, fused_add_norm=True, final_pool_type="mean",
                    if_abs_pos_embed=True, drop_path_rate=0.0, scan_order=g["scan_order"], hcs=False)
    m.load_state_dict(g["state_dict"], strict=True)
    assert set(g["grads"]) == {k for k, _ in m.named_parameters()}
    # tokenisation order of the per-channel patch embedding is a pure index map (bit-exact against a Conv3d)
    x = torch.randn(2, 3, 32, 64)
    with torch.no_grad():
        ours, tpp, _, _, chans = m.patch_embed(x)
        ref = m.patch_embed.proj(x.unsqueeze(1)) + m.patch_embed.channel_embed.weight.t()[None, :, :, None, None]
    ref = ref.permute(0, 1, 3, 4, 2) if g["scan_order"] == "Channel-First" else ref
    ref = ref.flatten(2).transpose(1, 2)
    assert tpp == 3 and chans == [0, 1, 2] and ours.shape == ref.shape
    assert torch.allclose(ours, ref, atol=1e-5)


@pytest.mark.parametrize("name", ["cmixer2d_d32_4x6_t3_layer0_rows", "cmixer2d_d32_4x6_t3_layer2_channels"])
def test_2dcompress_mixer_oracle_matches_reference_vectors(name):
    """FastChannelVim 2dcompress mixer (mamba_simple_channel_faster_2dcompress.py): both layer kinds are (outer, pool, inner)
    layouts of the same oracle; forward and every gradient against the reference's own module."""
    g = load_golden(name)
    out, dh, grads = _oracle_grads(g, layout=g["layout"])
    assert_close(out, g["out"], 5e-5, "out")
    assert_close(dh, g["dhidden"], 5e-5, "dhidden")
    for k, want in g["grads"].items():
        assert_close(grads[k], want, 5e-5, "d" + k)


def test_masked_block_rotate_indices_and_state_dict():
    """Block_masked: the id rotation table equals the reference's double loop (models_mamba_faster_mae_vimdecoder_v2.py
    :320-328) and the block stack loads a state dict written by the reference's own blocks."""
    from fastvim_b200.vision_masked import Block_masked, create_block_masked

    H, W = 4, 6
    want = torch.zeros(H * W, dtype=torch.long)
    for i in range(H):
        for j in range(W):
            want[i * W + j] = j * H + i
    assert torch.equal(Block_masked.compute_rotate_indices(H, W), want)
    g = load_golden("mblocks_d32_4x6_keep10")
    layers = torch.nn.ModuleList([create_block_masked(32, rms_norm=True, residual_in_fp32=True, fused_add_norm=True,
                                                      layer_idx=i, token_size=g["token_size"]) for i in range(g["depth"])])
    sd = {k[len("layers."):]: v for k, v in g["state_dict"].items() if k.startswith("layers.")}
    layers.load_state_dict(sd, strict=True)


CSCAN = ["cscan_L128_c8_D", "cscan_L254_c2_noD", "cscan_L196_c14_D_z"]


@pytest.mark.parametrize("name", CSCAN)
def test_compressed_scan_oracle_matches_reference_vectors(name):
    """6-tensor compressed scan (fastvim_kernel/.../faster_mamba_ssm/ops/selective_scan_interface.py:162-252): oracle forward,
    last state and autograd gradients against the reference's own selective_scan_ref."""
    g = load_golden(name)
    lv = {k: (v.clone().requires_grad_() if v is not None else None) for k, v in g["inputs"].items()}
    out, st = O.compressed_scan_oracle(lv["u"], lv["u_compressed"], lv["delta"], lv["A"], lv["B"], lv["C"], lv["D"], z=lv["z"],
                                       delta_bias=lv["delta_bias"], delta_softplus=True, return_last_state=True)
    out.backward(g["dout"])
    assert_close(out, g["out"], 2e-5, "out")
    assert_close(st, g["last_state"], 2e-5, "last_state")
    for k, want in g["grads"].items():
        assert_close(lv[k].grad, want, 2e-5, "d" + k)


@pytest.mark.parametrize("name", CSCAN)
def test_compressed_scan_host_glue_with_cpu_stand_ins(name, monkeypatch):
    """The host glue of interface.selective_scan_fn_compressed (argument order, compression factor, D skip on the
    full-resolution u, z gate, last state) checked on CPU: the two kernels it calls are replaced by oracle stand-ins with the
    kernels' exact signatures.  (The kernels themselves are covered by the GPU parity tests.)"""
    from fastvim_b200 import interface, ops

    def scan_fwd(u, delta, A, B, Cm, D, z, delta_bias, delta_softplus, want_last_state=False):
        assert B.dim() == 4 and Cm.dim() == 4 and u.is_contiguous() and delta.is_contiguous()
        out, last = O.selective_scan_oracle(u, delta, A, B, Cm, D, z, delta_bias, delta_softplus, return_last_state=True)
        return out, (last if want_last_state else None)

    def bcast(s, xc, Dskip, outer, pool, inner=1):
        assert s.shape[-1] == outer * inner and (xc is None or xc.shape[-1] == outer * pool * inner)
        v = O.broadcast_oracle(s, outer, pool, inner)
        return v if Dskip is None else v + Dskip[None, :, None] * xc

    monkeypatch.setattr(ops, "selective_scan_fwd", scan_fwd)
    monkeypatch.setattr(ops, "bcast_skip_bdl_fwd", bcast)
    g = load_golden(name)
    i = g["inputs"]
    with torch.no_grad():
        out, st = interface.selective_scan_fn_compressed(i["u"], i["u_compressed"], i["delta"], i["A"], i["B"], i["C"], i["D"],
                                                         z=i["z"], delta_bias=i["delta_bias"], delta_softplus=True,
                                                         return_last_state=True)
        out_only = interface.selective_scan_fn_compressed(i["u"], i["u_compressed"], i["delta"], i["A"], i["B"], i["C"], i["D"],
                                                          z=i["z"], delta_bias=i["delta_bias"], delta_softplus=True)
    assert_close(out, g["out"], 2e-5, "out")
    assert_close(st, g["last_state"], 2e-5, "last_state")
    assert torch.equal(out, out_only)
    with pytest.raises(ValueError):
        interface.selective_scan_fn_compressed(i["u"][..., :-1], i["u_compressed"], i["delta"], i["A"], i["B"], i["C"])


# ------------------------------------------------------------------ model-level oracles added in round 2
@pytest.mark.parametrize("name", ["channelvim_small_cf", "channelvim_small_sf"])
def test_channel_model_oracle_matches_reference_vectors(name):
    """``channelvim_oracle`` (FastChannelVim classifier, BASELINE.json configs[3]) is pinned on the reference's own
    ``VisionMamba`` of models_channel_mamba_faster.py: logits and every parameter gradient."""
    g = load_golden(name)
    sd = {k: v.clone().double().requires_grad_(True) for k, v in g["state_dict"].items()}
    logits = O.channelvim_oracle(g["images"].double(), sd, depth=g["kwargs"]["depth"], scan_order=g["scan_order"])
    logits.backward(g["dlogits"].double())
    assert_close(logits, g["logits"], 5e-5, "logits")
    for k, want in g["grads"].items():
        assert_close(sd[k].grad, want, 1e-4, "d " + k)


def test_masked_blocks_oracle_matches_reference_vectors():
    """``masked_blocks_oracle`` (FastMaskVim encoder stack with the odd-layer id rotation) is pinned on the reference's own
    ``Block_masked`` stack: output, input gradient and every parameter gradient."""
    g = load_golden("mblocks_d32_4x6_keep10")
    sd = {k: v.clone().double().requires_grad_(True) for k, v in g["state_dict"].items()}
    h = g["hidden"].clone().double().requires_grad_(True)
    out = O.masked_blocks_oracle(h, sd, g["ids_keep"], g["token_size"], depth=g["depth"])
    out.backward(g["dout"].double())
    assert_close(out, g["out"], 5e-5, "out")
    assert_close(h.grad, g["dhidden"], 1e-4, "dhidden")
    for k, want in g["grads"].items():
        assert_close(sd[k].grad, want, 1e-4, "d " + k)
