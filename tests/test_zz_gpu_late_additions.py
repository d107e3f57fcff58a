"""GPU tests written after the round's last GPU slot (their host logic is covered on CPU by tests/test_oracle_golden.py and
tests/test_composed_host_logic.py with oracle stand-ins for the kernels); kept in one file that runs after every other GPU
test.

(1) the 6-tensor compressed selective_scan_fn (interface.selective_scan_fn_compressed; reference
fastvim_kernel/mamba-1p1p1/faster_mamba_ssm/ops/selective_scan_interface.py:129-252) against vectors produced by the
reference's own selective_scan_ref -- forward, last state and every gradient (the reference's CUDA backward raises for z
and is fp32-only; here both work).
(2) collapse_method="max" under autograd (cell_imaging/config/FastChannelVimS_maxpool.yaml): forward and every gradient of
the FastChannelVim mixer (both scan orders) and the FastVim mixer against the reference's own modules."""
import pytest
import torch

import fastvim_oracle as O
from util import TOL, assert_close, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["cscan_L128_c8_D", "cscan_L254_c2_noD", "cscan_L196_c14_D_z"])
def test_compressed_scan_fwd_bwd_vs_reference_golden(name):
    from fastvim_b200.interface import selective_scan_fn_compressed

    g = load_golden(name)
    lv = {k: (v.cuda().requires_grad_() if v is not None else None) for k, v in g["inputs"].items()}
    out, st = selective_scan_fn_compressed(lv["u"], lv["u_compressed"], lv["delta"], lv["A"], lv["B"], lv["C"], lv["D"],
                                           z=lv["z"], delta_bias=lv["delta_bias"], delta_softplus=True, return_last_state=True)
    assert_close(out, g["out"], 1e-4, "out")
    assert_close(st, g["last_state"], 1e-4, "last_state")
    out.backward(g["dout"].cuda())
    for k, want in g["grads"].items():
        assert lv[k].grad is not None, k
        assert_close(lv[k].grad, want, 1e-4, "d" + k)


def test_compressed_scan_bf16_vs_oracle():
    from fastvim_b200.interface import selective_scan_fn_compressed

    torch.manual_seed(0)
    bs, dim, L, cfac, N = 2, 64, 196, 14, 16
    Lc = L // cfac
    u = torch.randn(bs, dim, L).bfloat16()
    uc = u.float().reshape(bs, dim, Lc, cfac).mean(3).bfloat16()
    delta, A = (0.5 * torch.rand(bs, dim, Lc)).bfloat16(), -0.5 * torch.rand(dim, N) - 0.05
    B, C, D, bias = torch.randn(bs, N, Lc).bfloat16(), torch.randn(bs, N, Lc).bfloat16(), torch.randn(dim), 0.5 * torch.rand(dim)
    want = O.compressed_scan_oracle(u.float(), uc.float(), delta.float(), A, B.float(), C.float(), D, None, bias, True)
    with torch.no_grad():
        got = selective_scan_fn_compressed(u.cuda(), uc.cuda(), delta.cuda(), A.cuda(), B.cuda(), C.cuda(), D.cuda(), None,
                                           bias.cuda(), True)
    assert got.dtype == torch.bfloat16 and got.shape == (bs, dim, L)
    assert_close(got, want, TOL[torch.bfloat16], "compressed scan bf16")


@pytest.mark.parametrize("name", ["cmixer_max_d32_4x6_t3_channel_first_grads", "cmixer_max_d32_6x4_t2_spatial_first_grads",
                                  "mixer_max_d32_4x6_grads"])
def test_max_pool_training_vs_reference_golden_fp32(name):
    from fastvim_b200.mixer import Mamba
    from fastvim_b200.mixer_channel import Mamba as ChannelMamba

    g = load_golden(name)
    if g["kind"] == "channel":
        m = ChannelMamba(32, token_size=list(g["token_size"]), layer_idx=0, scan_order=g["scan_order"], collapse_method="max")
        call = lambda h: m(h, g["tokens_per_patch"])
    else:
        m = Mamba(32, token_size=list(g["token_size"]), layer_idx=0, collapse_method="max")
        call = lambda h: m(h)
    m.load_state_dict(g["params"], strict=True)
    m = m.cuda().eval()
    with torch.no_grad():
        assert_close(call(g["hidden"].cuda()), g["out"], 1e-4, "out (inference kernels)")
    m.train()
    h = g["hidden"].cuda().requires_grad_()
    out = call(h)
    out.backward(g["dout"].cuda())
    assert_close(out, g["out"], 1e-4, "out (training path)")
    assert_close(h.grad, g["dhidden"], 1e-4, "dhidden")
    got = dict(m.named_parameters())
    for k, want in g["grads"].items():
        assert_close(got[k].grad, want, 1e-4, "d" + k)
