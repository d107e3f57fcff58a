"""GPU: the 6-tensor compressed selective_scan_fn (interface.selective_scan_fn_compressed; reference
fastvim_kernel/mamba-1p1p1/faster_mamba_ssm/ops/selective_scan_interface.py:129-252) against vectors produced by the
reference's own selective_scan_ref -- forward, last state and every gradient (the reference's CUDA backward raises for z
and is fp32-only; here both work).  Kept in its own file: added after the round's last GPU slot, it runs after every
other GPU test."""
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
