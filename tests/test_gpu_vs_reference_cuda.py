"""GPU parity against the reference's OWN CUDA selective scan (mamba-1p1p1/csrc/selective_scan, compiled
unmodified for sm_100a by oracle/build_ref.py into oracle/_ref/selective_scan_cuda.so -- checker only).

BASELINE.json: "Results must match the reference's own CUDA selective_scan ... on identical synthetic
inputs": fp32 within 1e-4 relative, bf16 within 2e-2 relative.  Inputs follow the reference's own test
generator (tests/ops/test_selective_scan.py:61-122: seed 0, A = -0.5 rand, delta = 0.5 rand, ...)."""
import importlib.util
import os

import pytest
import torch

from util import assert_close

pytestmark = pytest.mark.gpu
SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "selective_scan_cuda.so")


@pytest.fixture(scope="module")
def ref_cuda():
    if not os.path.exists(SO):
        pytest.skip("oracle/_ref/selective_scan_cuda.so not built (python oracle/build_ref.py in the build container)")
    spec = importlib.util.spec_from_file_location("selective_scan_cuda", SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _inputs(batch, dim, L, N, dtype, has_z, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    r = lambda *s: torch.rand(*s, device="cuda", generator=g)
    n = lambda *s: torch.randn(*s, device="cuda", generator=g)
    A = -0.5 * r(dim, N)
    B, C = n(batch, 1, N, L).to(dtype), n(batch, 1, N, L).to(dtype)
    D = n(dim)
    z = n(batch, dim, L).to(dtype) if has_z else None
    delta_bias = 0.5 * r(dim)
    u = n(batch, dim, L).to(dtype)
    delta = (0.5 * r(batch, dim, L)).to(dtype)
    return u, delta, A, B, C, D, z, delta_bias


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(256, 384, 14, 16), (32, 768, 112, 16), (1, 384, 128, 16), (2, 4, 2048, 8),
                                   (2, 4, 4096, 8), (3, 20, 300, 16)])
@pytest.mark.parametrize("has_z", [False, True])
def test_selective_scan_fwd_matches_reference_cuda(ref_cuda, dtype, shape, has_z):
    from fastvim_b200.interface import selective_scan_fn

    batch, dim, L, N = shape
    u, delta, A, B, C, D, z, db = _inputs(batch, dim, L, N, dtype, has_z)
    outs = ref_cuda.fwd(u, delta, A, B, C, D, z, db, True)
    want = outs[-1] if has_z else outs[0]
    want_last = outs[1][:, :, -1, 1::2]   # selective_scan_interface.py:50
    with torch.no_grad():
        got, last = selective_scan_fn(u, delta, A, B, C, D, z=z, delta_bias=db, delta_softplus=True,
                                      return_last_state=True)
    # both sides round the output to `dtype` once; bf16: 1 ulp = 2^-8 relative
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    assert_close(got, want, tol, "out vs reference CUDA")
    assert_close(last, want_last.float(), 1e-4 if dtype == torch.float32 else 2e-2, "last_state vs reference CUDA")


# the four op shapes of SURVEY.md 8d (C2, C3, C4, C5) + one ragged one
BWD_SHAPES = [(256, 384, 14, 16), (128, 1536, 14, 16), (32, 768, 112, 16), (1, 384, 128, 16), (3, 20, 300, 16)]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", BWD_SHAPES)
@pytest.mark.parametrize("has_z", [False, True])
def test_selective_scan_bwd_matches_reference_cuda(ref_cuda, dtype, shape, has_z):
    """``fv_selective_scan_bwd`` beside the reference's own ``selective_scan_cuda.bwd`` (selective_scan.cpp:338-492,
    called as in selective_scan_interface.py:59-102) on identical tensors: every gradient."""
    from fastvim_b200 import ops

    batch, dim, L, N = shape
    u, delta, A, B, C, D, z, db = _inputs(batch, dim, L, N, dtype, has_z)
    g = torch.Generator(device="cuda").manual_seed(7)
    dout = torch.randn(batch, dim, L, device="cuda", generator=g).to(dtype)
    outs = ref_cuda.fwd(u, delta, A, B, C, D, z, db, True)
    out, x = outs[0], outs[1]
    want = ref_cuda.bwd(u, delta, A, B, C, D, z, db, dout, x, out if has_z else None, None, True, False)
    w_du, w_ddelta, w_dA, w_dB, w_dC, w_dD, w_dbias = want[:7]
    got = ops.selective_scan_bwd(dout, u, delta, A, B, C, D, z, db, True)
    du, ddelta, dA, dB, dC, dD, dz, dbias = got
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    # fp32: both kernels use fast-math exp2 / atomics in different orders -> 2e-4 on the reductions over (batch, L)
    rtol = 2e-4 if dtype == torch.float32 else 2e-2
    assert_close(du, w_du, tol, "du")
    assert_close(ddelta, w_ddelta, tol, "ddelta")
    assert_close(dA, w_dA, rtol, "dA")
    assert_close(dB, w_dB, rtol, "dB")
    assert_close(dC, w_dC, rtol, "dC")
    assert_close(dD, w_dD, rtol, "dD")
    assert_close(dbias, w_dbias, rtol, "ddelta_bias")
    if has_z:
        assert_close(dz, want[7], tol, "dz")
