"""CPU stand-ins with the exact signatures of the ``fastvim_b200.ops`` kernel wrappers that the operator-by-operator
(``composed``) path and the operator API call, built on the oracle.  They let the HOST LOGIC of those paths (argument
order, layouts, autograd plumbing) run in the CPU test suite; the kernels themselves are covered by the GPU parity tests."""
import torch

import fastvim_oracle as O


def causal_conv1d_fwd(x, weight, bias, silu=True):
    return O.causal_conv1d_oracle(x, weight.to(x.dtype), None if bias is None else bias.to(x.dtype),
                                  activation="silu" if silu else None).contiguous()


def causal_conv1d_bwd(x, weight, bias, dout, silu=True):
    with torch.enable_grad():
        xl = x.detach().clone().requires_grad_()
        wl = weight.detach().float().clone().requires_grad_()
        bl = None if bias is None else bias.detach().float().clone().requires_grad_()
        out = O.causal_conv1d_oracle(xl, wl.to(x.dtype), None if bl is None else bl.to(x.dtype),
                                     activation="silu" if silu else None)
        grads = torch.autograd.grad(out, [xl, wl] + ([bl] if bl is not None else []), dout)
    return grads[0], grads[1], (grads[2] if bl is not None else None)


def pool_bdl_fwd(xc, outer, pool, inner=1, mode="mean", scale=1.0):
    return O.pool_oracle(xc, outer, pool, inner, mode, scale).contiguous()


def bcast_skip_bdl_fwd(s, xc, Dskip, outer, pool, inner=1):
    v = O.broadcast_oracle(s, outer, pool, inner)
    return (v if Dskip is None else v + Dskip.to(v.dtype)[None, :, None] * xc).contiguous()


def rowdot_bdl(a, c):
    return (a.float() * c.float()).sum((0, 2))


def selective_scan_fwd(u, delta, A, B, Cm, D, z, delta_bias, delta_softplus, want_last_state=False):
    assert B.dim() == 4 and Cm.dim() == 4
    out, last = O.selective_scan_oracle(u, delta, A, B, Cm, D, z, delta_bias, delta_softplus, return_last_state=True)
    return out, (last if want_last_state else None)


def selective_scan_bwd(dout, u, delta, A, B, Cm, D, z, delta_bias, delta_softplus):
    with torch.enable_grad():
        leaves = [t.detach().clone().requires_grad_() if t is not None else None for t in (u, delta, A, B, Cm, D, z, delta_bias)]
        out = O.selective_scan_oracle(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], leaves[5], leaves[6], leaves[7],
                                      delta_softplus)
        live = [t for t in leaves if t is not None]
        g = list(torch.autograd.grad(out, live, dout))
    res = [g.pop(0) if t is not None else None for t in leaves]
    du, ddelta, dA, dB, dC, dD, dz, dbias = res
    return du, ddelta, dA, dB, dC, dD, dz, dbias


def install(monkeypatch):
    from fastvim_b200 import ops

    for name in ("causal_conv1d_fwd", "causal_conv1d_bwd", "pool_bdl_fwd", "bcast_skip_bdl_fwd", "rowdot_bdl",
                 "selective_scan_fwd", "selective_scan_bwd"):
        monkeypatch.setattr(ops, name, globals()[name])
