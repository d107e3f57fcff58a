#!/usr/bin/env python
"""Kernel-level timing of the C-ABI kernels at the BASELINE shapes (CUDA events, rotating buffers
larger than L2 so every launch is L2-cold).  Prints one JSON line per kernel.

    python tools/kbench.py [--shape t224|b224|t2048|c4] [--iters 20] [--only gate,scan,...]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastvim_b200 import ops  # noqa: E402

SHAPES = {  # B, rows, cols, d_model
    "t224": (256, 14, 14, 192), "s224": (256, 14, 14, 384), "b224": (128, 14, 14, 768),
    "t2048": (1, 128, 128, 192), "t448": (64, 28, 28, 192),
}


def timeit(fn, nrot, iters):
    """Average device time per launch: `iters` launches over rotating buffers captured in one CUDA
    graph (so host launch overhead is not in the number), replayed 3 times, best taken."""
    for i in range(3):
        fn(i % nrot)
    torch.cuda.synchronize()
    if os.environ.get("KBENCH_EAGER"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(i % nrot)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e-3
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i % nrot)
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters * 1e-3)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="t224")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default="")
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--rotated", action="store_true")
    a = ap.parse_args()
    Bt, rows, cols, dm = SHAPES[a.shape]
    D, N, R = 2 * dm, 16, (dm + 15) // 16
    L, Lp = rows * cols, rows
    dt = torch.bfloat16 if a.dtype == "bf16" else torch.float32
    s = 2 if dt == torch.bfloat16 else 4
    dev = "cuda"
    geom = ops.Geometry.grid(rows, cols, a.rotated)
    per = Bt * L * 2 * D * s
    nrot = max(2, int(300e6 // per) + 1)
    torch.manual_seed(0)
    xz = [torch.randn(Bt, L, 2 * D, device=dev).to(dt) for _ in range(nrot)]
    cw, cb = torch.randn(2, D, 4, device=dev) * 0.5, torch.randn(2, D, device=dev) * 0.5
    u = [torch.randn(2, Bt, Lp, D, device=dev).to(dt) for _ in range(nrot)]
    xdbl = [(torch.randn(2, Bt * Lp, R + 2 * N, device=dev) * 0.5).to(dt) for _ in range(nrot)]
    dt_w, dt_b = torch.randn(2, D, R, device=dev) * R ** -0.5, torch.rand(2, D, device=dev) - 4.0
    A_log = torch.log(torch.arange(1, N + 1, device=dev).float()).repeat(2, D, 1).contiguous()
    sv = [torch.randn(2, Bt, Lp, D, device=dev) for _ in range(nrot)]
    Dk, lw, lb = torch.ones(2, D, device=dev), torch.ones(D, device=dev), torch.zeros(D, device=dev)
    y = [torch.empty(Bt, L, D, device=dev, dtype=dt) for _ in range(nrot)]
    hs = [torch.randn(Bt, L, dm, device=dev).to(dt) for _ in range(nrot)]
    res = [torch.randn(Bt, L, dm, device=dev) for _ in range(nrot)]
    nw = torch.ones(dm, device=dev)
    only = set(a.only.split(",")) if a.only else None

    def rep(name, t, nbytes):
        print(json.dumps({"kernel": name, "shape": a.shape, "dtype": a.dtype, "us": round(t * 1e6, 2),
                          "alg_MB": round(nbytes / 1e6, 2), "GBps": round(nbytes / t / 1e9, 1),
                          "env": {k: v for k, v in os.environ.items() if k.startswith(("FV_", "FASTVIM_"))}}), flush=True)

    if not only or "conv_pool" in only:
        t = timeit(lambda i: ops.conv_pool_fwd(xz[i][..., :D], geom, cw, cb), nrot, a.iters)
        rep("conv_pool_fwd", t, Bt * L * D * s + 2 * Bt * Lp * D * s)
    if not only or "scan" in only:
        t = timeit(lambda i: ops.scan_fwd(u[i], xdbl[i], geom, R, N, dt_w, dt_b, A_log, True), nrot, a.iters)
        rep("scan_fwd", t, 2 * Bt * Lp * D * s + 2 * Bt * Lp * (R + 2 * N) * s + 2 * Bt * Lp * D * 4)
    if not only or "gate" in only:
        t = timeit(lambda i: ops.gate_fwd(xz[i][..., :D], xz[i][..., D:], sv[i], geom, cw, cb, Dk, lw, lb, 1e-5, out=y[i]),
                   nrot, a.iters)
        rep("gate_fwd", t, 3 * Bt * L * D * s + 2 * Bt * Lp * D * 4)
    if (not only or "block" in only) and ops.block_fwd_supported(geom, Bt, D, dt, R, N):
        xw = (torch.randn(2, R + 2 * N, D, device=dev) * D ** -0.5).to(dt)
        dtwa = dt_w
        xwp = ops.block_pack_xproj(xw)
        t = timeit(lambda i: ops.block_fwd(xz[i][..., :D], xz[i][..., D:], geom, cw, cb, xw, dtwa, dt_b, A_log, Dk,
                                           lw, lb, 1e-5, 1.0, R, N, True, xproj_w_packed=xwp), nrot, a.iters)
        rep("block_fwd", t, 3 * Bt * L * D * s)
    if only and "bwd" in only:   # backward kernels of the training path (MixerFn.backward)
        dxz = [torch.empty(Bt, L, 2 * D, device=dev, dtype=dt) for _ in range(2)]
        dy = [torch.randn(Bt, L, D, device=dev).to(dt) for _ in range(nrot)]
        e_ = [torch.randn(Bt, L, D, device=dev).to(dt) for _ in range(nrot)]
        t = timeit(lambda i: ops.gate_bwd(xz[i][..., :D], xz[i][..., D:], dy[i], sv[i], geom, cw, cb, Dk, lw, lb, 1e-5,
                                          dxz[i & 1][..., D:]), nrot, a.iters)
        rep("gate_bwd", t, 5 * Bt * L * D * s + 3 * Bt * Lp * D * 4)
        tpg = ops.bwd_tiles_per_group(geom, Bt, D, dt)
        dsp = [torch.randn(tpg, Bt, Lp, D, device=dev) for _ in range(2)]
        t = timeit(lambda i: ops.scan_bwd(dsp[i & 1], u[i], xdbl[i], geom, R, N, dt_w, dt_b, A_log, True), nrot, a.iters)
        rep("scan_bwd(+reduce_planes)", t, (tpg * 4 + 6 * s) * Bt * Lp * D + 2 * Bt * Lp * (R + 4 * N) * s)
        t = timeit(lambda i: ops.conv_pool_bwd(xz[i][..., :D], e_[i], u[i], geom, cw, cb, Dk, 1.0, dxz[i & 1][..., :D]),
                   nrot, a.iters)
        rep("conv_pool_bwd", t, 3 * Bt * L * D * s + 2 * Bt * Lp * D * s)
        if ops.gate_bwd_v_supported(geom, Bt, D, dt):
            vv = [torch.randn(Bt, L, D, device=dev).to(dt) for _ in range(nrot)]
            t = timeit(lambda i: ops.gate_bwd_v(vv[i], xz[i][..., D:], dy[i], geom, lw, lb, 1e-5, dxz[i & 1][..., D:]), nrot, a.iters)
            rep("gate_bwd_v", t, 5 * Bt * L * D * s + Bt * Lp * D * 4)
            t = timeit(lambda i: ops.conv_pool_bwd(xz[i][..., :D], e_[i], u[i], geom, cw, cb, Dk, 1.0, dxz[i & 1][..., :D],
                                                   want_dD=True), nrot, a.iters)
            rep("conv_pool_bwd(+dD)", t, 3 * Bt * L * D * s + 2 * Bt * Lp * D * s)
    if not only or "add_norm" in only:
        t = timeit(lambda i: ops.add_norm_fwd(hs[i], res[i], nw, None, 1e-5, True), nrot, a.iters)
        rep("add_norm_fwd", t, Bt * L * dm * (s + 4) * 2)
    if not only or "gemm" in only:
        w_in = torch.randn(2 * D, dm, device=dev).to(dt)
        w_out = torch.randn(dm, D, device=dev).to(dt)
        t = timeit(lambda i: torch.nn.functional.linear(hs[i], w_in), nrot, a.iters)
        rep("in_proj(cublas)", t, Bt * L * (dm + 2 * D) * s)
        t = timeit(lambda i: torch.nn.functional.linear(y[i], w_out), nrot, a.iters)
        rep("out_proj(cublas)", t, Bt * L * (dm + D) * s)
        if dt == torch.bfloat16 and ops.gemm_supported(Bt * L, 2 * D, dm) and ops.gemm_supported(Bt * L, dm, D):
            xzo = [torch.empty(Bt, L, 2 * D, device=dev, dtype=dt) for _ in range(nrot)]
            t = timeit(lambda i: ops.gemm_bf16_tn(hs[i], w_in, out=xzo[i].view(Bt * L, 2 * D)), nrot, a.iters)
            rep("in_proj(tcgen05)", t, Bt * L * (dm + 2 * D) * s)
            t = timeit(lambda i: ops.gemm_bf16_tn(y[i], w_out), nrot, a.iters)
            rep("out_proj(tcgen05)", t, Bt * L * (dm + D) * s)
        if dt == torch.bfloat16 and (not only or "bwdgemm" in only or "gemm" in only):
            # backward GEMMs of in_proj / out_proj: dgrad (bf16 out) and wgrad (fp32 out), cuBLAS vs csrc/gemm_tc2.cu
            M = Bt * L
            dxz = [torch.randn(M, 2 * D, device=dev).to(dt) for _ in range(nrot)]
            dout = [torch.randn(M, dm, device=dev).to(dt) for _ in range(nrot)]
            h2 = [t.view(M, dm) for t in hs]
            y2 = [t.reshape(M, D) for t in y]
            fl = 2.0 * M * dm * 2 * D
            t = timeit(lambda i: dxz[i] @ w_in, nrot, a.iters); rep("in_proj_dgrad(cublas)", t, M * (dm + 2 * D) * s); print(json.dumps({"tflops": fl / t * 1e-12}))
            t = timeit(lambda i: ops.gemm_bf16(dxz[i], w_in, b_mn=True), nrot, a.iters); rep("in_proj_dgrad(tcgen05)", t, M * (dm + 2 * D) * s); print(json.dumps({"tflops": fl / t * 1e-12}))
            t = timeit(lambda i: torch.mm(dxz[i].t(), h2[i], out_dtype=torch.float32), nrot, a.iters); rep("in_proj_wgrad(cublas)", t, M * (dm + 2 * D) * s); print(json.dumps({"tflops": fl / t * 1e-12}))
            t = timeit(lambda i: ops.gemm_bf16(dxz[i], h2[i], a_mn=True, b_mn=True, out_f32=True), nrot, a.iters); rep("in_proj_wgrad(tcgen05)", t, M * (dm + 2 * D) * s); print(json.dumps({"tflops": fl / t * 1e-12}))
            fl = 2.0 * M * dm * D
            t = timeit(lambda i: dout[i] @ w_out, nrot, a.iters); rep("out_proj_dgrad(cublas)", t, M * (dm + D) * s); print(json.dumps({"tflops": fl / t * 1e-12}))
            t = timeit(lambda i: ops.gemm_bf16(dout[i], w_out, b_mn=True), nrot, a.iters); rep("out_proj_dgrad(tcgen05)", t, M * (dm + D) * s); print(json.dumps({"tflops": fl / t * 1e-12}))
            t = timeit(lambda i: torch.mm(dout[i].t(), y2[i], out_dtype=torch.float32), nrot, a.iters); rep("out_proj_wgrad(cublas)", t, M * (dm + D) * s); print(json.dumps({"tflops": fl / t * 1e-12}))
            t = timeit(lambda i: ops.gemm_bf16(dout[i], y2[i], a_mn=True, b_mn=True, out_f32=True), nrot, a.iters); rep("out_proj_wgrad(tcgen05)", t, M * (dm + D) * s); print(json.dumps({"tflops": fl / t * 1e-12}))
        if dt == torch.bfloat16 and ops.gemm_out_norm_supported(Bt * L, dm, D):
            # out_proj + residual add + RMSNorm: one launch (epilogue fusion) vs the two launches it replaces
            yz = [torch.randn(Bt, L, D, device=dev).to(dt) for _ in range(nrot)]
            nb = Bt * L * (D * s + dm * 10)
            t = timeit(lambda i: ops.gemm_out_norm(yz[i], w_out, res[i], nw, 1e-5), nrot, a.iters)
            rep("out_proj+add_norm(fused epilogue)", t, nb)
            t = timeit(lambda i: ops.add_norm_fwd(ops.gemm_bf16_tn(yz[i], w_out), res[i], nw, None, 1e-5, True), nrot, a.iters)
            rep("out_proj+add_norm(two launches)", t, nb + 2 * Bt * L * dm * s)
        pe_in = [torch.randn(Bt * L, 768, device=dev).to(dt) for _ in range(nrot)]
        w_pe, b_pe = torch.randn(dm, 768, device=dev).to(dt), torch.randn(dm, device=dev)
        t = timeit(lambda i: torch.nn.functional.linear(pe_in[i], w_pe, b_pe.to(dt)), nrot, a.iters)
        rep("patch_embed(cublas)", t, Bt * L * (768 + dm) * s)
        if dt == torch.bfloat16 and ops.gemm_supported(Bt * L, dm, 768):
            t = timeit(lambda i: ops.gemm_bf16_tn(pe_in[i], w_pe, bias=b_pe), nrot, a.iters)
            rep("patch_embed(tcgen05, streamed W)", t, Bt * L * (768 + dm) * s)


if __name__ == "__main__":
    main()
