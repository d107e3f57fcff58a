// K2b-bwd, streaming form -- backward of (direction average + LayerNorm + SiLU gate) from the SAVED pre-norm value v.
//
// Reference: autograd through mamba_ssm/modules/mamba_simple_faster.py:434-453 (flip/add//2, LayerNorm, * silu(z)) and
// the D-skip / repeat_interleave of :356-358 (fused-autograd form selective_scan_interface.py:662-675).
//
// Round 1's fv_gate_bwd recomputed v from x (both convolutions, 2 SiLUs per element) inside 8-token shared-memory
// tiles with three block-wide phases: 386 us per launch at FastVim-B, 5.5x its HBM time, and a two-pass streaming
// variant that also recomputed was slower still.  The fix is not to recompute: the fused forward
// (block_cluster.cu) holds v = (s_f + s_b + D_f xc_f + D_b xc_b) / 2 in registers anyway and now writes it once (bf16,
// as the reference's own (out + out_b.flip) / 2 is).  With v given, the backward is channel-local except for FOUR
// per-token sums over d_inner, and all four can be formed in ONE reduction:
//   Sv = sum v, Svv = sum v^2                        -> mean, rstd
//   Sg = sum dxh, Sgv = sum dxh v,  dxh = dy silu(z) gamma   -> c1 = Sg / D, c2 = rstd (Sgv - mean Sg) / D
//   xhat = (v - mean) rstd;  dz = dy (xhat gamma + beta) silu'(z);  dv = rstd (dxh - c1 - xhat c2);  e = dv / 2
// so the kernel streams v, z, dy once and writes dz, e once: 5 T of traffic, no shared-memory tiles, no recompute.
// Mapping: a group of D / 384 warps owns one pooled row of one image and walks its `pool` tokens (lane = 4-channel chunks
// lane, lane + 32, lane + 64 of the warp's 384 channels, the layout of the forward gate): the pooled gradient
// ds[j] = sum over the row's tokens of e is a register accumulator written once (no partial planes, no atomics), next
// token's loads are issued before the current token's math, dgamma / dbeta accumulate in registers over all rows of the
// thread and leave through one shared-memory reduction + one atomic per CTA and channel.  The D-skip gradients
// dD_f = sum e xc_f, dD_b = sum e xc_b need the conv outputs and are formed by fv_conv_pool_bwd, which recomputes them
// for the conv backward anyway.

#include "block_common.cuh"

namespace fv {

int sm_count();
int check_geom(const fv_geom* g, const char* who);

constexpr int GV_THREADS = 256;
constexpr int GV_WARPS = GV_THREADS / 32;
constexpr int GV_NC = 3;  // 4-channel chunks per lane: one warp covers 384 channels

struct GateBwdVArgs {
    Geom g;
    const bf16* v;      // (B, L, D) memory token order
    const bf16* z;
    int64_t ldz, zbs;
    const bf16* dy;
    int64_t lddy, dybs;
    const float* lnw;
    const float* lnb;
    float eps;
    bf16* dz;           // same strides as z
    bf16* e;            // (B, L, D)
    float* ds;          // (B, Lp, D)
    float* dlnw;
    float* dlnb;
    int nitems;         // B * outer
};

// silu(z) and silu'(z) of a pair from ONE tanh per element: s = sigmoid(z) = 0.5 + 0.5 tanh(z / 2)
__device__ __forceinline__ void silu_grad2(float2 zv, float2& sl, float2& dsl) {
    const float2 half2c = make_float2(0.5f, 0.5f), one2 = make_float2(1.f, 1.f);
    const float2 h = __fmul2_rn(zv, half2c);
    const float2 s = __ffma2_rn(make_float2(bk_tanh(h.x), bk_tanh(h.y)), half2c, half2c);
    sl = __fmul2_rn(zv, s);
    dsl = __ffma2_rn(sl, __ffma2_rn(s, make_float2(-1.f, -1.f), one2), s);   // s + silu (1 - s)
}

// NWG: warps per token group (d_inner / 384)
template <bool NORM, int NWG>
__global__ void __launch_bounds__(GV_THREADS, 2) gate_bwd_v_kernel(const GateBwdVArgs a) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    extern __shared__ __align__(16) unsigned char gv_smem[];
    constexpr int NGRP = GV_WARPS / NWG;   // token groups per CTA
    const Geom& g = a.g;
    const int D = g.D, P = g.pool, outer = g.outer;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = warp / NWG, wg = warp - grp * NWG;   // group, warp within the group
    const int cbase = wg * 384 + lane * 4;               // first channel of chunk 0 (chunks at +128 i)
    float* sgam = reinterpret_cast<float*>(gv_smem);     // [D]
    float* sbet = sgam + D;                               // [D]
    float4* red = reinterpret_cast<float4*>(sbet + D);    // [2][GV_WARPS]
    if (NORM) {
        for (int i = tid; i < D; i += GV_THREADS) {
            sgam[i] = a.lnw[i];
            sbet[i] = a.lnb ? a.lnb[i] : 0.f;
        }
    }
    __syncthreads();
    const float invD = 1.f / (float)D;
    const int so = (int)g.so, sp = (int)g.sp;
    float2 dgam[GV_NC][2], dbet[GV_NC][2];
#pragma unroll
    for (int i = 0; i < GV_NC; ++i) dgam[i][0] = dgam[i][1] = dbet[i][0] = dbet[i][1] = make_float2(0.f, 0.f);
    const int step = gridDim.x * NGRP;
    const int n_iter = (a.nitems + step - 1) / step;
    int par = 0;
    for (int it = 0; it < n_iter; ++it) {
        const int item = (it * gridDim.x + blockIdx.x) * NGRP + grp;
        const bool live = item < a.nitems;
        const int b = live ? item / outer : 0, j = live ? item - b * outer : 0;
        const bf16* vb = a.v + (int64_t)b * g.L * D + cbase;
        const bf16* zb = a.z + (int64_t)b * a.zbs + cbase;
        const bf16* dyb = a.dy + (int64_t)b * a.dybs + cbase;
        bf16* dzb = a.dz + (int64_t)b * a.zbs + cbase;
        bf16* eb = a.e + (int64_t)b * g.L * D + cbase;
        float2 dsa[GV_NC][2];
#pragma unroll
        for (int i = 0; i < GV_NC; ++i) dsa[i][0] = dsa[i][1] = make_float2(0.f, 0.f);
        uint2 qv[GV_NC], qz[GV_NC], qd[GV_NC];
        {
            const int row = j * so;   // token p = 0
#pragma unroll
            for (int i = 0; i < GV_NC; ++i) {
                qv[i] = qz[i] = qd[i] = make_uint2(0u, 0u);
                if (live) {
                    qv[i] = __ldg(reinterpret_cast<const uint2*>(vb + (int64_t)row * D + 128 * i));
                    qz[i] = __ldg(reinterpret_cast<const uint2*>(zb + (int64_t)row * a.ldz + 128 * i));
                    qd[i] = __ldg(reinterpret_cast<const uint2*>(dyb + (int64_t)row * a.lddy + 128 * i));
                }
            }
        }
        for (int p = 0; p < P; ++p) {
            const int row = j * so + p * sp;
            uint2 cv[GV_NC], cz[GV_NC], cd[GV_NC];
#pragma unroll
            for (int i = 0; i < GV_NC; ++i) { cv[i] = qv[i]; cz[i] = qz[i]; cd[i] = qd[i]; }
            if (live && p + 1 < P) {   // next token's rows: in flight during this token's math
                const int rn = row + sp;
#pragma unroll
                for (int i = 0; i < GV_NC; ++i) {
                    qv[i] = __ldg(reinterpret_cast<const uint2*>(vb + (int64_t)rn * D + 128 * i));
                    qz[i] = __ldg(reinterpret_cast<const uint2*>(zb + (int64_t)rn * a.ldz + 128 * i));
                    qd[i] = __ldg(reinterpret_cast<const uint2*>(dyb + (int64_t)rn * a.lddy + 128 * i));
                }
            }
            // ---- per-element forward pieces and the four per-token sums
            float2 vv[GV_NC][2], gg[GV_NC][2], dn[GV_NC][2], dyd[GV_NC][2];   // v, dxh = dn gamma, dn = dy silu(z), dy silu'(z)
            float2 s_v = make_float2(0.f, 0.f), s_vv = s_v, s_g = s_v, s_gv = s_v;
#pragma unroll
            for (int i = 0; i < GV_NC; ++i) {
                const uint32_t wv[2] = {cv[i].x, cv[i].y}, wz[2] = {cz[i].x, cz[i].y}, wd[2] = {cd[i].x, cd[i].y};
                float4 gm = make_float4(1.f, 1.f, 1.f, 1.f);
                if (NORM) gm = *reinterpret_cast<const float4*>(sgam + cbase + 128 * i);
                const float2 gm2[2] = {make_float2(gm.x, gm.y), make_float2(gm.z, gm.w)};
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float2 v2 = unpack2(wv[k]), z2 = unpack2(wz[k]), d2 = unpack2(wd[k]);
                    float2 sl, dsl;
                    silu_grad2(z2, sl, dsl);
                    vv[i][k] = v2;
                    dn[i][k] = __fmul2_rn(d2, sl);
                    dyd[i][k] = __fmul2_rn(d2, dsl);
                    gg[i][k] = __fmul2_rn(dn[i][k], gm2[k]);
                    if (NORM) {
                        s_v = __fadd2_rn(s_v, v2);
                        s_vv = __ffma2_rn(v2, v2, s_vv);
                        s_g = __fadd2_rn(s_g, gg[i][k]);
                        s_gv = __ffma2_rn(gg[i][k], v2, s_gv);
                    }
                }
            }
            float nmean = 0.f, rstd = 1.f, c1 = 0.f, c2 = 0.f;
            if (NORM) {
                float4 r = make_float4(s_v.x + s_v.y, s_vv.x + s_vv.y, s_g.x + s_g.y, s_gv.x + s_gv.y);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    r.x += __shfl_xor_sync(0xffffffffu, r.x, o);
                    r.y += __shfl_xor_sync(0xffffffffu, r.y, o);
                    r.z += __shfl_xor_sync(0xffffffffu, r.z, o);
                    r.w += __shfl_xor_sync(0xffffffffu, r.w, o);
                }
                if (NWG > 1) {
                    if (lane == 0) red[par * GV_WARPS + warp] = r;
                    __syncthreads();
                    r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int w = 0; w < NWG; ++w) {   // fixed order: identical on every warp of the group
                        const float4 q = red[par * GV_WARPS + grp * NWG + w];
                        r.x += q.x; r.y += q.y; r.z += q.z; r.w += q.w;
                    }
                    par ^= 1;
                }
                const float mean = r.x * invD;
                rstd = rsqrtf(fmaxf(fmaf(r.y, invD, -mean * mean), 0.f) + a.eps);
                nmean = -mean;
                c1 = r.z * invD;
                c2 = rstd * (r.w - mean * r.z) * invD;
            }
            // ---- outputs of this token
            if (live) {
                const float2 nm2 = make_float2(nmean, nmean), rs2 = make_float2(rstd, rstd);
                const float2 nc1 = make_float2(-c1, -c1), nc2 = make_float2(-c2, -c2), hr2 = make_float2(0.5f * rstd, 0.5f * rstd);
#pragma unroll
                for (int i = 0; i < GV_NC; ++i) {
                    float4 bt = make_float4(0.f, 0.f, 0.f, 0.f), gm = make_float4(1.f, 1.f, 1.f, 1.f);
                    if (NORM) {
                        bt = *reinterpret_cast<const float4*>(sbet + cbase + 128 * i);
                        gm = *reinterpret_cast<const float4*>(sgam + cbase + 128 * i);
                    }
                    const float2 bt2[2] = {make_float2(bt.x, bt.y), make_float2(bt.z, bt.w)};
                    const float2 gm2[2] = {make_float2(gm.x, gm.y), make_float2(gm.z, gm.w)};
                    uint32_t odz[2], oe[2];
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        float2 ev, nrm;
                        if (NORM) {
                            const float2 xh = __fmul2_rn(__fadd2_rn(vv[i][k], nm2), rs2);
                            nrm = __ffma2_rn(xh, gm2[k], bt2[k]);
                            ev = __fmul2_rn(__ffma2_rn(xh, nc2, __fadd2_rn(gg[i][k], nc1)), hr2);   // dv / 2
                            dgam[i][k] = __ffma2_rn(dn[i][k], xh, dgam[i][k]);
                            dbet[i][k] = __fadd2_rn(dbet[i][k], dn[i][k]);
                        } else {
                            nrm = vv[i][k];
                            ev = __fmul2_rn(dn[i][k], make_float2(0.5f, 0.5f));
                        }
                        odz[k] = pack2(__fmul2_rn(dyd[i][k], nrm));
                        oe[k] = pack2(ev);
                        dsa[i][k] = __fadd2_rn(dsa[i][k], ev);
                    }
                    *reinterpret_cast<uint2*>(dzb + (int64_t)row * a.ldz + 128 * i) = make_uint2(odz[0], odz[1]);
                    *reinterpret_cast<uint2*>(eb + (int64_t)row * D + 128 * i) = make_uint2(oe[0], oe[1]);
                }
            }
        }
        if (live) {
            float* dsr = a.ds + ((int64_t)b * outer + j) * D + cbase;
#pragma unroll
            for (int i = 0; i < GV_NC; ++i)
                *reinterpret_cast<float4*>(dsr + 128 * i) = make_float4(dsa[i][0].x, dsa[i][0].y, dsa[i][1].x, dsa[i][1].y);
        }
    }
    if (NORM) {
        // dgamma / dbeta: sum the CTA's token groups in shared memory (group by group, fixed order), then one atomic per channel
        __syncthreads();
        float* acc = sgam;   // reuse: [2][D]
        for (int gsel = 0; gsel < NGRP; ++gsel) {
            if (grp == gsel) {
#pragma unroll
                for (int i = 0; i < GV_NC; ++i) {
                    float4* pg = reinterpret_cast<float4*>(acc + cbase + 128 * i);
                    float4* pb = reinterpret_cast<float4*>(acc + D + cbase + 128 * i);
                    float4 og = make_float4(0.f, 0.f, 0.f, 0.f), ob = og;
                    if (gsel > 0) { og = *pg; ob = *pb; }
                    *pg = make_float4(og.x + dgam[i][0].x, og.y + dgam[i][0].y, og.z + dgam[i][1].x, og.w + dgam[i][1].y);
                    *pb = make_float4(ob.x + dbet[i][0].x, ob.y + dbet[i][0].y, ob.z + dbet[i][1].x, ob.w + dbet[i][1].y);
                }
            }
            __syncthreads();
        }
        for (int i = tid; i < D; i += GV_THREADS) {
            atomicAdd(a.dlnw + i, acc[i]);
            if (a.dlnb) atomicAdd(a.dlnb + i, acc[D + i]);
        }
    }
}

}  // namespace fv

extern "C" int fv_gate_bwd_v_supported(const fv_geom* g, int dtype) {
    if (!g || dtype != FV_BF16 || g->inner != 1 || g->dim <= 0) return 0;
    const int nwg = g->dim / 384;
    return g->dim % 384 == 0 && (nwg == 1 || nwg == 2 || nwg == 4 || nwg == 8);
}

extern "C" int fv_gate_bwd_v(const fv_geom* g_, int dtype, const void* v, const void* z, int64_t ldz, int64_t z_bstride,
                             const void* dy, int64_t lddy, int64_t dy_bstride, const float* ln_w, const float* ln_b, float eps,
                             void* dz, void* e, float* ds, float* dln_w, float* dln_b, void* stream) {
    using namespace fv;
    if (int rc = check_geom(g_, "fv_gate_bwd_v")) return rc;
    FV_REQUIRE(v && z && dy && dz && e && ds, "fv_gate_bwd_v: null pointer");
    FV_REQUIRE(fv_gate_bwd_v_supported(g_, dtype), "fv_gate_bwd_v: needs bf16, a plain (outer, pool, 1) geometry and dim in "
                                                   "{384, 768, 1536, 3072}; use fv_gate_bwd otherwise");
    FV_REQUIRE(!ln_w || dln_w, "fv_gate_bwd_v: dln_w is required with LayerNorm");
    FV_REQUIRE(ldz % 4 == 0 && z_bstride % 4 == 0 && lddy % 4 == 0 && dy_bstride % 4 == 0, "fv_gate_bwd_v: strides must be multiples of 4");
    FV_REQUIRE(((uintptr_t)v % 8) == 0 && ((uintptr_t)z % 8) == 0 && ((uintptr_t)dy % 8) == 0 && ((uintptr_t)dz % 8) == 0 &&
                   ((uintptr_t)e % 8) == 0 && ((uintptr_t)ds % 16) == 0, "fv_gate_bwd_v: misaligned pointer");
    const int64_t Lmem = (int64_t)g_->outer * g_->pool;
    FV_REQUIRE(Lmem * (ldz > lddy ? ldz : lddy) < (1ll << 31), "fv_gate_bwd_v: image too large for 32-bit row offsets");
    GateBwdVArgs a;
    a.g = make_geom(g_);
    a.v = (const bf16*)v; a.z = (const bf16*)z; a.ldz = ldz; a.zbs = z_bstride;
    a.dy = (const bf16*)dy; a.lddy = lddy; a.dybs = dy_bstride;
    a.lnw = ln_w; a.lnb = ln_b; a.eps = eps;
    a.dz = (bf16*)dz; a.e = (bf16*)e; a.ds = ds; a.dlnw = dln_w; a.dlnb = dln_b;
    a.nitems = g_->batch * g_->outer;
    const int nwg = g_->dim / 384;
    void (*kern)(const GateBwdVArgs) = nullptr;
#define FV_GV(N_) (nwg == 1 ? gate_bwd_v_kernel<N_, 1> : nwg == 2 ? gate_bwd_v_kernel<N_, 2> : nwg == 4 ? gate_bwd_v_kernel<N_, 4> : gate_bwd_v_kernel<N_, 8>)
    kern = ln_w ? FV_GV(true) : FV_GV(false);
#undef FV_GV
    const size_t smem = (size_t)2 * g_->dim * 4 + 2 * GV_WARPS * sizeof(float4);
    if (smem > 48 * 1024) {
        cudaError_t er = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        FV_REQUIRE(er == cudaSuccess, "fv_gate_bwd_v: cudaFuncSetAttribute: %s", cudaGetErrorString(er));
    }
    const int ngrp = GV_WARPS / nwg;
    int64_t want = (a.nitems + ngrp - 1) / ngrp;
    const int64_t resident = (int64_t)sm_count() * 2;
    const int grid = (int)(want < resident ? want : resident);
    FV_LAUNCH_PDL((kern), grid, GV_THREADS, smem, (cudaStream_t)stream, a);
    return finish_launch("gate_bwd_v");
}
