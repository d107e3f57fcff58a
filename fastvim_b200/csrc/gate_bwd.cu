// K2b-bwd -- backward of the scan epilogue (broadcast + D skip + direction average + LayerNorm + SiLU gate).
//
// Reference: autograd through the live branch of Mamba.forward, mamba_ssm/modules/mamba_simple_faster.py:356-358,
// 412-416, 434-453 (repeat_interleave, out += D * x, flip/add//2, LayerNorm, * silu(z)); the fused-autograd form
// of the first two is selective_scan_interface.py:662-675 (dD = sum dout * conv_out, dconv_out = dout * D,
// pooled dout = dout.view(B, D, Lp, Wc).sum(-1)).
//
// Forward (csrc/gate.cu): v = (s_f[j] + s_b[j] + D_f xc_f + D_b xc_b) / 2;  y = LN(v; gamma, beta) * silu(z).
// Given dy this kernel recomputes v (both convs from the staged x rows) and the LayerNorm statistics,
// then produces in one pass over x, z, dy:
//   dz   = dy * LN(v) * silu'(z)                                   -> written next to dx in the d(xz) buffer
//   e    = dv / 2,  dv = rstd * (dxh - mean(dxh) - xhat * mean(dxh * xhat)),  dxh = dy * silu(z) * gamma
//                                                                   -> (B, L, D), consumed by the conv backward
//   ds   = sum over the tile's tokens of e (identical for both scan directions)
//                                                                   -> one fp32 plane per tile of a pooled group
//   dD_f = sum e * xc_f, dD_b = sum e * xc_b, dgamma = sum dy silu(z) xhat, dbeta = sum dy silu(z)
//                                                                   -> register accumulators, one atomicAdd per CTA
// Persistent grid; tiles and row tables as in the forward kernel (tiles.cuh); x (+3 halo rows each side),
// z and dy rows are staged with cp.async.
#include <cstdlib>

#include "common.cuh"
#include "tiles.cuh"

namespace fv {

// silu(z) and d silu(z)/dz of 4 values
__device__ __forceinline__ void silu_and_grad4(float4 zv, float4& sl, float4& ds) {
    float s;
    s = sigmoidf_(zv.x); sl.x = zv.x * s; ds.x = s * fmaf(zv.x, 1.f - s, 1.f);
    s = sigmoidf_(zv.y); sl.y = zv.y * s; ds.y = s * fmaf(zv.y, 1.f - s, 1.f);
    s = sigmoidf_(zv.z); sl.z = zv.z * s; ds.z = s * fmaf(zv.z, 1.f - s, 1.f);
    s = sigmoidf_(zv.w); sl.w = zv.w * s; ds.w = s * fmaf(zv.w, 1.f - s, 1.f);
}
__device__ __forceinline__ float4 mul4(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float hsum4(float4 a) { return (a.x + a.y) + (a.z + a.w); }
__device__ __forceinline__ float dot4(float4 a, float4 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }

// block-wide per-token reduction of per-thread float2 partials: psum[TT][nthreads] -> out[TT] (raw sums)
template <int TT>
__device__ __forceinline__ void reduce_tokens(const float2* psum, float2* out, int np, int nthreads) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = nthreads >> 5;
    for (int p = warp; p < np; p += nwarps) {
        float a = 0.f, b = 0.f;
        for (int k = lane; k < nthreads; k += 32) {
            const float2 q = psum[p * nthreads + k];
            a += q.x;
            b += q.y;
        }
        a = warp_sum(a);
        b = warp_sum(b);
        if (lane == 0) out[p] = make_float2(a, b);
    }
}

template <typename T, int TT, bool NORM, int MAXT>
__global__ void __launch_bounds__(MAXT)
gate_bwd_kernel(Geom g, int64_t ntiles, int tiles_per_img, int tiles_per_group, int tile_len, int vec16,
                const T* __restrict__ x, const T* __restrict__ z, int64_t ldxz, int64_t xzbs,
                const T* __restrict__ dy, int64_t lddy, int64_t dybs, const float* __restrict__ s,
                const float* __restrict__ cw, const float* __restrict__ cb, const float* __restrict__ Dskip,
                const float* __restrict__ lnw, const float* __restrict__ lnb, float eps, T* __restrict__ dz,
                T* __restrict__ e_out, float* __restrict__ ds_planes, float* __restrict__ dDskip,
                float* __restrict__ dlnw, float* __restrict__ dlnb) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nthreads = blockDim.x;
    const int D = g.D;
    // smem: 4 x fp32[TT][D] (v/xhat, xc_f, xc_b, dxh) | psum[TT][nthreads] | stat[TT], stat2[TT] | xs | zs | dys | tab
    float* vbuf = reinterpret_cast<float*>(smem_raw);
    float* xcf = vbuf + TT * D;
    float* xcb = xcf + TT * D;
    float* dxb = xcb + TT * D;
    float2* psum = reinterpret_cast<float2*>(dxb + TT * D);
    float2* stat = psum + TT * nthreads;
    float2* stat2 = stat + (TT + 1) / 2 * 2;
    T* xs = reinterpret_cast<T*>(stat2 + (TT + 1) / 2 * 2);
    T* zs = xs + (TT + 6) * D;
    T* dys = zs + TT * D;
    TileTab<TT>* tab = reinterpret_cast<TileTab<TT>*>(dys + TT * D);

    const int d0 = threadIdx.x * 4;
    const bool live = d0 < D;
    const int dd = live ? d0 : 0;
    const int64_t splane = (int64_t)g.B * g.Lp * D;
    const Taps tf = load_taps(cw, cb, D, 0, dd), tb = load_taps(cw, cb, D, 1, dd);
    const float4 Dfh = scale4(ld4(Dskip + dd), 0.5f), Dbh = scale4(ld4(Dskip + D + dd), 0.5f);
    float4 gam = make_float4(1.f, 1.f, 1.f, 1.f), bet = zero4();
    if (NORM) {
        gam = ld4(lnw + dd);
        if (lnb) bet = ld4(lnb + dd);
    }
    const float invD = 1.f / (float)D;
    float4 acc_dg = zero4(), acc_db = zero4(), acc_dDf = zero4(), acc_dDb = zero4();

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();  // previous tile fully consumed
        if (threadIdx.x < 32) fill_tiletab<TT>(g, tile, ntiles, tiles_per_img, tiles_per_group, tile_len, ldxz, xzbs, tab);
        __syncthreads();
        const int np = tab->np, b = tab->b;
        stage_rows(g, x + (int64_t)b * xzbs, ldxz, tab->rows, np + 6, xs, vec16 & 1);
        stage_rows(g, z + (int64_t)b * xzbs, ldxz, tab->rows + 3, np, zs, vec16 & 1);
        stage_rows(g, dy + (int64_t)b * dybs, lddy, tab->rows + 3, np, dys, (vec16 >> 1) & 1);
        cp_async_commit();
        const int j = tab->jt[0];
        const float* sp = s + ((int64_t)b * g.Lp + j) * D + dd;
        const float4 sv = scale4(ld4(sp) + ld4(sp + splane), 0.5f);
        cp_async_wait<0>();
        __syncthreads();

        // ---- P1: recompute xc_f, xc_b, v; LayerNorm partials
        for (int p = 0; p < np; ++p) {
            float4 w[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) w[k] = ld4(xs + (p + k) * D + dd);
            float4 af, ab;
            conv_both<false>(w, tf, tb, af, ab);
            float4 v = fma4(Dbh, ab, fma4(Dfh, af, sv));
            if (!live) v = zero4();
            if (live) {
                st4(vbuf + p * D + d0, v);
                st4(xcf + p * D + d0, af);
                st4(xcb + p * D + d0, ab);
            }
            psum[p * nthreads + threadIdx.x] = make_float2(hsum4(v), dot4(v, v));
        }
        if (NORM) {
            __syncthreads();
            reduce_tokens<TT>(psum, stat, np, nthreads);
            __syncthreads();
        }
        // ---- P2: dz, dgamma, dbeta, dxh; second set of partials
        for (int p = 0; p < np; ++p) {
            float4 v = ld4(vbuf + p * D + dd);
            float mean = 0.f, rstd = 1.f;
            if (NORM) {
                const float2 q = stat[p];
                mean = q.x * invD;
                rstd = rsqrtf(fmaxf(q.y * invD - mean * mean, 0.f) + eps);
            }
            const float4 xh = NORM ? scale4(make_float4(v.x - mean, v.y - mean, v.z - mean, v.w - mean), rstd) : v;
            float4 zz, dsz;
            silu_and_grad4(ld4(zs + p * D + dd), zz, dsz);
            const float4 dyv = live ? ld4(dys + p * D + dd) : zero4();
            const float4 dln = mul4(dyv, zz);
            const float4 lnv = NORM ? fma4(xh, gam, bet) : v;
            if (live) st4(dz + tab->yoff[p] + d0, mul4(mul4(dyv, lnv), dsz));
            const float4 dxh = NORM ? mul4(dln, gam) : dln;
            if (NORM) {
                acc_dg = fma4(dln, xh, acc_dg);
                acc_db = acc_db + dln;
                if (live) {
                    st4(vbuf + p * D + d0, xh);
                    st4(dxb + p * D + d0, dxh);
                }
                psum[p * nthreads + threadIdx.x] = make_float2(hsum4(dxh), dot4(dxh, xh));
            } else if (live) {
                st4(dxb + p * D + d0, dxh);
            }
        }
        if (NORM) {
            __syncthreads();
            reduce_tokens<TT>(psum, stat2, np, nthreads);
            __syncthreads();
        }
        // ---- P3: e = dv / 2, pooled ds, dD
        float4 ds_acc = zero4();
        for (int p = 0; p < np; ++p) {
            float4 dv = ld4(dxb + p * D + dd);
            if (NORM) {
                const float2 q = stat[p];
                const float mean = q.x * invD;
                const float rstd = rsqrtf(fmaxf(q.y * invD - mean * mean, 0.f) + eps);
                const float2 c = stat2[p];
                const float c1 = c.x * invD, c2 = c.y * invD;
                const float4 xh = ld4(vbuf + p * D + dd);
                dv = make_float4(rstd * (dv.x - c1 - xh.x * c2), rstd * (dv.y - c1 - xh.y * c2),
                                 rstd * (dv.z - c1 - xh.z * c2), rstd * (dv.w - c1 - xh.w * c2));
            }
            const float4 e = scale4(dv, 0.5f);
            if (live) {
                st4(e_out + ((int64_t)b * g.L + tab->rows[p + 3]) * D + d0, e);
                ds_acc = ds_acc + e;
                acc_dDf = fma4(e, ld4(xcf + p * D + d0), acc_dDf);
                acc_dDb = fma4(e, ld4(xcb + p * D + d0), acc_dDb);
            }
        }
        if (live) {
            const int rem = (int)(tile - (int64_t)b * tiles_per_img);
            const int q = rem - (rem / tiles_per_group) * tiles_per_group;
            st4(ds_planes + q * splane + ((int64_t)b * g.Lp + j) * D + d0, ds_acc);
        }
    }
    if (live) {
        const float a[16] = {acc_dDf.x, acc_dDf.y, acc_dDf.z, acc_dDf.w, acc_dDb.x, acc_dDb.y, acc_dDb.z, acc_dDb.w,
                             acc_dg.x,  acc_dg.y,  acc_dg.z,  acc_dg.w,  acc_db.x,  acc_db.y,  acc_db.z,  acc_db.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            atomicAdd(dDskip + d0 + k, a[k]);
            atomicAdd(dDskip + D + d0 + k, a[4 + k]);
            if (NORM) {
                atomicAdd(dlnw + d0 + k, a[8 + k]);
                atomicAdd(dlnb + d0 + k, a[12 + k]);
            }
        }
    }
}

int check_geom(const fv_geom* g, const char* who);

template <typename T, int TT>
static size_t gate_bwd_smem(int D, int threads) {
    return (size_t)4 * TT * D * 4 + (size_t)TT * threads * 8 + (size_t)2 * ((TT + 1) / 2 * 2) * 8 +
           (size_t)(3 * TT + 6) * D * sizeof(T) + sizeof(TileTab<TT>);
}

template <typename T, int TT>
static int launch_gate_bwd(const Geom& g, int tpg, int tile_len, const T* x, const T* z, int64_t ldxz, int64_t xzbs,
                           const T* dy, int64_t lddy, int64_t dybs, const float* s, const float* cw, const float* cb,
                           const float* Dskip, const float* lnw, const float* lnb, float eps, T* dz, T* e_out,
                           float* ds_planes, float* dDskip, float* dlnw, float* dlnb, cudaStream_t st) {
    const int threads = ((g.D / 4) + 31) / 32 * 32;
    const size_t smem = gate_bwd_smem<T, TT>(g.D, threads);
    FV_REQUIRE(smem <= 227 * 1024, "fv_gate_bwd: shared memory %zu too large (dim %d)", smem, g.D);
    const int tiles_per_img = g.Lp * tpg;
    const int64_t ntiles = (int64_t)tiles_per_img * g.B;
    void (*kern)(Geom, int64_t, int, int, int, int, const T*, const T*, int64_t, int64_t, const T*, int64_t, int64_t,
                 const float*, const float*, const float*, const float*, const float*, const float*, float, T*, T*,
                 float*, float*, float*, float*);
    const bool norm = lnw != nullptr;
    if (threads <= 128) kern = norm ? gate_bwd_kernel<T, TT, true, 128> : gate_bwd_kernel<T, TT, false, 128>;
    else if (threads <= 256) kern = norm ? gate_bwd_kernel<T, TT, true, 256> : gate_bwd_kernel<T, TT, false, 256>;
    else if (threads <= 512) kern = norm ? gate_bwd_kernel<T, TT, true, 512> : gate_bwd_kernel<T, TT, false, 512>;
    else kern = norm ? gate_bwd_kernel<T, TT, true, 1024> : gate_bwd_kernel<T, TT, false, 1024>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        FV_REQUIRE(e == cudaSuccess, "fv_gate_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    int occ = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
    FV_REQUIRE(e == cudaSuccess && occ > 0, "fv_gate_bwd: occupancy query failed (%s)", cudaGetErrorString(e));
    const int64_t resident = (int64_t)sm_count() * occ;
    dim3 grid((unsigned)(ntiles < resident ? ntiles : resident)), block(threads);
    const int vec16 = (rows_vec16<T>(g.D, x, ldxz, xzbs) && ((uintptr_t)z % 16) == 0 ? 1 : 0) |
                      (rows_vec16<T>(g.D, dy, lddy, dybs) ? 2 : 0);
    FV_LAUNCH_PDL((kern), grid, block, smem, st, g, ntiles, tiles_per_img, tpg, tile_len, vec16, x, z, ldxz, xzbs, dy, lddy, dybs, s,
                                    cw, cb, Dskip, lnw, lnb, eps, dz, e_out, ds_planes, dDskip, dlnw, dlnb);
    return finish_launch("gate_bwd");
}

template <typename T>
static int dispatch_gate_bwd(const Geom& g, int tpg, int tile_len, const T* x, const T* z, int64_t ldxz, int64_t xzbs,
                             const T* dy, int64_t lddy, int64_t dybs, const float* s, const float* cw, const float* cb,
                             const float* Dskip, const float* lnw, const float* lnb, float eps, T* dz, T* e_out,
                             float* ds_planes, float* dDskip, float* dlnw, float* dlnb, cudaStream_t st) {
#define FV_GB_ARGS g, tpg, tile_len, x, z, ldxz, xzbs, dy, lddy, dybs, s, cw, cb, Dskip, lnw, lnb, eps, dz, e_out, ds_planes, dDskip, dlnw, dlnb, st
    if (tile_len <= 2) return launch_gate_bwd<T, 2>(FV_GB_ARGS);
    if (tile_len <= 4) return launch_gate_bwd<T, 4>(FV_GB_ARGS);
    if (tile_len <= 7) return launch_gate_bwd<T, 7>(FV_GB_ARGS);
    return launch_gate_bwd<T, 8>(FV_GB_ARGS);
#undef FV_GB_ARGS
}

}  // namespace fv

// Tiling of the backward kernels over a pooled group: number of tiles (= number of ds planes).
extern "C" int fv_bwd_tiles_per_group(const fv_geom* g_, int dtype) {
    using namespace fv;
    if (!g_ || g_->inner != 1 || g_->pool <= 0) return -1;
    const int threads = ((g_->dim / 4) + 31) / 32 * 32;
    const size_t budget = 200 * 1024;
    int maxlen = 8;
    const size_t need8 = dtype == FV_F32 ? gate_bwd_smem<float, 8>(g_->dim, threads) : gate_bwd_smem<bf16, 8>(g_->dim, threads);
    if (need8 > budget) maxlen = 4;
    const size_t need4 = dtype == FV_F32 ? gate_bwd_smem<float, 4>(g_->dim, threads) : gate_bwd_smem<bf16, 4>(g_->dim, threads);
    if (maxlen == 4 && need4 > budget) maxlen = 2;
    static const char* force = getenv("FASTVIM_BWD_MAXLEN");   // tuning switch (tools/kbench.py): cap the tile length
    if (force) {
        const int f = atoi(force);
        if (f >= 1 && f < maxlen) maxlen = f;
    }
    return (g_->pool + maxlen - 1) / maxlen;
}

extern "C" int fv_gate_bwd(const fv_geom* g_, int dtype, const void* x, const void* z, int64_t ldxz, int64_t xz_bstride,
                           const void* dy, int64_t lddy, int64_t dy_bstride, const float* s, const float* conv_w,
                           const float* conv_b, const float* Dskip, const float* ln_w, const float* ln_b, float eps,
                           void* dz, void* e_out, float* ds_planes, float* dDskip, float* dln_w, float* dln_b,
                           void* stream) {
    using namespace fv;
    if (int rc = check_geom(g_, "fv_gate_bwd")) return rc;
    FV_REQUIRE(x && z && dy && s && conv_w && Dskip && dz && e_out && ds_planes && dDskip, "fv_gate_bwd: null pointer");
    FV_REQUIRE(!ln_w || (dln_w && dln_b), "fv_gate_bwd: dln_w / dln_b required with LayerNorm");
    FV_REQUIRE(g_->inner == 1, "fv_gate_bwd: channel layouts (inner > 1) are forward-only");
    FV_REQUIRE(ldxz % 4 == 0 && xz_bstride % 4 == 0 && lddy % 4 == 0 && dy_bstride % 4 == 0,
               "fv_gate_bwd: strides must be multiples of 4 elements");
    FV_REQUIRE(g_->dim <= 4096 && g_->batch <= 65535, "fv_gate_bwd: dim > 4096 or batch > 65535");
    const int tpg = fv_bwd_tiles_per_group(g_, dtype);
    const int tile_len = (g_->pool + tpg - 1) / tpg;
    Geom g = make_geom(g_);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FV_F32)
        return dispatch_gate_bwd<float>(g, tpg, tile_len, (const float*)x, (const float*)z, ldxz, xz_bstride, (const float*)dy, lddy,
                                        dy_bstride, s, conv_w, conv_b, Dskip, ln_w, ln_b, eps, (float*)dz, (float*)e_out, ds_planes,
                                        dDskip, dln_w, dln_b, st);
    if (dtype == FV_BF16)
        return dispatch_gate_bwd<bf16>(g, tpg, tile_len, (const bf16*)x, (const bf16*)z, ldxz, xz_bstride, (const bf16*)dy, lddy,
                                       dy_bstride, s, conv_w, conv_b, Dskip, ln_w, ln_b, eps, (bf16*)dz, (bf16*)e_out, ds_planes,
                                       dDskip, dln_w, dln_b, st);
    return fail("fv_gate_bwd: unsupported dtype %d", dtype);
}
