// K1 -- depthwise causal conv1d (both scan directions) + SiLU + pooling, one pass over x.
//
// What it replaces in the reference (paths relative to /root/reference):
//   x.flip([-1])                                   mamba_ssm/modules/mamba_simple_faster.py:272
//   causal_conv1d_fn(x, conv1d) / (x_flip, conv1d_b)  :274-285   (causal-conv1d 1.1.3, un-vendored)
//   x.reshape(pre_x_shape).mean(3) [* scaling] / .max(3)  :287-305
// i.e. one flip copy, two conv kernels and two reductions (>= 8 full-resolution passes)
// become one read of x.  In original token order the b-direction is the anti-causal conv
// out_b[t] = silu(b_b + sum_k w_b[k] x[t+3-k]), so no flip is needed.
//
// Mapping: token-major x (B, L, D); one thread owns 4 consecutive channels of one pooled
// position j and slides a 7-row register window over the `pool` tokens of that position
// (+3 halo rows each side, shared through L1/L2 with the neighbouring positions).  A warp
// therefore reads 32 x 4 channels = 256 B (bf16) / 512 B (fp32) contiguous per row: fully
// coalesced, independent of the token permutation (rotated layers only change row numbers).
// HBM-bound by design: algorithmic bytes = B*L*D*s read + 2*B*Lp*D*s written.
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"

namespace fv {

int sm_count();

template <typename T>
__device__ __forceinline__ float4 load_row4(const Geom& g, const T* xb, int64_t ldx, int d0, int t) {
    if (t < 0 || t >= g.L) return zero4();
    return ld4(xb + seq_to_row(g, t) * ldx + d0);
}

template <typename T, bool MAXPOOL, bool INNER1>
__global__ void __launch_bounds__(256)
conv_pool_fwd_kernel(Geom g, const T* __restrict__ x, int64_t ldx, int64_t xbs,
                     const float* __restrict__ cw, const float* __restrict__ cb, float scale,
                     T* __restrict__ u) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    constexpr bool FAST = is_fast<T>::value;
    const int nvec = g.D >> 2;
    int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= (int64_t)g.B * g.Lp * nvec) return;
    const int v = (int)(item % nvec);
    const int64_t bj = item / nvec;
    const int j = (int)(bj % g.Lp), b = (int)(bj / g.Lp);
    const int d0 = v * 4;
    const T* xb = x + (int64_t)b * xbs;
    const Taps tf = load_taps(cw, cb, g.D, 0, d0), tb = load_taps(cw, cb, g.D, 1, d0);

    float4 accf = MAXPOOL ? make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY) : zero4();
    float4 accb = accf;
    float4 w[7];
    if (INNER1) {
        constexpr int PF = 4;  // rows prefetched ahead of the window
        const int t0 = j * g.pool;
#pragma unroll
        for (int i = 0; i < 6; ++i) w[i + 1] = load_row4(g, xb, ldx, d0, t0 - 3 + i);
        float4 ring[PF];
#pragma unroll
        for (int i = 0; i < PF; ++i) ring[i] = load_row4(g, xb, ldx, d0, t0 + 3 + i);
        for (int p0 = 0; p0 < g.pool; p0 += PF) {
#pragma unroll
            for (int i = 0; i < PF; ++i) {
                const int p = p0 + i;
                if (p < g.pool) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) w[k] = w[k + 1];
                    w[6] = ring[i];
                    ring[i] = (p + PF < g.pool) ? load_row4(g, xb, ldx, d0, t0 + 3 + p + PF) : zero4();
                    float4 xf, xr;
                    conv_both<FAST>(w, tf, tb, xf, xr);
                    accf = MAXPOOL ? max4(accf, xf) : accf + xf;
                    accb = MAXPOOL ? max4(accb, xr) : accb + xr;
                }
            }
        }
    } else {
        // channel layouts (inner > 1).  Channel-First FastChannelVim keeps the sequence in memory order (row = t):
        // then the 7-row window needs no index arithmetic at all; other stride sets take the generic map.
        const bool natural = g.si == 1 && g.sp == g.inner && g.so == (int64_t)g.pool * g.inner;
        const int o = j / g.inner, i = j - o * g.inner;
        const T* xd = xb + d0;
        for (int p = 0; p < g.pool; ++p) {
            const int c = (o * g.pool + p) * g.inner + i;   // = pooled_to_seq(g, j, p)
            if (natural) {
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    const int t = c - 3 + k;
                    w[k] = (t >= 0 && t < g.L) ? ld4(xd + (int64_t)t * ldx) : zero4();
                }
            } else {
#pragma unroll
                for (int k = 0; k < 7; ++k) w[k] = load_row4(g, xb, ldx, d0, c - 3 + k);
            }
            float4 xf, xr;
            conv_both<FAST>(w, tf, tb, xf, xr);
            accf = MAXPOOL ? max4(accf, xf) : accf + xf;
            accb = MAXPOOL ? max4(accb, xr) : accb + xr;
        }
    }
    if (!MAXPOOL) {
        const float m = scale / (float)g.pool;
        accf = scale4(accf, m);
        accb = scale4(accb, m);
    }
    const int64_t plane = (int64_t)g.B * g.Lp * g.D;
    T* uo = u + ((int64_t)b * g.Lp + j) * g.D + d0;
    st4(uo, accf);
    st4(uo + plane, accb);
}

// v2 for the plain (outer, pool, 1) geometry: CTA = one pooled position (b, j), all channels
// (4 per thread).  The pool+6 token rows are staged with 16-byte cp.async (zero-filled outside
// the sequence), in double-buffered chunks of TP tokens when the pooled group is long (2048^2:
// pool = 128), so every global load of a chunk is in flight at once.
template <typename T, bool MAXPOOL, int MAXT>
__global__ void __launch_bounds__(MAXT)
conv_pool_staged_kernel(Geom g, int TP, int nbuf, int vec16, const T* __restrict__ x, int64_t ldx, int64_t xbs,
                        const float* __restrict__ cw, const float* __restrict__ cb, float scale,
                        T* __restrict__ u, const float* __restrict__ Dskip, T* __restrict__ wout) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    constexpr bool FAST = is_fast<T>::value;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* xs = reinterpret_cast<T*>(smem_raw);
    constexpr int G = 7;
    const int D = g.D;
    const int bufsz = (TP + 6) * D;
    int* rowtab = reinterpret_cast<int*>(xs + nbuf * bufsz);  // [g.pool + 6]
    const int j = blockIdx.x, b = blockIdx.y;
    const int d0 = threadIdx.x * 4;
    const bool live = d0 < g.D;
    const int dd = live ? d0 : 0;
    const T* xb = x + (int64_t)b * xbs;
    const int tbase = j * g.pool;
    const int nchunk = (g.pool + TP - 1) / TP;

    fill_rowtab(g, tbase - 3, g.pool + 6, rowtab);
    __syncthreads();
    stage_rows(g, xb, ldx, rowtab, min(TP, g.pool) + 6, xs, vec16 != 0);
    cp_async_commit();
    constexpr float PRE = FAST ? 0.5f : 1.f;
    const Taps tf = load_taps(cw, cb, g.D, 0, dd, PRE), tb = load_taps(cw, cb, g.D, 1, dd, PRE);
    float4 accf = MAXPOOL ? make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY) : zero4();
    float4 accb = accf;
    // optional second output: the D-skip term w = (D_f xc_f + D_b xc_b) / 2 of every token, for fv_gate_w_fwd
    const bool want_w = wout != nullptr && live;
    const float4 Df = want_w ? scale4(ld4(Dskip + dd), 0.5f) : zero4(), Db = want_w ? scale4(ld4(Dskip + g.D + dd), 0.5f) : zero4();
    T* wimg = want_w ? wout + (int64_t)b * g.L * g.D + dd : nullptr;
    for (int c = 0; c < nchunk; ++c) {
        const int p_lo = c * TP, np = min(TP, g.pool - p_lo);
        if (c + 1 < nchunk) {
            const int q_lo = p_lo + TP, nq = min(TP, g.pool - q_lo);
            stage_rows(g, xb, ldx, rowtab + q_lo, nq + 6, xs + ((c + 1) % nbuf) * bufsz, vec16 != 0);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const T* cur = xs + (c % nbuf) * bufsz + dd;
        // groups of G tokens, fully unrolled: G + 6 smem rows per group, window in registers
        for (int p0 = 0; p0 < np; p0 += G) {
            float4 r[G + 6];
            const T* xp = cur + p0 * D;
#pragma unroll
            for (int k = 0; k < 6; ++k) r[k] = ld4(xp + k * D);
#pragma unroll
            for (int i = 0; i < G; ++i) {
                if (p0 + i < np) {
                    r[i + 6] = ld4(xp + (i + 6) * D);
                    float4 xf, xr;
                    conv_both_pre<FAST>(r[i], r[i + 1], r[i + 2], r[i + 3], r[i + 4], r[i + 5], r[i + 6], tf, tb, xf, xr);
                    accf = MAXPOOL ? max4(accf, xf) : accf + xf;
                    accb = MAXPOOL ? max4(accb, xr) : accb + xr;
                    if (want_w)
                        st4(wimg + (int64_t)rowtab[p_lo + p0 + i + 3] * g.D,
                            make_float4(fmaf(Db.x, xr.x, Df.x * xf.x), fmaf(Db.y, xr.y, Df.y * xf.y), fmaf(Db.z, xr.z, Df.z * xf.z),
                                        fmaf(Db.w, xr.w, Df.w * xf.w)));
                }
            }
        }
        __syncthreads();  // all reads of this buffer done before it is refilled two chunks later
    }
    if (!live) return;
    if (!MAXPOOL) {
        const float m = scale / (float)g.pool;
        accf = scale4(accf, m);
        accb = scale4(accb, m);
    }
    const int64_t plane = (int64_t)g.B * g.Lp * g.D;
    T* uo = u + ((int64_t)b * g.Lp + j) * g.D + d0;
    st4(uo, accf);
    st4(uo + plane, accb);
}

// v3 for long pooled groups and few images (2048^2: one image, pool = 128): the staged kernel above would run only
// Lp = 128 CTAs, each walking 128 tokens serially (23 us for 12.8 MB).  Here a thread-block CLUSTER of NS CTAs owns
// one pooled position: CTA `rank` convolves tokens [rank*np, (rank+1)*np) of the group (its own 3-token halos staged
// with cp.async like above), the partial sums / maxima meet in rank 0 through distributed shared memory
// (cluster.map_shared_rank), and rank 0 scales and stores.  NS x more CTAs, no workspace, no atomics.
template <typename T, bool MAXPOOL, int MAXT>
__global__ void __launch_bounds__(MAXT)
conv_pool_cluster_kernel(Geom g, int NS, int np_seg, int vec16, const T* __restrict__ x, int64_t ldx, int64_t xbs,
                         const float* __restrict__ cw, const float* __restrict__ cb, float scale,
                         T* __restrict__ u, const float* __restrict__ Dskip, T* __restrict__ wout) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    constexpr bool FAST = is_fast<T>::value;
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int G = 7;
    const int D = g.D;
    float4* part = reinterpret_cast<float4*>(smem_raw);                     // [2][blockDim.x] partial results of this CTA
    T* xs = reinterpret_cast<T*>(part + 2 * blockDim.x);                    // [np_seg + 6][D]
    int* rowtab = reinterpret_cast<int*>(xs + (size_t)(np_seg + 6) * D);    // [np_seg + 6]
    const int rank = (int)cluster.block_rank();
    const int j = blockIdx.x / NS, b = blockIdx.y;
    const int d0 = threadIdx.x * 4;
    const bool live = d0 < g.D;
    const int dd = live ? d0 : 0;
    const T* xb = x + (int64_t)b * xbs;
    const int p_lo = rank * np_seg, np = max(0, min(np_seg, g.pool - p_lo));
    const int tbase = j * g.pool + p_lo;

    fill_rowtab(g, tbase - 3, np + 6, rowtab);
    __syncthreads();
    if (np > 0) stage_rows(g, xb, ldx, rowtab, np + 6, xs, vec16 != 0);
    cp_async_commit();
    constexpr float PRE = FAST ? 0.5f : 1.f;
    const Taps tf = load_taps(cw, cb, g.D, 0, dd, PRE), tb = load_taps(cw, cb, g.D, 1, dd, PRE);
    float4 accf = MAXPOOL ? make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY) : zero4();
    float4 accb = accf;
    const bool want_w = wout != nullptr && live;
    const float4 Df = want_w ? scale4(ld4(Dskip + dd), 0.5f) : zero4(), Db = want_w ? scale4(ld4(Dskip + g.D + dd), 0.5f) : zero4();
    T* wimg = want_w ? wout + (int64_t)b * g.L * g.D + dd : nullptr;
    cp_async_wait<0>();
    __syncthreads();
    for (int p0 = 0; p0 < np; p0 += G) {
        float4 r[G + 6];
        const T* xp = xs + dd + p0 * D;
#pragma unroll
        for (int k = 0; k < 6; ++k) r[k] = ld4(xp + k * D);
#pragma unroll
        for (int i = 0; i < G; ++i) {
            if (p0 + i < np) {
                r[i + 6] = ld4(xp + (i + 6) * D);
                float4 xf, xr;
                conv_both_pre<FAST>(r[i], r[i + 1], r[i + 2], r[i + 3], r[i + 4], r[i + 5], r[i + 6], tf, tb, xf, xr);
                accf = MAXPOOL ? max4(accf, xf) : accf + xf;
                accb = MAXPOOL ? max4(accb, xr) : accb + xr;
                if (want_w)
                    st4(wimg + (int64_t)rowtab[p0 + i + 3] * g.D,
                        make_float4(fmaf(Db.x, xr.x, Df.x * xf.x), fmaf(Db.y, xr.y, Df.y * xf.y), fmaf(Db.z, xr.z, Df.z * xf.z),
                                    fmaf(Db.w, xr.w, Df.w * xf.w)));
            }
        }
    }
    part[threadIdx.x] = accf;
    part[blockDim.x + threadIdx.x] = accb;
    cluster.sync();  // every CTA's partials are visible cluster-wide
    if (rank == 0 && live) {
        for (int r = 1; r < NS; ++r) {
            const float4* rp = cluster.map_shared_rank(part, r);
            const float4 pf = rp[threadIdx.x], pb = rp[blockDim.x + threadIdx.x];
            accf = MAXPOOL ? max4(accf, pf) : accf + pf;
            accb = MAXPOOL ? max4(accb, pb) : accb + pb;
        }
        if (!MAXPOOL) {
            const float m = scale / (float)g.pool;
            accf = scale4(accf, m);
            accb = scale4(accb, m);
        }
        const int64_t plane = (int64_t)g.B * g.Lp * g.D;
        T* uo = u + ((int64_t)b * g.Lp + j) * g.D + d0;
        st4(uo, accf);
        st4(uo + plane, accb);
    }
    cluster.sync();  // keep every CTA's shared memory alive until rank 0 has read it
}

template <typename T>
static int launch_conv_pool_cluster(const Geom& g, int NS, const T* x, int64_t ldx, int64_t xbs, const float* cw,
                                    const float* cb, float scale, int pool_mode, T* u, cudaStream_t st,
                                    const float* Dskip = nullptr, T* wout = nullptr) {
    const int threads = ((g.D / 4) + 31) / 32 * 32;
    const int np_seg = (g.pool + NS - 1) / NS;
    const size_t smem = (size_t)2 * threads * sizeof(float4) + (size_t)(np_seg + 6) * g.D * sizeof(T) + (size_t)(np_seg + 6) * 4;
    void (*kern)(Geom, int, int, int, const T*, int64_t, int64_t, const float*, const float*, float, T*, const float*, T*);
    const bool mx = pool_mode == FV_POOL_MAX;
    if (threads <= 128) kern = mx ? conv_pool_cluster_kernel<T, true, 128> : conv_pool_cluster_kernel<T, false, 128>;
    else if (threads <= 256) kern = mx ? conv_pool_cluster_kernel<T, true, 256> : conv_pool_cluster_kernel<T, false, 256>;
    else if (threads <= 512) kern = mx ? conv_pool_cluster_kernel<T, true, 512> : conv_pool_cluster_kernel<T, false, 512>;
    else kern = mx ? conv_pool_cluster_kernel<T, true, 1024> : conv_pool_cluster_kernel<T, false, 1024>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        FV_REQUIRE(e == cudaSuccess, "fv_conv_pool_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(g.Lp * NS), (unsigned)g.B);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    pdl_attr(&attr[1]);
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)NS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    const int vec16 = (int)rows_vec16<T>(g.D, x, ldx, xbs);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, g, NS, np_seg, vec16, x, ldx, xbs, cw, cb, scale, u, Dskip, wout);
    FV_REQUIRE(e == cudaSuccess, "fv_conv_pool_fwd: cluster launch failed: %s", cudaGetErrorString(e));
    return finish_launch("conv_pool_fwd");
}

template <typename T>
static int launch_conv_pool_staged(const Geom& g, const T* x, int64_t ldx, int64_t xbs, const float* cw,
                                   const float* cb, float scale, int pool_mode, T* u, cudaStream_t st,
                                   const float* Dskip = nullptr, T* wout = nullptr) {
    const int threads = ((g.D / 4) + 31) / 32 * 32;
    int TP = g.pool < 32 ? g.pool : 32;
    while (TP > 4 && (size_t)2 * (TP + 6) * g.D * sizeof(T) > 96 * 1024) TP /= 2;
    const int nchunk = (g.pool + TP - 1) / TP;
    const int nbuf = nchunk > 1 ? 2 : 1;
    const size_t smem = (size_t)nbuf * (TP + 6) * g.D * sizeof(T) + (size_t)(g.pool + 6) * 4;
    FV_REQUIRE(g.Lp <= 2147483647 && g.B <= 65535, "fv_conv_pool_fwd: batch > 65535");
    dim3 grid(g.Lp, g.B), block(threads);
    void (*kern)(Geom, int, int, int, const T*, int64_t, int64_t, const float*, const float*, float, T*, const float*, T*);
    const bool mx = pool_mode == FV_POOL_MAX;
    if (threads <= 128) kern = mx ? conv_pool_staged_kernel<T, true, 128> : conv_pool_staged_kernel<T, false, 128>;
    else if (threads <= 256) kern = mx ? conv_pool_staged_kernel<T, true, 256> : conv_pool_staged_kernel<T, false, 256>;
    else if (threads <= 512) kern = mx ? conv_pool_staged_kernel<T, true, 512> : conv_pool_staged_kernel<T, false, 512>;
    else kern = mx ? conv_pool_staged_kernel<T, true, 1024> : conv_pool_staged_kernel<T, false, 1024>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        FV_REQUIRE(e == cudaSuccess, "fv_conv_pool_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    FV_LAUNCH_PDL((kern), grid, block, smem, st, g, TP, nbuf, (int)rows_vec16<T>(g.D, x, ldx, xbs), x, ldx, xbs, cw, cb, scale, u,
                  Dskip, wout);
    return finish_launch("conv_pool_fwd");
}

template <typename T>
static int launch_conv_pool(const Geom& g, const T* x, int64_t ldx, int64_t xbs, const float* cw,
                            const float* cb, float scale, int pool_mode, T* u, cudaStream_t st,
                            const float* Dskip = nullptr, T* wout = nullptr) {
    if (g.inner == 1 && g.D <= 4096) {
        // long pooled groups, too few (image, pooled position) pairs to fill the GPU: one cluster per position
        if (g.pool >= 32 && (int64_t)g.Lp * g.B < 2 * sm_count() && g.B <= 65535) {
            // cluster size: enough CTAs to fill the GPU, but one wave (5 CTAs / SM by registers).  2048^2 (128 pooled rows of
            // 128 tokens), measured: 8 CTAs per row (1024 CTAs, 1.4 waves) 19.8 us, 4 (512) 14.4 us, 2 (256) 18.3 us
            int NS = 8;
            while (NS > 2 && (g.pool / NS < 8 || (int64_t)g.Lp * g.B * NS > 5ll * sm_count())) NS >>= 1;
            if (const char* e = getenv("FASTVIM_CONV_CLUSTER_NS")) NS = atoi(e) == 4 ? 4 : (atoi(e) == 2 ? 2 : 8);   // A/B timing
            return launch_conv_pool_cluster<T>(g, NS, x, ldx, xbs, cw, cb, scale, pool_mode, u, st, Dskip, wout);
        }
        return launch_conv_pool_staged<T>(g, x, ldx, xbs, cw, cb, scale, pool_mode, u, st, Dskip, wout);
    }
    if (wout) return fail("fv_conv_pool_w_fwd: the D-skip output needs the plain geometry here (inner == 1, dim <= 4096)");
    const int64_t items = (int64_t)g.B * g.Lp * (g.D / 4);
    const int threads = 256;
    const int64_t blocks = (items + threads - 1) / threads;
    FV_REQUIRE(blocks < (1ll << 31), "conv_pool: grid too large");
    dim3 grid((unsigned)blocks), block(threads);
    const bool in1 = g.inner == 1;
    if (pool_mode == FV_POOL_MAX) {
        if (in1) FV_LAUNCH_PDL((conv_pool_fwd_kernel<T, true, true>), grid, block, 0, st, g, x, ldx, xbs, cw, cb, scale, u);
        else FV_LAUNCH_PDL((conv_pool_fwd_kernel<T, true, false>), grid, block, 0, st, g, x, ldx, xbs, cw, cb, scale, u);
    } else {
        if (in1) FV_LAUNCH_PDL((conv_pool_fwd_kernel<T, false, true>), grid, block, 0, st, g, x, ldx, xbs, cw, cb, scale, u);
        else FV_LAUNCH_PDL((conv_pool_fwd_kernel<T, false, false>), grid, block, 0, st, g, x, ldx, xbs, cw, cb, scale, u);
    }
    return finish_launch("conv_pool_fwd");
}

int check_geom(const fv_geom* g, const char* who);

// plain geometry (inner == 1), bf16: the staged / cluster kernels with the D-skip output (fv_conv_pool_w_fwd, conv_gate_w.cu)
int conv_pool_plain_w(const Geom& g, const bf16* x, int64_t ldx, int64_t xbs, const float* cw, const float* cb, float scale,
                      int pool_mode, bf16* u, const float* Dskip, bf16* wout, cudaStream_t st) {
    return launch_conv_pool<bf16>(g, x, ldx, xbs, cw, cb, scale, pool_mode, u, st, Dskip, wout);
}

}  // namespace fv

extern "C" int fv_conv_pool_fwd(const fv_geom* g_, int dtype, const void* x, int64_t ldx,
                                int64_t x_bstride, const float* conv_w, const float* conv_b,
                                float scale, int pool_mode, void* u_out, void* stream) {
    using namespace fv;
    if (int rc = check_geom(g_, "fv_conv_pool_fwd")) return rc;
    FV_REQUIRE(x && conv_w && u_out, "fv_conv_pool_fwd: null pointer");
    FV_REQUIRE(ldx % 4 == 0 && x_bstride % 4 == 0, "fv_conv_pool_fwd: ldx/x_bstride must be multiples of 4 elements");
    FV_REQUIRE(pool_mode == FV_POOL_MEAN || pool_mode == FV_POOL_MAX, "fv_conv_pool_fwd: bad pool_mode %d", pool_mode);
    Geom g = make_geom(g_);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FV_F32)
        return launch_conv_pool<float>(g, (const float*)x, ldx, x_bstride, conv_w, conv_b, scale, pool_mode, (float*)u_out, st);
    if (dtype == FV_BF16)
        return launch_conv_pool<bf16>(g, (const bf16*)x, ldx, x_bstride, conv_w, conv_b, scale, pool_mode, (bf16*)u_out, st);
    return fail("fv_conv_pool_fwd: unsupported dtype %d", dtype);
}
