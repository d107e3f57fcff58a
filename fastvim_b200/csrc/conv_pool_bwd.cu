// K1-bwd -- backward of conv (both directions) + SiLU + mean pooling, merged with the D-skip path.
//
// Reference (paths relative to /root/reference): FastVim_MambaInnerFnNoOutProj_withoutZ.backward,
// mamba_ssm/ops/selective_scan_interface.py:662-668 (dconv_out = dout * D), :740-748 (un-pool:
// dconv_out += repeat_interleave(dconv_compressed) / num_of_col) and :751-753 (causal_conv1d_bwd of the
// un-vendored causal-conv1d package), once per direction on flipped tensors in the module
// (mamba_simple_faster.py:272-305).
//
// In original token coordinates, with e = dL/dv / 2 from the gate backward and
// du_f, du_b the gradients of the pooled conv outputs (scan backward + x_proj backward):
//   dxc_f[t] = e[t] D_f + du_f[j(t)] sf / pool          dxc_b[t] = e[t] D_b + du_b[j(t)] sf / pool
//   G_f[t]   = dxc_f[t] silu'(cf[t]),  cf[t] = b_f + sum_k w_f[k] x[t-3+k]
//   G_b[t]   = dxc_b[t] silu'(cb[t]),  cb[t] = b_b + sum_k w_b[k] x[t+3-k]
//   dx[t]    = sum_k w_f[k] G_f[t+3-k] + sum_k w_b[k] G_b[t-3+k]
//   dw_f[k]  = sum_t G_f[t] x[t-3+k],  dw_b[k] = sum_t G_b[t] x[t+3-k],  db_f = sum_t G_f[t],  db_b = sum_t G_b[t]
// One pass over x and e, one write of dx (next to dz in the d(xz) buffer).  Persistent grid; tile =
// TT <= 8 tokens of one pooled group, all channels (4 per thread); x and e rows (+3 halo rows each side)
// staged with cp.async; G_f / G_b live in thread-private shared-memory columns (no barrier between the
// two passes); weight / bias gradients accumulate in registers and leave with one atomicAdd per CTA.
#include "common.cuh"
#include "tiles.cuh"

namespace fv {

template <typename T, int TT, int MAXT>
__global__ void __launch_bounds__(MAXT)
conv_pool_bwd_kernel(Geom g, int64_t ntiles, int tiles_per_img, int tiles_per_group, int tile_len, int vec16,
                     const T* __restrict__ x, int64_t ldx, int64_t xbs, const T* __restrict__ e,
                     const T* __restrict__ du, const float* __restrict__ cw, const float* __restrict__ cb,
                     const float* __restrict__ Dskip, float scale, T* __restrict__ dx,
                     float* __restrict__ dcw, float* __restrict__ dcb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = g.D;
    constexpr int NR = TT + 6;
    // smem: Gf[NR][D] fp32 | Gb[NR][D] fp32 | xs[NR][D] T | es[NR][D] T | tab | jh[NR]
    float* Gf = reinterpret_cast<float*>(smem_raw);
    float* Gb = Gf + NR * D;
    T* xs = reinterpret_cast<T*>(Gb + NR * D);
    T* es = xs + NR * D;
    TileTab<TT>* tab = reinterpret_cast<TileTab<TT>*>(es + NR * D);
    int* jh = reinterpret_cast<int*>(tab + 1);

    const int d0 = threadIdx.x * 4;
    const bool live = d0 < D;
    const int dd = live ? d0 : 0;
    const int64_t uplane = (int64_t)g.B * g.Lp * D;
    const Taps tf = load_taps(cw, cb, D, 0, dd), tb = load_taps(cw, cb, D, 1, dd);
    const float4 Df = ld4(Dskip + dd), Db = ld4(Dskip + D + dd);
    const float pscale = scale / (float)g.pool;
    float4 awf[4], awb[4], abf = zero4(), abb = zero4();
#pragma unroll
    for (int k = 0; k < 4; ++k) awf[k] = zero4(), awb[k] = zero4();

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < 32) {
            fill_tiletab<TT>(g, tile, ntiles, tiles_per_img, tiles_per_group, tile_len, ldx, xbs, tab);
            // pooled position of every staged position (halo included)
            const int b_ = (int)(tile / tiles_per_img), rem = (int)(tile - (int64_t)b_ * tiles_per_img);
            const int jj = rem / tiles_per_group, q = rem - jj * tiles_per_group;
            const int t = jj * g.pool + q * tile_len - 3 + (int)threadIdx.x;
            if (threadIdx.x < NR) jh[threadIdx.x] = (t >= 0 && t < g.L) ? t / g.pool : 0;
        }
        __syncthreads();
        const int np = tab->np, b = tab->b;
        stage_rows(g, x + (int64_t)b * xbs, ldx, tab->rows, np + 6, xs, vec16 & 1);
        stage_rows(g, e + (int64_t)b * g.L * D, (int64_t)D, tab->rows, np + 6, es, (vec16 >> 1) & 1);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        if (!live) continue;  // (all barriers of the next iteration are reached by every thread)

        // ---- pass 1: G_f on positions [3, np+6), G_b on [0, np+3)
        const T* dub = du + (int64_t)b * g.Lp * D + d0;
        int jprev = -1;
        float4 duf = zero4(), dub4 = zero4();
        for (int i = 0; i < np + 6; ++i) {
            float4 gf = zero4(), gb = zero4();
            if (tab->rows[i] >= 0) {
                const int j = jh[i];
                if (j != jprev) {
                    duf = scale4(ld4(dub + (int64_t)j * D), pscale);
                    dub4 = scale4(ld4(dub + uplane + (int64_t)j * D), pscale);
                    jprev = j;
                }
                const float4 ev = ld4(es + i * D + d0);
                if (i >= 3) {
                    float4 c = tf.b;
#pragma unroll
                    for (int k = 0; k < 4; ++k) c = fma4(tf.w[k], ld4(xs + (i - 3 + k) * D + d0), c);
                    const float4 dxc = fma4(ev, Df, duf);
                    gf = make_float4(dxc.x * dsilu(c.x), dxc.y * dsilu(c.y), dxc.z * dsilu(c.z), dxc.w * dsilu(c.w));
                }
                if (i < np + 3) {
                    float4 c = tb.b;
#pragma unroll
                    for (int k = 0; k < 4; ++k) c = fma4(tb.w[k], ld4(xs + (i + 3 - k) * D + d0), c);
                    const float4 dxc = fma4(ev, Db, dub4);
                    gb = make_float4(dxc.x * dsilu(c.x), dxc.y * dsilu(c.y), dxc.z * dsilu(c.z), dxc.w * dsilu(c.w));
                }
            }
            st4(Gf + i * D + d0, gf);
            st4(Gb + i * D + d0, gb);
        }
        // ---- pass 2: dx and the weight / bias gradients of the owned positions [3, np+3)
        for (int p = 0; p < np; ++p) {
            const int i = p + 3;
            float4 acc = zero4();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                acc = fma4(tf.w[k], ld4(Gf + (i + 3 - k) * D + d0), acc);
                acc = fma4(tb.w[k], ld4(Gb + (i - 3 + k) * D + d0), acc);
            }
            st4(dx + tab->yoff[p] + d0, acc);
            const float4 gf = ld4(Gf + i * D + d0), gb = ld4(Gb + i * D + d0);
            abf = abf + gf;
            abb = abb + gb;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                awf[k] = fma4(gf, ld4(xs + (i - 3 + k) * D + d0), awf[k]);
                awb[k] = fma4(gb, ld4(xs + (i + 3 - k) * D + d0), awb[k]);
            }
        }
    }
    if (live) {
        // conv_w layout (2, D, 4): channel-major, tap fastest
        const float f[4][4] = {{awf[0].x, awf[1].x, awf[2].x, awf[3].x}, {awf[0].y, awf[1].y, awf[2].y, awf[3].y},
                               {awf[0].z, awf[1].z, awf[2].z, awf[3].z}, {awf[0].w, awf[1].w, awf[2].w, awf[3].w}};
        const float r[4][4] = {{awb[0].x, awb[1].x, awb[2].x, awb[3].x}, {awb[0].y, awb[1].y, awb[2].y, awb[3].y},
                               {awb[0].z, awb[1].z, awb[2].z, awb[3].z}, {awb[0].w, awb[1].w, awb[2].w, awb[3].w}};
        const float bf_[4] = {abf.x, abf.y, abf.z, abf.w}, bb_[4] = {abb.x, abb.y, abb.z, abb.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                atomicAdd(dcw + ((int64_t)(d0 + c)) * 4 + k, f[c][k]);
                atomicAdd(dcw + ((int64_t)D + d0 + c) * 4 + k, r[c][k]);
            }
            if (dcb) {
                atomicAdd(dcb + d0 + c, bf_[c]);
                atomicAdd(dcb + D + d0 + c, bb_[c]);
            }
        }
    }
}

int check_geom(const fv_geom* g, const char* who);

template <typename T, int TT>
static size_t conv_bwd_smem(int D) {
    return (size_t)2 * (TT + 6) * D * 4 + (size_t)2 * (TT + 6) * D * sizeof(T) + sizeof(TileTab<TT>) + (TT + 6) * 4;
}

template <typename T, int TT>
static int launch_conv_bwd(const Geom& g, int tpg, int tile_len, const T* x, int64_t ldx, int64_t xbs, const T* e,
                           const T* du, const float* cw, const float* cb, const float* Dskip, float scale, T* dx,
                           float* dcw, float* dcb, cudaStream_t st) {
    const int threads = ((g.D / 4) + 31) / 32 * 32;
    const size_t smem = conv_bwd_smem<T, TT>(g.D);
    FV_REQUIRE(smem <= 227 * 1024, "fv_conv_pool_bwd: shared memory %zu too large (dim %d)", smem, g.D);
    const int tiles_per_img = g.Lp * tpg;
    const int64_t ntiles = (int64_t)tiles_per_img * g.B;
    void (*kern)(Geom, int64_t, int, int, int, int, const T*, int64_t, int64_t, const T*, const T*, const float*,
                 const float*, const float*, float, T*, float*, float*);
    if (threads <= 128) kern = conv_pool_bwd_kernel<T, TT, 128>;
    else if (threads <= 256) kern = conv_pool_bwd_kernel<T, TT, 256>;
    else if (threads <= 512) kern = conv_pool_bwd_kernel<T, TT, 512>;
    else kern = conv_pool_bwd_kernel<T, TT, 1024>;
    if (smem > 48 * 1024) {
        cudaError_t er = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        FV_REQUIRE(er == cudaSuccess, "fv_conv_pool_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(er));
    }
    int occ = 0;
    cudaError_t er = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
    FV_REQUIRE(er == cudaSuccess && occ > 0, "fv_conv_pool_bwd: occupancy query failed (%s)", cudaGetErrorString(er));
    const int64_t resident = (int64_t)sm_count() * occ;
    dim3 grid((unsigned)(ntiles < resident ? ntiles : resident)), block(threads);
    const int vec16 = (rows_vec16<T>(g.D, x, ldx, xbs) ? 1 : 0) | (rows_vec16<T>(g.D, e, g.D, (int64_t)g.L * g.D) ? 2 : 0);
    kern<<<grid, block, smem, st>>>(g, ntiles, tiles_per_img, tpg, tile_len, vec16, x, ldx, xbs, e, du, cw, cb, Dskip,
                                    scale, dx, dcw, dcb);
    return finish_launch("conv_pool_bwd");
}

}  // namespace fv

extern "C" int fv_conv_pool_bwd(const fv_geom* g_, int dtype, const void* x, int64_t ldx, int64_t x_bstride,
                                const void* e, const void* du, const float* conv_w, const float* conv_b,
                                const float* Dskip, float scale, int pool_mode, void* dx, float* dconv_w,
                                float* dconv_b, void* stream) {
    using namespace fv;
    if (int rc = check_geom(g_, "fv_conv_pool_bwd")) return rc;
    FV_REQUIRE(x && e && du && conv_w && Dskip && dx && dconv_w, "fv_conv_pool_bwd: null pointer");
    FV_REQUIRE(pool_mode == FV_POOL_MEAN, "fv_conv_pool_bwd: only mean pooling has a backward (as in the reference's fused path)");
    FV_REQUIRE(g_->inner == 1, "fv_conv_pool_bwd: channel layouts (inner > 1) are forward-only");
    FV_REQUIRE(ldx % 4 == 0 && x_bstride % 4 == 0, "fv_conv_pool_bwd: strides must be multiples of 4 elements");
    FV_REQUIRE(g_->dim <= 4096 && g_->batch <= 65535, "fv_conv_pool_bwd: dim > 4096 or batch > 65535");
    Geom g = make_geom(g_);
    const size_t budget = 200 * 1024;
    int maxlen = 8;
    if ((dtype == FV_F32 ? conv_bwd_smem<float, 8>(g.D) : conv_bwd_smem<bf16, 8>(g.D)) > budget) maxlen = 4;
    if (maxlen == 4 && (dtype == FV_F32 ? conv_bwd_smem<float, 4>(g.D) : conv_bwd_smem<bf16, 4>(g.D)) > budget) maxlen = 2;
    const int tpg = (g.pool + maxlen - 1) / maxlen;
    const int tile_len = (g.pool + tpg - 1) / tpg;
    cudaStream_t st = (cudaStream_t)stream;
#define FV_CB(T_)                                                                                                      \
    do {                                                                                                               \
        if (tile_len <= 2)                                                                                             \
            return launch_conv_bwd<T_, 2>(g, tpg, tile_len, (const T_*)x, ldx, x_bstride, (const T_*)e, (const T_*)du,  \
                                          conv_w, conv_b, Dskip, scale, (T_*)dx, dconv_w, dconv_b, st);                \
        if (tile_len <= 4)                                                                                             \
            return launch_conv_bwd<T_, 4>(g, tpg, tile_len, (const T_*)x, ldx, x_bstride, (const T_*)e, (const T_*)du,  \
                                          conv_w, conv_b, Dskip, scale, (T_*)dx, dconv_w, dconv_b, st);                \
        if (tile_len <= 7)                                                                                             \
            return launch_conv_bwd<T_, 7>(g, tpg, tile_len, (const T_*)x, ldx, x_bstride, (const T_*)e, (const T_*)du,  \
                                          conv_w, conv_b, Dskip, scale, (T_*)dx, dconv_w, dconv_b, st);                \
        return launch_conv_bwd<T_, 8>(g, tpg, tile_len, (const T_*)x, ldx, x_bstride, (const T_*)e, (const T_*)du,      \
                                      conv_w, conv_b, Dskip, scale, (T_*)dx, dconv_w, dconv_b, st);                    \
    } while (0)
    if (dtype == FV_F32) FV_CB(float);
    if (dtype == FV_BF16) FV_CB(bf16);
#undef FV_CB
    return fail("fv_conv_pool_bwd: unsupported dtype %d", dtype);
}
