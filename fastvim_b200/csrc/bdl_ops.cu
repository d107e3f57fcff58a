// Operator-API helpers on the reference's (batch, dim, seqlen) layout (seqlen contiguous).
//
// The reference's fused autograd functions -- mamba_inner_fn_no_out_proj, ..._withoutZ and
// FastVim_mamba_inner_fn_no_out_proj_withoutZ (mamba_ssm/ops/selective_scan_interface.py:1652-1753,
// forward bodies :208-330, :452-605) -- take (B, D, L) tensors and call, in order:
//   causal_conv1d_cuda.causal_conv1d_fwd(x, w, bias, None, True)      :496-498   (third-party causal-conv1d 1.1.3)
//   conv1d_out.reshape(pre_x_shape).mean(3) [* scaling_factor]        :503-508
//   selective_scan_cuda.fwd(...)                                      :556-566   -> fv_selective_scan_fwd
//   out.repeat_interleave(num_of_col, 2); out += D.unsqueeze(-1) * conv1d_out   :570-571
// These three kernels serve that API one-to-one in the same layout (fastvim_b200.interface).  The model's own
// path never uses them: fastvim_b200.mixer runs token-major through K1/K2a/K2b or the fused fv_block_fwd.
#include "common.cuh"

namespace fv {

// out[b,d,l] = act(bias[d] + sum_k w[d,k] * x[b,d,l-3+k]); one thread per (b, d, run of E timesteps)
template <typename T, int E>
__global__ void __launch_bounds__(256)
causal_conv1d_bdl_kernel(int64_t nrows, int dim, int64_t L, int64_t chunks, const T* __restrict__ x, int64_t xbs,
                         int64_t xds, const float* __restrict__ w, const float* __restrict__ bias, int silu,
                         T* __restrict__ out) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= nrows * chunks) return;
    const int64_t row = item / chunks, l0 = (item - row * chunks) * E;
    const int64_t b = row / dim;
    const int d = (int)(row - b * dim);
    const T* xr = x + b * xbs + (int64_t)d * xds;
    const float w0 = w[d * 4], w1 = w[d * 4 + 1], w2 = w[d * 4 + 2], w3 = w[d * 4 + 3];
    const float bs = bias ? bias[d] : 0.f;
    float v[E + 3];
#pragma unroll
    for (int i = 0; i < E + 3; ++i) {
        const int64_t l = l0 - 3 + i;
        v[i] = (l >= 0 && l < L) ? ld1(xr + l) : 0.f;
    }
    T* orow = out + row * L;
#pragma unroll
    for (int i = 0; i < E; ++i) {
        if (l0 + i < L) {
            float a = fmaf(w3, v[i + 3], fmaf(w2, v[i + 2], fmaf(w1, v[i + 1], fmaf(w0, v[i], bs))));
            if (silu) a = silu_exact(a);
            st1(orow + l0 + i, a);
        }
    }
}

// (B, D, outer*pool*inner) -> (B, D, outer*inner): mean (times scale) or max over the pool axis
template <typename T>
__global__ void __launch_bounds__(256)
pool_bdl_kernel(int64_t nrows, int outer, int pool, int inner, const T* __restrict__ x, int is_max, float scale,
                T* __restrict__ out) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    const int Lp = outer * inner;
    const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= nrows * Lp) return;
    const int64_t row = item / Lp;
    const int j = (int)(item - row * Lp), o = j / inner, i = j - o * inner;
    const T* p = x + row * (int64_t)outer * pool * inner + ((int64_t)o * pool) * inner + i;
    float acc = is_max ? -INFINITY : 0.f;
    for (int q = 0; q < pool; ++q) {
        const float v = ld1(p + (int64_t)q * inner);
        acc = is_max ? fmaxf(acc, v) : acc + v;
    }
    st1(out + item, is_max ? acc : acc * scale);
}

// out[b,d,t] = s[b,d,pool_index(t)] + D[d] * xc[b,d,t]
template <typename T>
__global__ void __launch_bounds__(256)
bcast_skip_bdl_kernel(int64_t nrows, int dim, int outer, int pool, int inner, const T* __restrict__ s,
                      const T* __restrict__ xc, const float* __restrict__ Dskip, T* __restrict__ out) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    const int64_t L = (int64_t)outer * pool * inner;
    const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= nrows * L) return;
    const int64_t row = item / L;
    const int t = (int)(item - row * L);
    const int q = t / inner, i = t - q * inner, j = (q / pool) * inner + i;
    const int d = (int)(row % dim);
    float v = ld1(s + row * (int64_t)outer * inner + j);
    if (Dskip) v = fmaf(Dskip[d], ld1(xc + item), v);
    st1(out + item, v);
}

// Backward of the depthwise causal conv (+SiLU) on (B, D, L): replaces causal_conv1d_cuda.causal_conv1d_bwd
// (call site selective_scan_interface.py:751-753; third-party causal-conv1d 1.1.3).  With pre[t] = bias + sum_k w[k] x[t-3+k]
// and g[t] = dout[t] * silu'(pre[t]):   dx[s] = sum_k w[k] g[s+3-k],   dw[k] += sum_t g[t] x[t-3+k],   db += sum_t g[t].
// One warp per (b, d) row, E consecutive steps per lane per round: a lane reads x[l0-3 .. l0+E+3) and dout[l0 .. l0+E+3),
// forms g on [l0, l0+E+3) and its E outputs; dw / db partials stay in registers over the row, then one warp reduction
// and five atomics per row (caller zero-fills dw (dim, 4) and db (dim)).
template <typename T, int E>
__global__ void __launch_bounds__(128)
causal_conv1d_bdl_bwd_kernel(int64_t nrows, int dim, int64_t L, const T* __restrict__ x, int64_t xbs, int64_t xds,
                             const float* __restrict__ w, const float* __restrict__ bias, int silu,
                             const T* __restrict__ dout, T* __restrict__ dx, int64_t dxbs, int64_t dxds,
                             float* __restrict__ dw, float* __restrict__ db) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const int64_t b = row / dim;
    const int d = (int)(row - b * dim);
    const T* xr = x + b * xbs + (int64_t)d * xds;
    const T* gr = dout + row * L;
    T* dxr = dx + b * dxbs + (int64_t)d * dxds;
    const float w0 = w[d * 4], w1 = w[d * 4 + 1], w2 = w[d * 4 + 2], w3 = w[d * 4 + 3];
    const float bs = bias ? bias[d] : 0.f;
    float aw0 = 0.f, aw1 = 0.f, aw2 = 0.f, aw3 = 0.f, ab = 0.f;
    for (int64_t c0 = 0; c0 < L; c0 += 32 * E) {
        const int64_t l0 = c0 + (int64_t)lane * E;
        float v[E + 6], g[E + 3];
#pragma unroll
        for (int i = 0; i < E + 6; ++i) {
            const int64_t l = l0 - 3 + i;
            v[i] = (l >= 0 && l < L) ? ld1(xr + l) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < E + 3; ++i) {
            const int64_t t = l0 + i;
            float gi = 0.f;
            if (t < L) {
                gi = ld1(gr + t);
                if (silu) gi *= dsilu(fmaf(w3, v[i + 3], fmaf(w2, v[i + 2], fmaf(w1, v[i + 1], fmaf(w0, v[i], bs)))));
            }
            g[i] = gi;
        }
#pragma unroll
        for (int i = 0; i < E; ++i) {
            if (l0 + i < L) {
                // dx[s] = w3 g[s] + w2 g[s+1] + w1 g[s+2] + w0 g[s+3]
                st1(dxr + l0 + i, fmaf(w0, g[i + 3], fmaf(w1, g[i + 2], fmaf(w2, g[i + 1], w3 * g[i]))));
                aw0 = fmaf(g[i], v[i], aw0);
                aw1 = fmaf(g[i], v[i + 1], aw1);
                aw2 = fmaf(g[i], v[i + 2], aw2);
                aw3 = fmaf(g[i], v[i + 3], aw3);
                ab += g[i];
            }
        }
    }
    aw0 = warp_sum(aw0); aw1 = warp_sum(aw1); aw2 = warp_sum(aw2); aw3 = warp_sum(aw3); ab = warp_sum(ab);
    if (lane == 0) {
        atomicAdd(dw + d * 4 + 0, aw0); atomicAdd(dw + d * 4 + 1, aw1);
        atomicAdd(dw + d * 4 + 2, aw2); atomicAdd(dw + d * 4 + 3, aw3);
        if (db) atomicAdd(db + d, ab);
    }
}

// out[d] += sum_{b, l} a[b,d,l] * c[b,d,l]  (gradient of the D skip: dD = sum dout * conv1d_out,
// selective_scan_interface.py:640-642): one warp per (b, d) row, one atomic per row; caller zero-fills out.
template <typename T>
__global__ void __launch_bounds__(128)
rowdot_bdl_kernel(int64_t nrows, int dim, int64_t L, const T* __restrict__ a, const T* __restrict__ c,
                  float* __restrict__ out) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= nrows) return;
    float acc = 0.f;
    for (int64_t l = lane; l < L; l += 32) acc = fmaf(ld1(a + row * L + l), ld1(c + row * L + l), acc);
    acc = warp_sum(acc);
    if (lane == 0) atomicAdd(out + (int)(row % dim), acc);
}

static inline unsigned grid_for(int64_t items) { return (unsigned)((items + 255) / 256); }

}  // namespace fv

extern "C" int fv_causal_conv1d_fwd(int dtype, int batch, int dim, int64_t L, const void* x, int64_t x_bstride,
                                    int64_t x_dstride, const float* w, const float* bias, int silu, void* out,
                                    void* stream) {
    using namespace fv;
    FV_REQUIRE(batch > 0 && dim > 0 && L > 0, "fv_causal_conv1d_fwd: non-positive size");
    FV_REQUIRE(x && w && out, "fv_causal_conv1d_fwd: null pointer");
    const int64_t nrows = (int64_t)batch * dim;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FV_F32) {
        const int64_t chunks = (L + 3) / 4;
        FV_LAUNCH_PDL((causal_conv1d_bdl_kernel<float, 4>), grid_for(nrows * chunks), 256, 0, st, 
            nrows, dim, L, chunks, (const float*)x, x_bstride, x_dstride, w, bias, silu, (float*)out);
    } else if (dtype == FV_BF16) {
        const int64_t chunks = (L + 7) / 8;
        FV_LAUNCH_PDL((causal_conv1d_bdl_kernel<bf16, 8>), grid_for(nrows * chunks), 256, 0, st, 
            nrows, dim, L, chunks, (const bf16*)x, x_bstride, x_dstride, w, bias, silu, (bf16*)out);
    } else {
        return fail("fv_causal_conv1d_fwd: unsupported dtype %d", dtype);
    }
    return finish_launch("causal_conv1d_bdl");
}

extern "C" int fv_pool_bdl_fwd(int dtype, int batch, int dim, int outer, int pool, int inner, const void* x,
                               int pool_mode, float scale, void* out, void* stream) {
    using namespace fv;
    FV_REQUIRE(batch > 0 && dim > 0 && outer > 0 && pool > 0 && inner > 0, "fv_pool_bdl_fwd: non-positive size");
    FV_REQUIRE(x && out, "fv_pool_bdl_fwd: null pointer");
    const int64_t nrows = (int64_t)batch * dim, items = nrows * outer * inner;
    cudaStream_t st = (cudaStream_t)stream;
    const float sc = scale / (float)pool;
    if (dtype == FV_F32)
        FV_LAUNCH_PDL((pool_bdl_kernel<float>), grid_for(items), 256, 0, st, nrows, outer, pool, inner, (const float*)x,
                                                                pool_mode == FV_POOL_MAX, sc, (float*)out);
    else if (dtype == FV_BF16)
        FV_LAUNCH_PDL((pool_bdl_kernel<bf16>), grid_for(items), 256, 0, st, nrows, outer, pool, inner, (const bf16*)x,
                                                               pool_mode == FV_POOL_MAX, sc, (bf16*)out);
    else
        return fail("fv_pool_bdl_fwd: unsupported dtype %d", dtype);
    return finish_launch("pool_bdl");
}

extern "C" int fv_bcast_skip_bdl_fwd(int dtype, int batch, int dim, int outer, int pool, int inner, const void* s,
                                     const void* xc, const float* Dskip, void* out, void* stream) {
    using namespace fv;
    FV_REQUIRE(batch > 0 && dim > 0 && outer > 0 && pool > 0 && inner > 0, "fv_bcast_skip_bdl_fwd: non-positive size");
    FV_REQUIRE(s && out && (xc || !Dskip), "fv_bcast_skip_bdl_fwd: null pointer");
    const int64_t nrows = (int64_t)batch * dim, items = nrows * outer * pool * inner;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FV_F32)
        FV_LAUNCH_PDL((bcast_skip_bdl_kernel<float>), grid_for(items), 256, 0, st, nrows, dim, outer, pool, inner, (const float*)s,
                                                                      (const float*)xc, Dskip, (float*)out);
    else if (dtype == FV_BF16)
        FV_LAUNCH_PDL((bcast_skip_bdl_kernel<bf16>), grid_for(items), 256, 0, st, nrows, dim, outer, pool, inner, (const bf16*)s,
                                                                     (const bf16*)xc, Dskip, (bf16*)out);
    else
        return fail("fv_bcast_skip_bdl_fwd: unsupported dtype %d", dtype);
    return finish_launch("bcast_skip_bdl");
}

extern "C" int fv_causal_conv1d_bwd(int dtype, int batch, int dim, int64_t L, const void* x, int64_t x_bstride,
                                    int64_t x_dstride, const float* w, const float* bias, int silu, const void* dout,
                                    void* dx, int64_t dx_bstride, int64_t dx_dstride, float* dw, float* dbias,
                                    void* stream) {
    using namespace fv;
    FV_REQUIRE(batch > 0 && dim > 0 && L > 0, "fv_causal_conv1d_bwd: non-positive size");
    FV_REQUIRE(x && w && dout && dx && dw && (dbias || !bias), "fv_causal_conv1d_bwd: null pointer");
    const int64_t nrows = (int64_t)batch * dim;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((nrows + 3) / 4);
    if (dtype == FV_F32)
        FV_LAUNCH_PDL((causal_conv1d_bdl_bwd_kernel<float, 4>), grid, 128, 0, st, nrows, dim, L, (const float*)x, x_bstride, x_dstride, w, bias,
                                                                     silu, (const float*)dout, (float*)dx, dx_bstride,
                                                                     dx_dstride, dw, dbias);
    else if (dtype == FV_BF16)
        FV_LAUNCH_PDL((causal_conv1d_bdl_bwd_kernel<bf16, 8>), grid, 128, 0, st, nrows, dim, L, (const bf16*)x, x_bstride, x_dstride, w, bias,
                                                                    silu, (const bf16*)dout, (bf16*)dx, dx_bstride, dx_dstride,
                                                                    dw, dbias);
    else
        return fail("fv_causal_conv1d_bwd: unsupported dtype %d", dtype);
    return finish_launch("causal_conv1d_bdl_bwd");
}

extern "C" int fv_rowdot_bdl(int dtype, int batch, int dim, int64_t L, const void* a, const void* c, float* out,
                             void* stream) {
    using namespace fv;
    FV_REQUIRE(batch > 0 && dim > 0 && L > 0, "fv_rowdot_bdl: non-positive size");
    FV_REQUIRE(a && c && out, "fv_rowdot_bdl: null pointer");
    const int64_t nrows = (int64_t)batch * dim;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((nrows + 3) / 4);
    if (dtype == FV_F32)
        FV_LAUNCH_PDL((rowdot_bdl_kernel<float>), grid, 128, 0, st, nrows, dim, L, (const float*)a, (const float*)c, out);
    else if (dtype == FV_BF16)
        FV_LAUNCH_PDL((rowdot_bdl_kernel<bf16>), grid, 128, 0, st, nrows, dim, L, (const bf16*)a, (const bf16*)c, out);
    else
        return fail("fv_rowdot_bdl: unsupported dtype %d", dtype);
    return finish_launch("rowdot_bdl");
}
