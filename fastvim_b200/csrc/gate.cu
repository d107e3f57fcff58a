// K2b -- scan epilogue: broadcast-back + D skip + direction average + LayerNorm(d_inner) + SiLU(z) gate.
//
// What it replaces in the reference (paths relative to /root/reference):
//   out.repeat_interleave(num_of_col, 2); out += D * x            mamba_simple_faster.py:356-358
//   the same for the b direction on x_flip                        :412-416
//   layernorm(rearrange(out + out_b.flip(-1)) / 2) * silu(z)      :434-441   (no-norm form :445-453)
// (>= 10 full-resolution passes incl. a flip and a (B,D,L)->(B,L,D) transpose copy) by ONE
// kernel that reads x and z once and writes the gated out_proj input once: 3 full-resolution
// tensors, the compulsory minimum.  The conv outputs xc_f / xc_b needed by the D skip are
// recomputed from x (7-row sliding register window) rather than stored by K1.
//
// Mapping: token-major; a CTA owns ALL d_inner channels (4 per thread) of up to TT tokens of
// one pooled position j, so s[b, j, :] is loaded once and the LayerNorm reduction over
// d_inner stays inside the CTA: pre-norm values go to shared memory, per-token (sum, sumsq)
// partials are warp-shuffled, one barrier, then normalise * gamma + beta, * silu(z), store.
// In channel-sharded mode (2048^2 single image, d_inner split over GPUs) the CTA sees only a
// shard of the channels: it writes the pre-norm value and the shard's per-token partial
// statistics; fv_norm_gate_apply finishes after the all-reduce of the (B, L, 2) statistics.
#include "common.cuh"

namespace fv {

template <typename T>
__device__ __forceinline__ float4 gload_row4(const Geom& g, const T* xb, int64_t ldx, int d0, int t, bool live) {
    if (!live || t < 0 || t >= g.L) return zero4();
    return ld4(xb + seq_to_row(g, t) * ldx + d0);
}

template <typename T, bool INNER1, int MAXT>
__global__ void __launch_bounds__(MAXT)
gate_fwd_kernel(Geom g, int TT, const T* __restrict__ x, const T* __restrict__ z, int64_t ldxz,
                int64_t xzbs, const float* __restrict__ s, const float* __restrict__ cw,
                const float* __restrict__ cb, const float* __restrict__ Dskip,
                const float* __restrict__ lnw, const float* __restrict__ lnb, float eps,
                T* __restrict__ y, int64_t ldy, int64_t ybs, float* __restrict__ stats) {
    constexpr bool FAST = is_fast<T>::value;
    extern __shared__ __align__(16) float smem[];
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* vbuf = smem;                          // [TT][D]
    float2* part = reinterpret_cast<float2*>(smem + (size_t)TT * g.D);  // [TT][nwarps]

    const int j = blockIdx.y, b = blockIdx.z;
    const int p_lo = blockIdx.x * TT;
    const int np = min(TT, g.pool - p_lo);
    const int d0 = threadIdx.x * 4;
    const bool live = d0 < g.D;
    const int dd = live ? d0 : 0;
    const T* xb = x + (int64_t)b * xzbs;
    const T* zb = z + (int64_t)b * xzbs;
    T* yb = y + (int64_t)b * ybs;
    const bool has_norm = lnw != nullptr;
    const bool sharded = stats != nullptr;

    const Taps tf = load_taps(cw, cb, g.D, 0, dd), tb = load_taps(cw, cb, g.D, 1, dd);
    const float4 sv = ld4(s + ((int64_t)b * g.Lp + j) * g.D + dd);
    const float4 Df = ld4(Dskip + dd), Db = ld4(Dskip + g.D + dd);

    float4 w[7];
    constexpr int PF = 4;
    float4 ring[PF];
    const int t0 = INNER1 ? j * g.pool + p_lo : 0;
    if (INNER1) {
#pragma unroll
        for (int i = 0; i < 6; ++i) w[i + 1] = gload_row4(g, xb, ldxz, dd, t0 - 3 + i, live);
#pragma unroll
        for (int i = 0; i < PF; ++i) ring[i] = gload_row4(g, xb, ldxz, dd, t0 + 3 + i, live);
    }
    for (int p0 = 0; p0 < np; p0 += PF) {
#pragma unroll
        for (int i = 0; i < PF; ++i) {
            const int p = p0 + i;
            if (p < np) {
                int t;
                if (INNER1) {
                    t = t0 + p;
#pragma unroll
                    for (int k = 0; k < 6; ++k) w[k] = w[k + 1];
                    w[6] = ring[i];
                    ring[i] = (p + PF < np) ? gload_row4(g, xb, ldxz, dd, t0 + 3 + p + PF, live) : zero4();
                } else {
                    t = pooled_to_seq(g, j, p_lo + p);
#pragma unroll
                    for (int k = 0; k < 7; ++k) w[k] = gload_row4(g, xb, ldxz, dd, t - 3 + k, live);
                }
                float4 af = tf.b, ab = tb.b;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    af = fma4(tf.w[k], w[k], af);
                    ab = fma4(tb.w[k], w[6 - k], ab);
                }
                af = silu4<FAST>(af);
                ab = silu4<FAST>(ab);
                float4 v = scale4(fma4(Db, ab, fma4(Df, af, sv)), 0.5f);
                if (!live) v = zero4();
                if (has_norm || sharded) {
                    if (live) st4(vbuf + (size_t)p * g.D + d0, v);
                    float sum = (v.x + v.y) + (v.z + v.w);
                    float sq = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
                    sum = warp_sum(sum);
                    sq = warp_sum(sq);
                    if (lane == 0) part[p * nwarps + warp] = make_float2(sum, sq);
                } else if (live) {
                    const int64_t row = seq_to_row(g, t);
                    float4 zz = silu4<FAST>(ld4(zb + row * ldxz + d0));
                    st4(yb + row * ldy + d0, make_float4(v.x * zz.x, v.y * zz.y, v.z * zz.z, v.w * zz.w));
                }
            }
        }
    }
    if (!(has_norm || sharded)) return;
    __syncthreads();
    float4 gam = make_float4(1.f, 1.f, 1.f, 1.f), bet = zero4();
    if (has_norm && live) {
        gam = ld4(lnw + d0);
        if (lnb) bet = ld4(lnb + d0);
    }
    const float invD = 1.f / (float)g.D;
    for (int p = 0; p < np; ++p) {
        const int t = INNER1 ? t0 + p : pooled_to_seq(g, j, p_lo + p);
        const int64_t row = seq_to_row(g, t);
        float sum = 0.f, sq = 0.f;
        for (int wq = 0; wq < nwarps; ++wq) {
            float2 q = part[p * nwarps + wq];
            sum += q.x;
            sq += q.y;
        }
        if (sharded) {
            if (threadIdx.x == 0) {
                float* so = stats + ((int64_t)b * g.L + row) * 2;
                so[0] = sum;
                so[1] = sq;
            }
            if (live) st4(yb + row * ldy + d0, ld4(vbuf + (size_t)p * g.D + d0));
            continue;
        }
        if (!live) continue;
        const float mean = sum * invD;
        const float rstd = rsqrtf(fmaxf(sq * invD - mean * mean, 0.f) + eps);
        float4 v = ld4(vbuf + (size_t)p * g.D + d0);
        float4 zz = silu4<FAST>(ld4(zb + row * ldxz + d0));
        float4 o;
        o.x = fmaf((v.x - mean) * rstd, gam.x, bet.x) * zz.x;
        o.y = fmaf((v.y - mean) * rstd, gam.y, bet.y) * zz.y;
        o.z = fmaf((v.z - mean) * rstd, gam.z, bet.z) * zz.z;
        o.w = fmaf((v.w - mean) * rstd, gam.w, bet.w) * zz.w;
        st4(yb + row * ldy + d0, o);
    }
}

// Finishes the channel-sharded path: y holds pre-norm values of this shard, stats the
// all-reduced per-token (sum, sumsq) over the full d_inner.
template <typename T>
__global__ void __launch_bounds__(256)
norm_gate_apply_kernel(Geom g, int full_dim, T* __restrict__ y, int64_t ldy, int64_t ybs,
                       const T* __restrict__ z, int64_t ldz, int64_t zbs,
                       const float* __restrict__ stats, const float* __restrict__ lnw,
                       const float* __restrict__ lnb, float eps) {
    constexpr bool FAST = is_fast<T>::value;
    const int nvec = g.D >> 2;
    const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= (int64_t)g.B * g.L * nvec) return;
    const int d0 = (int)(item % nvec) * 4;
    const int64_t bt = item / nvec;
    const int64_t row = bt % g.L, b = bt / g.L;
    const float sum = stats[bt * 2], sq = stats[bt * 2 + 1];
    const float mean = sum / (float)full_dim;
    const float rstd = rsqrtf(fmaxf(sq / (float)full_dim - mean * mean, 0.f) + eps);
    float4 v = ld4(y + b * ybs + row * ldy + d0);
    float4 zz = silu4<FAST>(ld4(z + b * zbs + row * ldz + d0));
    float4 gam = lnw ? ld4(lnw + d0) : make_float4(1.f, 1.f, 1.f, 1.f), bet = lnb ? ld4(lnb + d0) : zero4();
    float4 o;
    if (lnw) {
        o.x = fmaf((v.x - mean) * rstd, gam.x, bet.x) * zz.x;
        o.y = fmaf((v.y - mean) * rstd, gam.y, bet.y) * zz.y;
        o.z = fmaf((v.z - mean) * rstd, gam.z, bet.z) * zz.z;
        o.w = fmaf((v.w - mean) * rstd, gam.w, bet.w) * zz.w;
    } else {
        o = make_float4(v.x * zz.x, v.y * zz.y, v.z * zz.z, v.w * zz.w);
    }
    st4(y + b * ybs + row * ldy + d0, o);
}

int check_geom(const fv_geom* g, const char* who);

template <typename T>
static int launch_gate(const Geom& g, const T* x, const T* z, int64_t ldxz, int64_t xzbs, const float* s,
                       const float* cw, const float* cb, const float* Dskip, const float* lnw,
                       const float* lnb, float eps, T* y, int64_t ldy, int64_t ybs, float* stats,
                       cudaStream_t st) {
    const int threads = ((g.D / 4) + 31) / 32 * 32;
    FV_REQUIRE(threads <= 1024, "fv_gate_fwd: dim %d > 4096 not supported", g.D);
    const int nwarps = threads / 32;
    int TT = g.pool < 16 ? g.pool : 16;
    auto smem_of = [&](int tt) { return (size_t)tt * g.D * 4 + (size_t)tt * nwarps * 8; };
    while (TT > 1 && smem_of(TT) > 64 * 1024) TT = (TT + 1) / 2;
    const size_t smem = smem_of(TT);
    FV_REQUIRE(smem <= 200 * 1024, "fv_gate_fwd: shared memory %zu too large", smem);
    FV_REQUIRE(g.Lp <= 65535 && g.B <= 65535, "fv_gate_fwd: Lp or batch > 65535");
    dim3 grid(ceil_div(g.pool, TT), g.Lp, g.B), block(threads);
    void (*kern)(Geom, int, const T*, const T*, int64_t, int64_t, const float*, const float*, const float*,
                 const float*, const float*, const float*, float, T*, int64_t, int64_t, float*);
    const bool in1 = g.inner == 1;
    if (threads <= 128) kern = in1 ? gate_fwd_kernel<T, true, 128> : gate_fwd_kernel<T, false, 128>;
    else if (threads <= 256) kern = in1 ? gate_fwd_kernel<T, true, 256> : gate_fwd_kernel<T, false, 256>;
    else if (threads <= 512) kern = in1 ? gate_fwd_kernel<T, true, 512> : gate_fwd_kernel<T, false, 512>;
    else kern = in1 ? gate_fwd_kernel<T, true, 1024> : gate_fwd_kernel<T, false, 1024>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        FV_REQUIRE(e == cudaSuccess, "fv_gate_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    kern<<<grid, block, smem, st>>>(g, TT, x, z, ldxz, xzbs, s, cw, cb, Dskip, lnw, lnb, eps, y, ldy, ybs, stats);
    return finish_launch("gate_fwd");
}

}  // namespace fv

extern "C" int fv_gate_fwd(const fv_geom* g_, int dtype, const void* x, const void* z, int64_t ldxz,
                           int64_t xz_bstride, const float* s, const float* conv_w,
                           const float* conv_b, const float* Dskip, const float* ln_w,
                           const float* ln_b, float eps, void* y, int64_t ldy, int64_t y_bstride,
                           float* stats, void* stream) {
    using namespace fv;
    if (int rc = check_geom(g_, "fv_gate_fwd")) return rc;
    FV_REQUIRE(x && z && s && conv_w && Dskip && y, "fv_gate_fwd: null pointer");
    FV_REQUIRE(ldxz % 4 == 0 && xz_bstride % 4 == 0 && ldy % 4 == 0 && y_bstride % 4 == 0,
               "fv_gate_fwd: strides must be multiples of 4 elements");
    Geom g = make_geom(g_);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FV_F32)
        return launch_gate<float>(g, (const float*)x, (const float*)z, ldxz, xz_bstride, s, conv_w, conv_b, Dskip, ln_w, ln_b, eps, (float*)y, ldy, y_bstride, stats, st);
    if (dtype == FV_BF16)
        return launch_gate<bf16>(g, (const bf16*)x, (const bf16*)z, ldxz, xz_bstride, s, conv_w, conv_b, Dskip, ln_w, ln_b, eps, (bf16*)y, ldy, y_bstride, stats, st);
    return fail("fv_gate_fwd: unsupported dtype %d", dtype);
}

extern "C" int fv_norm_gate_apply(const fv_geom* g_, int dtype, int full_dim, void* y, int64_t ldy,
                                  int64_t y_bstride, const void* z, int64_t ldz, int64_t z_bstride,
                                  const float* stats, const float* ln_w, const float* ln_b, float eps,
                                  void* stream) {
    using namespace fv;
    if (int rc = check_geom(g_, "fv_norm_gate_apply")) return rc;
    FV_REQUIRE(y && z && stats, "fv_norm_gate_apply: null pointer");
    FV_REQUIRE(full_dim >= g_->dim, "fv_norm_gate_apply: full_dim %d < dim %d", full_dim, g_->dim);
    Geom g = make_geom(g_);
    const int64_t items = (int64_t)g.B * g.L * (g.D / 4);
    dim3 grid((unsigned)((items + 255) / 256)), block(256);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FV_F32)
        norm_gate_apply_kernel<float><<<grid, block, 0, st>>>(g, full_dim, (float*)y, ldy, y_bstride, (const float*)z, ldz, z_bstride, stats, ln_w, ln_b, eps);
    else if (dtype == FV_BF16)
        norm_gate_apply_kernel<bf16><<<grid, block, 0, st>>>(g, full_dim, (bf16*)y, ldy, y_bstride, (const bf16*)z, ldz, z_bstride, stats, ln_w, ln_b, eps);
    else
        return fail("fv_norm_gate_apply: unsupported dtype %d", dtype);
    return finish_launch("norm_gate_apply");
}
