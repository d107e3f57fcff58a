// Fused residual-add + RMSNorm / LayerNorm, prenorm form (forward and backward).
//
// Replaces the Triton kernels of mamba_ssm/ops/triton/layernorm.py:66-121 (fwd) and :210-305 (bwd)
// as called by Block.forward (models/fastvim.py:167-190) and the final norm (:519-537):
//   residual_out = x + residual          (kept in fp32: residual_in_fp32=True, fastvim.py:706)
//   y = residual_out * rstd * w (+ b)     (RMSNorm)   |   (residual_out - mean) * rstd * w + b (LayerNorm)
// One warp per token row, 4 contiguous channels per lane per step (128-bit fp32 / 64-bit bf16
// accesses), row kept in registers between the reduction and the normalisation: every byte is
// read once and written once -- HBM-bound: (s + 4) B read + (s + 4) B written per element.
#include "common.cuh"

namespace fv {

constexpr int NORM_WARPS = 4;

template <typename T, int NV, bool RMS>
__global__ void __launch_bounds__(NORM_WARPS * 32)
add_norm_fwd_kernel(int64_t rows, int cols, const T* __restrict__ x, int64_t ldx,
                    const float* __restrict__ res_in, const float* __restrict__ w,
                    const float* __restrict__ bias, float eps, T* __restrict__ y, int64_t ldy,
                    float* __restrict__ res_out, float* __restrict__ mean_out,
                    float* __restrict__ rstd_out) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * NORM_WARPS + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int nvec = cols >> 2;
    float4 r[NV];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
            float4 a = ld4(x + row * ldx + v * 4);
            if (res_in) a = a + ld4(res_in + row * cols + v * 4);
            r[i] = a;
            if (res_out) st4(res_out + row * cols + v * 4, a);
            sum += RMS ? fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, a.w * a.w))) : (a.x + a.y) + (a.z + a.w);
        }
    }
    sum = warp_sum(sum);
    float mean = 0.f, rstd;
    if (RMS) {
        rstd = rsqrtf(sum / (float)cols + eps);
    } else {
        mean = sum / (float)cols;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int v = lane + i * 32;
            if (v < nvec) {
                float4 a = r[i];
                float dx = a.x - mean, dy = a.y - mean, dz = a.z - mean, dw = a.w - mean;
                sq += fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, dw * dw)));
            }
        }
        sq = warp_sum(sq);
        rstd = rsqrtf(sq / (float)cols + eps);
    }
    if (lane == 0) {
        if (mean_out) mean_out[row] = mean;
        if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
            float4 a = r[i], g = ld4(w + v * 4), bb = bias ? ld4(bias + v * 4) : zero4();
            float4 o;
            o.x = fmaf((a.x - mean) * rstd, g.x, bb.x);
            o.y = fmaf((a.y - mean) * rstd, g.y, bb.y);
            o.z = fmaf((a.z - mean) * rstd, g.z, bb.z);
            o.w = fmaf((a.w - mean) * rstd, g.w, bb.w);
            st4(y + row * ldy + v * 4, o);
        }
    }
}

// Backward of the prenorm add + norm: r = x + residual (saved, fp32), y = norm(r) * w (+ b).
//   g = dy * w;  RMS: dr = rstd (g - xhat mean(g xhat)),  xhat = r rstd
//                LN : dr = rstd (g - mean(g) - xhat mean(g xhat)),  xhat = (r - mean) rstd
//   dr += d(residual_out);  dx = dr (activation dtype), d(residual_in) = dr (fp32);  dw = sum dy xhat, db = sum dy
// (reference: _layer_norm_bwd_kernel, mamba_ssm/ops/triton/layernorm.py:210-305).  One warp per row, rows
// grid-strided so the per-lane dw / db accumulators stay in registers; one atomicAdd per warp at the end.
// The statistics are recomputed from the saved fp32 row (already in registers) instead of being stored.
template <typename T, int NV, bool RMS>
__global__ void __launch_bounds__(NORM_WARPS * 32)
add_norm_bwd_kernel(int64_t rows, int cols, const T* __restrict__ dy, int64_t lddy,
                    const float* __restrict__ dres_out, const float* __restrict__ res_out,
                    const float* __restrict__ w, float eps, T* __restrict__ dx, int64_t lddx,
                    float* __restrict__ dres_in, float* __restrict__ dw, float* __restrict__ db) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int nvec = cols >> 2;
    const float inv = 1.f / (float)cols;
    float4 wv[NV], aw[NV], ab[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        wv[i] = v < nvec ? ld4(w + v * 4) : zero4();
        aw[i] = zero4();
        ab[i] = zero4();
    }
    const int64_t warp0 = (int64_t)blockIdx.x * NORM_WARPS + (threadIdx.x >> 5);
    const int64_t nwarp = (int64_t)gridDim.x * NORM_WARPS;
    for (int64_t row = warp0; row < rows; row += nwarp) {
        float4 r[NV], gy[NV];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int v = lane + i * 32;
            r[i] = v < nvec ? ld4(res_out + row * cols + v * 4) : zero4();
            gy[i] = v < nvec ? ld4(dy + row * lddy + v * 4) : zero4();
            s1 += (r[i].x + r[i].y) + (r[i].z + r[i].w);
            s2 += fmaf(r[i].x, r[i].x, fmaf(r[i].y, r[i].y, fmaf(r[i].z, r[i].z, r[i].w * r[i].w)));
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        float mean = 0.f, rstd;
        if (RMS) {
            rstd = rsqrtf(s2 * inv + eps);
        } else {
            mean = s1 * inv;
            float sq = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int v = lane + i * 32;
                if (v < nvec) {
                    const float a = r[i].x - mean, b = r[i].y - mean, c = r[i].z - mean, d = r[i].w - mean;
                    sq += fmaf(a, a, fmaf(b, b, fmaf(c, c, d * d)));
                }
            }
            rstd = rsqrtf(warp_sum(sq) * inv + eps);
        }
        float c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            // r <- xhat, gy <- g = dy * w  (dw, db use dy before the scaling)
            r[i] = make_float4((r[i].x - mean) * rstd, (r[i].y - mean) * rstd, (r[i].z - mean) * rstd, (r[i].w - mean) * rstd);
            if (lane + i * 32 >= nvec) r[i] = zero4();
            aw[i] = fma4(gy[i], r[i], aw[i]);
            if (!RMS) ab[i] = ab[i] + gy[i];   // RMSNorm has no bias (ops/triton/layernorm.py:515-536)
            gy[i] = make_float4(gy[i].x * wv[i].x, gy[i].y * wv[i].y, gy[i].z * wv[i].z, gy[i].w * wv[i].w);
            c1 += (gy[i].x + gy[i].y) + (gy[i].z + gy[i].w);
            c2 += fmaf(gy[i].x, r[i].x, fmaf(gy[i].y, r[i].y, fmaf(gy[i].z, r[i].z, gy[i].w * r[i].w)));
        }
        c1 = RMS ? 0.f : warp_sum(c1) * inv;
        c2 = warp_sum(c2) * inv;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int v = lane + i * 32;
            if (v < nvec) {
                float4 o = make_float4(rstd * (gy[i].x - c1 - r[i].x * c2), rstd * (gy[i].y - c1 - r[i].y * c2),
                                       rstd * (gy[i].z - c1 - r[i].z * c2), rstd * (gy[i].w - c1 - r[i].w * c2));
                if (dres_out) o = o + ld4(dres_out + row * cols + v * 4);
                if (dres_in) st4(dres_in + row * cols + v * 4, o);
                if (dx) st4(dx + row * lddx + v * 4, o);
            }
        }
    }
    // weight / bias gradients: sum the CTA's warps in shared memory first, then ONE atomic per CTA and column
    // (per-warp atomics put ~9,000 serialized fp32 adds on each of the `cols` addresses)
    __shared__ float4 red[NORM_WARPS][32];
    const int wid = threadIdx.x >> 5;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1 && (RMS || !db)) break;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            __syncthreads();
            red[wid][lane] = pass ? ab[i] : aw[i];
            __syncthreads();
            const int v = lane + i * 32;
            if (wid == 0 && v < nvec) {
                float4 t = red[0][lane];
#pragma unroll
                for (int k = 1; k < NORM_WARPS; ++k) t = t + red[k][lane];
                float* dst = (pass ? db : dw) + v * 4;
                atomicAdd(dst + 0, t.x); atomicAdd(dst + 1, t.y); atomicAdd(dst + 2, t.z); atomicAdd(dst + 3, t.w);
            }
        }
    }
}

int sm_count();

template <typename T>
static int launch_add_norm_bwd(int64_t rows, int cols, const T* dy, int64_t lddy, const float* dres_out,
                               const float* res_out, const float* w, float eps, int is_rms, T* dx, int64_t lddx,
                               float* dres_in, float* dw, float* db, cudaStream_t st) {
    int64_t blocks = (rows + NORM_WARPS - 1) / NORM_WARPS;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    dim3 grid((unsigned)blocks), block(NORM_WARPS * 32);
    const int nvec = cols / 4;
#define FV_NB_CASE(NV_)                                                                                                 \
    if (nvec <= NV_ * 32) {                                                                                             \
        if (is_rms) FV_LAUNCH_PDL((add_norm_bwd_kernel<T, NV_, true>), grid, block, 0, st, rows, cols, dy, lddy, dres_out, res_out, w, eps, dx, lddx, dres_in, dw, db); \
        else FV_LAUNCH_PDL((add_norm_bwd_kernel<T, NV_, false>), grid, block, 0, st, rows, cols, dy, lddy, dres_out, res_out, w, eps, dx, lddx, dres_in, dw, db);       \
        return finish_launch("add_norm_bwd");                                                                           \
    }
    FV_NB_CASE(2) FV_NB_CASE(3) FV_NB_CASE(4) FV_NB_CASE(6) FV_NB_CASE(8) FV_NB_CASE(12) FV_NB_CASE(16)
#undef FV_NB_CASE
    return fail("fv_add_norm_bwd: cols %d > 2048 not supported", cols);
}

template <typename T>
static int launch_add_norm(int64_t rows, int cols, const T* x, int64_t ldx, const float* res_in,
                           const float* w, const float* bias, float eps, int is_rms, T* y, int64_t ldy,
                           float* res_out, float* mean_out, float* rstd_out, cudaStream_t st) {
    dim3 grid((unsigned)((rows + NORM_WARPS - 1) / NORM_WARPS)), block(NORM_WARPS * 32);
    const int nvec = cols / 4;
#define FV_NORM_CASE(NV_)                                                                                   \
    if (nvec <= NV_ * 32) {                                                                                 \
        if (is_rms) FV_LAUNCH_PDL((add_norm_fwd_kernel<T, NV_, true>), grid, block, 0, st, rows, cols, x, ldx, res_in, w, bias, eps, y, ldy, res_out, mean_out, rstd_out); \
        else FV_LAUNCH_PDL((add_norm_fwd_kernel<T, NV_, false>), grid, block, 0, st, rows, cols, x, ldx, res_in, w, bias, eps, y, ldy, res_out, mean_out, rstd_out);       \
        return finish_launch("add_norm_fwd");                                                               \
    }
    FV_NORM_CASE(2) FV_NORM_CASE(4) FV_NORM_CASE(8) FV_NORM_CASE(16)
#undef FV_NORM_CASE
    return fail("fv_add_norm_fwd: cols %d > 2048 not supported", cols);
}

}  // namespace fv

extern "C" int fv_add_norm_fwd(int dtype, int64_t rows, int cols, const void* x, int64_t ldx,
                               const float* residual_in, const float* weight, const float* bias,
                               float eps, int is_rms, void* y, int64_t ldy, float* residual_out,
                               float* mean_out, float* rstd_out, void* stream) {
    using namespace fv;
    FV_REQUIRE(x && weight && y, "fv_add_norm_fwd: null pointer");
    FV_REQUIRE(rows > 0 && cols > 0 && cols % 4 == 0, "fv_add_norm_fwd: bad shape rows %lld cols %d", (long long)rows, cols);
    FV_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0, "fv_add_norm_fwd: row strides must be multiples of 4");
    FV_REQUIRE(rows / NORM_WARPS < (1ll << 31), "fv_add_norm_fwd: too many rows");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FV_F32)
        return launch_add_norm<float>(rows, cols, (const float*)x, ldx, residual_in, weight, bias, eps, is_rms, (float*)y, ldy, residual_out, mean_out, rstd_out, st);
    if (dtype == FV_BF16)
        return launch_add_norm<bf16>(rows, cols, (const bf16*)x, ldx, residual_in, weight, bias, eps, is_rms, (bf16*)y, ldy, residual_out, mean_out, rstd_out, st);
    return fail("fv_add_norm_fwd: unsupported dtype %d", dtype);
}

extern "C" int fv_add_norm_bwd(int dtype, int64_t rows, int cols, const void* dy, int64_t lddy,
                               const float* dresidual_out, const float* residual_out, const float* weight,
                               float eps, int is_rms, void* dx, int64_t lddx, float* dresidual_in, float* dweight,
                               float* dbias, void* stream) {
    using namespace fv;
    FV_REQUIRE(dy && residual_out && weight && dweight && (dx || dresidual_in), "fv_add_norm_bwd: null pointer");
    FV_REQUIRE(rows > 0 && cols > 0 && cols % 4 == 0, "fv_add_norm_bwd: bad shape rows %lld cols %d", (long long)rows, cols);
    FV_REQUIRE(lddy % 4 == 0 && lddx % 4 == 0, "fv_add_norm_bwd: row strides must be multiples of 4");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FV_F32)
        return launch_add_norm_bwd<float>(rows, cols, (const float*)dy, lddy, dresidual_out, residual_out, weight, eps, is_rms, (float*)dx, lddx, dresidual_in, dweight, dbias, st);
    if (dtype == FV_BF16)
        return launch_add_norm_bwd<bf16>(rows, cols, (const bf16*)dy, lddy, dresidual_out, residual_out, weight, eps, is_rms, (bf16*)dx, lddx, dresidual_in, dweight, dbias, st);
    return fail("fv_add_norm_bwd: unsupported dtype %d", dtype);
}

// ---- token-side LayerNorm + SiLU gate of the hybrid-sharded mode ------------------------------------------------------
// y = LayerNorm_D(v; gamma, beta) * silu(z)   (mamba_simple_faster.py:437-441; without gamma: y = v * silu(z), :445-453)
// for rows that hold ALL d_inner channels of a token: after the channel -> token all-to-all of the single-image
// multi-GPU mode the LayerNorm over d_inner is token-local, so no statistics have to cross GPUs.  One warp per row,
// row in registers, two-pass variance like add_norm_fwd_kernel.
namespace fv {

template <typename T, int NV>
__global__ void __launch_bounds__(NORM_WARPS * 32)
ln_gate_fwd_kernel(int64_t rows, int cols, const T* __restrict__ v, int64_t ldv, const T* __restrict__ z, int64_t ldz,
                   const float* __restrict__ w, const float* __restrict__ bias, float eps, T* __restrict__ y, int64_t ldy) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * NORM_WARPS + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int nvec = cols >> 2;
    float4 r[NV];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + i * 32;
        if (c < nvec) {
            r[i] = ld4(v + row * ldv + c * 4);
            sum += (r[i].x + r[i].y) + (r[i].z + r[i].w);
        }
    }
    float mean = 0.f, rstd = 1.f;
    if (w) {
        mean = warp_sum(sum) / (float)cols;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + i * 32;
            if (c < nvec) {
                const float dx = r[i].x - mean, dy = r[i].y - mean, dz = r[i].z - mean, dw = r[i].w - mean;
                sq += fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, dw * dw)));
            }
        }
        rstd = rsqrtf(warp_sum(sq) / (float)cols + eps);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + i * 32;
        if (c < nvec) {
            float4 a = r[i];
            if (w) {
                const float4 g = ld4(w + c * 4), bb = bias ? ld4(bias + c * 4) : zero4();
                a.x = fmaf((a.x - mean) * rstd, g.x, bb.x); a.y = fmaf((a.y - mean) * rstd, g.y, bb.y);
                a.z = fmaf((a.z - mean) * rstd, g.z, bb.z); a.w = fmaf((a.w - mean) * rstd, g.w, bb.w);
            }
            const float4 zz = ld4(z + row * ldz + c * 4);
            a.x *= silu_exact(zz.x); a.y *= silu_exact(zz.y); a.z *= silu_exact(zz.z); a.w *= silu_exact(zz.w);
            st4(y + row * ldy + c * 4, a);
        }
    }
}

template <typename T>
static int launch_ln_gate(int64_t rows, int cols, const T* v, int64_t ldv, const T* z, int64_t ldz, const float* w,
                          const float* b, float eps, T* y, int64_t ldy, cudaStream_t st) {
    const int nv = (cols / 4 + 31) / 32;
    const unsigned grid = (unsigned)((rows + NORM_WARPS - 1) / NORM_WARPS);
#define FV_LG(NV_) FV_LAUNCH_PDL((ln_gate_fwd_kernel<T, NV_>), grid, NORM_WARPS * 32, 0, st, rows, cols, v, ldv, z, ldz, w, b, eps, y, ldy)
    if (nv <= 1) FV_LG(1);
    else if (nv <= 2) FV_LG(2);
    else if (nv <= 3) FV_LG(3);
    else if (nv <= 6) FV_LG(6);
    else if (nv <= 12) FV_LG(12);
    else if (nv <= 24) FV_LG(24);
    else return fail("fv_ln_gate_fwd: cols %d too large (max 3072)", cols);
#undef FV_LG
    return finish_launch("ln_gate_fwd");
}

}  // namespace fv

extern "C" int fv_ln_gate_fwd(int dtype, int64_t rows, int cols, const void* v, int64_t ldv, const void* z, int64_t ldz,
                              const float* ln_w, const float* ln_b, float eps, void* y, int64_t ldy, void* stream) {
    using namespace fv;
    FV_REQUIRE(v && z && y, "fv_ln_gate_fwd: null pointer");
    FV_REQUIRE(rows > 0 && cols > 0 && cols % 4 == 0 && rows / NORM_WARPS < (1ll << 31), "fv_ln_gate_fwd: bad shape rows %lld cols %d", (long long)rows, cols);
    FV_REQUIRE(ldv % 4 == 0 && ldz % 4 == 0 && ldy % 4 == 0, "fv_ln_gate_fwd: row strides must be multiples of 4");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FV_F32)
        return launch_ln_gate<float>(rows, cols, (const float*)v, ldv, (const float*)z, ldz, ln_w, ln_b, eps, (float*)y, ldy, st);
    if (dtype == FV_BF16)
        return launch_ln_gate<bf16>(rows, cols, (const bf16*)v, ldv, (const bf16*)z, ldz, ln_w, ln_b, eps, (bf16*)y, ldy, st);
    return fail("fv_ln_gate_fwd: unsupported dtype %d", dtype);
}
