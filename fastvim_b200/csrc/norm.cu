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

template <typename T>
static int launch_add_norm(int64_t rows, int cols, const T* x, int64_t ldx, const float* res_in,
                           const float* w, const float* bias, float eps, int is_rms, T* y, int64_t ldy,
                           float* res_out, float* mean_out, float* rstd_out, cudaStream_t st) {
    dim3 grid((unsigned)((rows + NORM_WARPS - 1) / NORM_WARPS)), block(NORM_WARPS * 32);
    const int nvec = cols / 4;
#define FV_NORM_CASE(NV_)                                                                                   \
    if (nvec <= NV_ * 32) {                                                                                 \
        if (is_rms) add_norm_fwd_kernel<T, NV_, true><<<grid, block, 0, st>>>(rows, cols, x, ldx, res_in, w, bias, eps, y, ldy, res_out, mean_out, rstd_out); \
        else add_norm_fwd_kernel<T, NV_, false><<<grid, block, 0, st>>>(rows, cols, x, ldx, res_in, w, bias, eps, y, ldy, res_out, mean_out, rstd_out);       \
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
