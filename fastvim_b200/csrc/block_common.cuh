// Device helpers shared by the fused block-interior kernels (block_fwd.cu: one CTA per image; block_cluster.cu: a thread-block
// cluster per image, channels split over its CTAs).
#pragma once
#include "common.cuh"

namespace fv {

constexpr int BK_NSTATE = 16;

__device__ __forceinline__ float2 unpack2(uint32_t v) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
}
__device__ __forceinline__ uint32_t pack2(float2 v) {
    __nv_bfloat162 r = __floats2bfloat162_rn(v.x, v.y);
    return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ uint32_t pack2(float a, float b) { return pack2(make_float2(a, b)); }

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float bk_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float bk_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// softplus with the reference threshold (fwd_kernel.cuh:153-156), 2 MUFU ops; see scan_pooled.cu
__device__ __forceinline__ float bk_softplus(float x) {
    constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
    const float e = bk_ex2(x * LOG2E);
    const float sp = x < -5.f ? e * fmaf(e, fmaf(e, 0.33333334f, -0.5f), 1.f) : LN2 * bk_lg2(1.f + e);
    return x <= 20.f ? sp : x;
}
// silu(x) for a pair, given h = x/2: h + h * tanh(h).  tanh.approx.f32 is ONE MUFU op per element with no
// conversions around it (the f16x2 form also costs one MUFU per element in SASS, plus a pack and two unpacks).
__device__ __forceinline__ float bk_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float2 silu2_from_half(float2 h) {
    return __ffma2_rn(h, make_float2(bk_tanh(h.x), bk_tanh(h.y)), h);
}

// memory token row of sequence position t (plain geometry: inner == 1)
template <int POOL_T>
__device__ __forceinline__ int64_t bk_row(const Geom& g, int t) {
    const int P = POOL_T ? POOL_T : g.pool;
    const int o = t / P, p = t - o * P;
    return o * g.so + p * g.sp;
}


}  // namespace fv
