// fastvim_b200 -- helpers shared by the "streaming" backward kernels (conv_pool_bwd.cu, gate_bwd_stream.cu):
// a thread owns a PAIR of adjacent channels and walks consecutive tokens with its state in registers.
#pragma once
#include "common.cuh"

namespace fv {

// a pair of adjacent channels of one token row: one 4-byte (bf16) / 8-byte (fp32) access
template <typename T> struct Pair;
template <> struct Pair<bf16> {
    typedef uint32_t type;
    static __device__ __forceinline__ type ld(const bf16* p) { return __ldg(reinterpret_cast<const unsigned int*>(p)); }
    static __device__ __forceinline__ type zero() { return 0u; }
    static __device__ __forceinline__ float2 up(type v) { return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v)); }
    static __device__ __forceinline__ void st(bf16* p, float2 v) { *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v.x, v.y); }
};
template <> struct Pair<float> {
    typedef float2 type;
    static __device__ __forceinline__ type ld(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
    static __device__ __forceinline__ type zero() { return make_float2(0.f, 0.f); }
    static __device__ __forceinline__ float2 up(type v) { return v; }
    static __device__ __forceinline__ void st(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
};
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }

// block size for the streaming kernel: the largest multiple of 32 that is <= 256 and divides dim / 2 (0: none)
static inline int stream_block(int D) {
    if (D % 64 != 0) return 0;
    const int pairs = D / 2;
    for (int t = 256; t >= 32; t -= 32)
        if (pairs % t == 0) return t;
    return 0;
}


}  // namespace fv
