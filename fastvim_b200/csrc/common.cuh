// fastvim_b200 -- shared device/host helpers (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "fastvim_b200.h"

namespace fv {

typedef __nv_bfloat16 bf16;

// ---- host side -------------------------------------------------------------------
void set_error(const char* fmt, ...);
int finish_launch(const char* what);  // counts the launch, maps cudaGetLastError to rc
int fail(const char* fmt, ...);       // set_error + return 1

#define FV_REQUIRE(cond, ...)                      \
    do {                                           \
        if (!(cond)) return ::fv::fail(__VA_ARGS__); \
    } while (0)

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- geometry --------------------------------------------------------------------
struct Geom {
    int B, D, outer, pool, inner, Lp, L;
    int64_t so, sp, si;
};
static inline Geom make_geom(const fv_geom* g) {
    Geom r;
    r.B = g->batch; r.D = g->dim; r.outer = g->outer; r.pool = g->pool; r.inner = g->inner;
    r.Lp = g->outer * g->inner; r.L = g->outer * g->pool * g->inner;
    r.so = g->tok_stride_outer; r.sp = g->tok_stride_pool; r.si = g->tok_stride_inner;
    return r;
}
// sequence position t in [0, L) -> memory token row of the image
__device__ __forceinline__ int64_t seq_to_row(const Geom& g, int t) {
    if (g.inner == 1) {
        int o = t / g.pool, p = t - o * g.pool;
        return o * g.so + p * g.sp;
    }
    int q = t / g.inner, i = t - q * g.inner;
    int o = q / g.pool, p = q - o * g.pool;
    return o * g.so + p * g.sp + i * g.si;
}
// pooled position j, slot p -> sequence position
__device__ __forceinline__ int pooled_to_seq(const Geom& g, int j, int p) {
    if (g.inner == 1) return j * g.pool + p;
    int o = j / g.inner, i = j - o * g.inner;
    return (o * g.pool + p) * g.inner + i;
}

// ---- 4-wide vector access (V = 4 channels per thread) -----------------------------
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const bf16* p) {
    uint2 r = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(bf16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = r;
}
__device__ __forceinline__ float ld1(const float* p) { return *p; }
__device__ __forceinline__ float ld1(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }
__device__ __forceinline__ void st1(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// ---- transcendental helpers --------------------------------------------------------
// The non-GEMM half of the block is co-limited by the MUFU pipe (16 ops/clk/SM) and HBM:
// per full-resolution element the path needs 5 SiLUs (2 conv directions in K1, the same 2
// recomputed in K2b, and the z gate).  exp+rcp costs 2 MUFU per SiLU.  For bf16 I/O we use
// silu(x) = h + h*tanh(h), h = x/2, with tanh.approx.f16x2: ONE MUFU op per TWO elements
// (abs error ~2^-11, four times below the bf16 rounding the result receives anyway).
// fp32 I/O keeps the exact form.
__device__ __forceinline__ float silu_exact(float x) { return __fdividef(x, 1.f + __expf(-x)); }
__device__ __forceinline__ void silu_pair_fast(float& a, float& b) {
    float ha = 0.5f * a, hb = 0.5f * b;
    __half2 h = __floats2half2_rn(ha, hb);
    uint32_t hi = *reinterpret_cast<uint32_t*>(&h), ho;
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(ho) : "r"(hi));
    float2 t = __half22float2(*reinterpret_cast<__half2*>(&ho));
    a = fmaf(ha, t.x, ha);
    b = fmaf(hb, t.y, hb);
}
template <bool FAST>
__device__ __forceinline__ float4 silu4(float4 v) {
    if (FAST) {
        silu_pair_fast(v.x, v.y);
        silu_pair_fast(v.z, v.w);
    } else {
        v.x = silu_exact(v.x); v.y = silu_exact(v.y); v.z = silu_exact(v.z); v.w = silu_exact(v.w);
    }
    return v;
}
// SiLU of a conv output computed with taps pre-scaled by 0.5 (FAST) / 1 (exact): for FAST the
// input is h = x/2 and silu(x) = h + h*tanh(h).
__device__ __forceinline__ void silu_pair_from_half(float& ha, float& hb) {
    __half2 h = __floats2half2_rn(ha, hb);
    uint32_t hi = *reinterpret_cast<uint32_t*>(&h), ho;
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(ho) : "r"(hi));
    float2 t = __half22float2(*reinterpret_cast<__half2*>(&ho));
    ha = fmaf(ha, t.x, ha);
    hb = fmaf(hb, t.y, hb);
}
template <bool FAST>
__device__ __forceinline__ float4 silu4_pre(float4 v) {
    if (FAST) {
        silu_pair_from_half(v.x, v.y);
        silu_pair_from_half(v.z, v.w);
    } else {
        v.x = silu_exact(v.x); v.y = silu_exact(v.y); v.z = silu_exact(v.z); v.w = silu_exact(v.w);
    }
    return v;
}
template <bool FAST>
__device__ __forceinline__ constexpr float tap_prescale() { return FAST ? 0.5f : 1.f; }
// d silu(x)/dx = s + x*s*(1-s), s = sigmoid(x)
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float dsilu(float x) {
    float s = sigmoidf_(x);
    return s * fmaf(x, 1.f - s, 1.f);
}
// softplus with the reference threshold (selective_scan_fwd_kernel.cuh:153-156)
__device__ __forceinline__ float softplus20(float x) { return x <= 20.f ? log1pf(__expf(x)) : x; }

template <typename T> struct is_fast : std::false_type {};
template <> struct is_fast<bf16> : std::true_type {};

__device__ __forceinline__ float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 fma4(float4 a, float4 b, float4 c) {
    return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 max4(float4 a, float4 b) {
    return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}
__device__ __forceinline__ float4 scale4(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Depthwise conv taps of 4 consecutive channels, transposed to one float4 per tap.
// conv_w is (dim, 4): out[t] = bias + sum_k w[k] * x[t-3+k]  (causal) -- oracle/fastvim_oracle.py
// causal_conv1d_oracle; the b-direction applies the same taps to the time-reversed sequence:
// out_b[t] = bias_b + sum_k w_b[k] * x[t+3-k].
struct Taps {
    float4 w[4];
    float4 b;
};
// pre: multiplies taps and bias (0.5 on the fast-SiLU path: the conv then yields h = x/2 directly)
__device__ __forceinline__ Taps load_taps(const float* cw, const float* cb, int D, int dir, int d0, float pre = 1.f) {
    Taps t;
    const float* p = cw + ((int64_t)dir * D + d0) * 4;
    float4 c0 = ld4(p), c1 = ld4(p + 4), c2 = ld4(p + 8), c3 = ld4(p + 12);
    t.w[0] = make_float4(c0.x, c1.x, c2.x, c3.x);
    t.w[1] = make_float4(c0.y, c1.y, c2.y, c3.y);
    t.w[2] = make_float4(c0.z, c1.z, c2.z, c3.z);
    t.w[3] = make_float4(c0.w, c1.w, c2.w, c3.w);
    t.b = cb ? ld4(cb + (int64_t)dir * D + d0) : zero4();
    if (pre != 1.f) {
#pragma unroll
        for (int k = 0; k < 4; ++k) t.w[k] = scale4(t.w[k], pre);
        t.b = scale4(t.b, pre);
    }
    return t;
}


// ---- async global -> shared staging (cp.async; LDGSTS in SASS) ------------------------
// The full-resolution kernels stage whole token rows (all channels handled by the CTA) into
// shared memory with 16-byte cp.async: the loads are bulk, coalesced and cost no registers, so
// the memory-level parallelism does not depend on occupancy.  Out-of-range rows are zero-filled
// (src-size 0), which is exactly the conv's zero padding.
__device__ __forceinline__ void cp_async16(void* smem, const void* gptr, bool valid) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gptr), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gptr, bool valid) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa), "l"(gptr), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Row table: rowtab[i] = memory token row of sequence position t_lo + i, or -1 outside [0, L).
// Filled once per tile by the first threads (the only integer divisions of the kernel), then the
// staging, compute and store loops index it instead of re-deriving rows.
__device__ __forceinline__ void fill_rowtab(const Geom& g, int t_lo, int nrows, int* rowtab) {
    for (int i = threadIdx.x; i < nrows; i += blockDim.x) {
        const int t = t_lo + i;
        rowtab[i] = (t >= 0 && t < g.L) ? (int)seq_to_row(g, t) : -1;
    }
}
// Stages rows rowtab[0..nrows) of one image (g.D channels, token-major) into dst[nrows][g.D];
// rows marked -1 are zero-filled.  vec16: rows, strides and base are 16-byte aligned (host-checked).
// Requires blockDim.x >= g.D / 4 (true for every caller: one thread per 4 channels).
template <typename T>
__device__ __forceinline__ void stage_rows(const Geom& g, const T* __restrict__ xb, int64_t ldx,
                                           const int* rowtab, int nrows, T* dst, bool vec16) {
    if (vec16) {
        constexpr int E = 16 / (int)sizeof(T);
        const int vpr = g.D / E;
        const int rsub = threadIdx.x / vpr, c = threadIdx.x - rsub * vpr, rpp = blockDim.x / vpr;
        if (rsub < rpp)
            for (int r = rsub; r < nrows; r += rpp) {
                const int row = rowtab[r];
                cp_async16(dst + (size_t)r * g.D + c * E, row >= 0 ? xb + (int64_t)row * ldx + c * E : xb, row >= 0);
            }
    } else {
        constexpr int E = 8 / (int)sizeof(T);
        const int vpr = g.D / E;
        for (int v = threadIdx.x; v < vpr; v += blockDim.x)
            for (int r = 0; r < nrows; ++r) {
                const int row = rowtab[r];
                cp_async8(dst + (size_t)r * g.D + v * E, row >= 0 ? xb + (int64_t)row * ldx + v * E : xb, row >= 0);
            }
    }
}
template <typename T>
static inline bool rows_vec16(int D, const void* p, int64_t ld, int64_t bs) {
    constexpr int E = 16 / (int)sizeof(T);
    return D % E == 0 && ld % E == 0 && bs % E == 0 && ((uintptr_t)p % 16) == 0;
}

// ---- TMA bulk copies (cp.async.bulk; UBLKCP in SASS) completing on an mbarrier ----------
// Whole token rows (all channels of a CTA, contiguous in memory) are moved global -> shared by
// the TMA engine: one lane issues one row, nobody spends issue slots on address arithmetic, and
// the consumer warps only wait on the mbarrier's phase parity.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------
// Kernels of the inference chain (in_proj GEMM -> fv_block_fwd -> out_proj GEMM (+ add + norm) -> ...) are launched with the
// programmatic-stream-serialization attribute: a kernel's CTAs may become resident and run their prologue (barrier init,
// TMEM allocation, tables, parameter loads) while the previous kernel drains; pdl_wait() blocks until the previous grid has
// completed and its memory is visible, and must precede the first read of its output AND the first global write.
// pdl_trigger() lets the NEXT kernel start launching (it fires once every CTA of this grid has called it or exited).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();      // core.cu: FASTVIM_PDL != "0" -- the FastVim-T inference chain kernels
bool pdl_all_enabled();  // core.cu: FASTVIM_PDL_ALL == "1" -- every other kernel (FV_LAUNCH_PDL)
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl_if(bool on, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                        Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = on ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    return launch_pdl_if(pdl_enabled(), kern, grid, block, smem, st, args...);
}

// kernel<<<grid, block, smem, st>>>(args...) with the PDL attribute; the kernel name goes in parentheses (template commas)
#define FV_LAUNCH_PDL(kern, grid, block, smem, st, ...) \
    ((void)::fv::launch_pdl_if(::fv::pdl_all_enabled(), kern, dim3(grid), dim3(block), (size_t)(smem), (cudaStream_t)(st), __VA_ARGS__))
// second attribute for cudaLaunchKernelEx launches that already carry a cluster dimension
static inline void pdl_attr(cudaLaunchAttribute* at) {
    at->id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at->val.programmaticStreamSerializationAllowed = pdl_all_enabled() ? 1 : 0;
}

// Both conv directions at one token from a 7-row window w[k] = x[t-3+k]:
//   f: silu(b_f + sum_k w_f[k] x[t-3+k])      b: silu(b_b + sum_k w_b[k] x[t+3-k])
template <bool FAST>
__device__ __forceinline__ void conv_both(const float4 (&w)[7], const Taps& tf, const Taps& tb, float4& xf,
                                          float4& xb_) {
    float4 af = tf.b, ab = tb.b;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        af = fma4(tf.w[k], w[k], af);
        ab = fma4(tb.w[k], w[6 - k], ab);
    }
    xf = silu4<FAST>(af);
    xb_ = silu4<FAST>(ab);
}
// Same with taps loaded through load_taps(..., tap_prescale<FAST>()), window given as 7 values.
template <bool FAST>
__device__ __forceinline__ void conv_both_pre(const float4& r0, const float4& r1, const float4& r2, const float4& r3,
                                              const float4& r4, const float4& r5, const float4& r6, const Taps& tf,
                                              const Taps& tb, float4& xf, float4& xb_) {
    float4 af = fma4(tf.w[0], r0, tf.b), ab = fma4(tb.w[0], r6, tb.b);
    af = fma4(tf.w[1], r1, af); ab = fma4(tb.w[1], r5, ab);
    af = fma4(tf.w[2], r2, af); ab = fma4(tb.w[2], r4, ab);
    af = fma4(tf.w[3], r3, af); ab = fma4(tb.w[3], r3, ab);
    xf = silu4_pre<FAST>(af);
    xb_ = silu4_pre<FAST>(ab);
}
// pooled position of sequence position t
__device__ __forceinline__ int seq_to_pooled(const Geom& g, int t) {
    if (g.inner == 1) return t / g.pool;
    const int q = t / g.inner, i = t - q * g.inner;
    return (q / g.pool) * g.inner + i;
}

}  // namespace fv
