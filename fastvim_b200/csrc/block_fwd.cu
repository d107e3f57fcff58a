// K-fused -- the whole SSM-block interior of one image in ONE kernel, x resident in shared memory.
//
// What it replaces in the reference (paths relative to /root/reference), i.e. everything between the
// in_proj output and the out_proj input of mamba_ssm/modules/mamba_simple_faster.py:269-453:
//   x.flip, 2x causal_conv1d_fn, 2x mean pool            :272-297
//   2x x_proj, 2x dt_proj                                :312-337, 368-394
//   2x selective_scan_fn(delta_softplus=True)            :343-354, 397-410  (fwd_kernel.cuh:67-303)
//   repeat_interleave + D skip, flip/add//2, LayerNorm, * silu(z)   :356-358, 412-416, 434-453
// and, in this repo, the four-launch path fv_conv_pool_fwd -> x_proj GEMM -> fv_scan_fwd -> fv_gate_fwd.
//
// B200 mapping.  At 224^2 one image's x is (196 tokens x 384 channels) bf16 = 147 KB: it fits in the 227 KB
// of shared memory of ONE SM.  So a persistent CTA (768 threads, one per SM) owns a whole image:
//   load   x token rows -> smem slab with TMA bulk copies (one 768-byte row per copy, any token permutation:
//          the odd-layer rotation of models/fastvim.py:192-210 is just the source row of each copy);
//   pass 1 depthwise conv (both directions, sliding 7-token register window, packed f32x2 FMAs) + SiLU +
//          mean pool -> pooled u (bf16, smem); the D-skip term w = (D_f xc_f + D_b xc_b)/2 is written IN PLACE
//          over x (bf16), so nothing is recomputed later;
//   x_proj [dt|B|C] = u W_x^T and dt_proj  delta_pre = dt W_dt^T: tiny (28 x 44 x 384) GEMMs on the tensor
//          cores (mma.sync bf16, fp32 accumulate), operands from smem / L2, results in smem;
//   scan   one thread per (channel, direction): 16 fp32 states in registers, softplus + exp2 recurrence over
//          the pooled rows, both directions concurrently (different warps), output s in smem;
//   gate   v = w + (s_f + s_b)/2, LayerNorm over d_inner (per-token partial sums through smem), * silu(z)
//          with z streamed from HBM one tile ahead through registers, y written once.
// While the gate pass walks the slab tile by tile, the producer warp refills the freed rows with the NEXT
// image's x (TMA, completion on one mbarrier), so the load of image i+1 overlaps the epilogue of image i.
// HBM traffic = x + z read once, y written once: the algorithmic 3*B*L*D*s.
// Restrictions (fv_block_fwd_supported): bf16, plain (outer, pool, 1) geometry, mean pooling, d_state 16,
// dim <= 384 and (L+6)*dim*2 + pooled buffers <= 227 KB; everything else uses the four-launch path.
#include <cuda.h>

#include "common.cuh"

namespace fv {

int sm_count();
int check_geom(const fv_geom* g, const char* who);

constexpr int BK_THREADS = 768;
constexpr int BK_WARPS = BK_THREADS / 32;
constexpr int BK_GT = 28;     // tokens per gate-pass tile: 7 slots x 4 tokens
constexpr int BK_NSTATE = 16;

struct BlockArgs {
    Geom g;
    const bf16* x;
    const bf16* z;
    int64_t ldxz, xzbs;
    const float* cw;
    const float* cb;
    const bf16* xw;   // (2, ncols, dim)  x_proj weights
    const bf16* dtw;  // (2, dim, R)      dt_proj weights
    const float* dtb;
    const float* A;
    int a_is_log;
    const float* Dskip;
    const float* lnw;
    const float* lnb;
    float eps, scale;  // scale = scaling_factor / pool
    bf16* y;
    int64_t ldy, ybs;
    bf16* u_out;      // (2, B, Lp, dim) or null   (saved for backward)
    bf16* xdbl_out;   // (2, B*Lp, ncols) or null
    float* s_out;     // (2, B, Lp, dim) or null
    int R, ncols, xld, uld, pld;
    int off_u, off_s, off_xdbl, off_stat, off_bar;  // byte offsets into dynamic smem (slab at 0)
};

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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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
// two exp2 with ONE MUFU op (ex2.approx.f16x2): halves the SFU load of the recurrence.  Relative error of
// the decay factor <= 2^-11, four times below the bf16 rounding of the activations it multiplies.
__device__ __forceinline__ float2 bk_ex2_pair_f16(float2 x) {
    __half2 h = __floats2half2_rn(x.x, x.y);
    uint32_t hi = *reinterpret_cast<uint32_t*>(&h), ho;
    asm("ex2.approx.f16x2 %0, %1;" : "=r"(ho) : "r"(hi));
    return __half22float2(*reinterpret_cast<__half2*>(&ho));
}

// memory token row of sequence position t (plain geometry: inner == 1)
template <int POOL_T>
__device__ __forceinline__ int64_t bk_row(const Geom& g, int t) {
    const int P = POOL_T ? POOL_T : g.pool;
    const int o = t / P, p = t - o * P;
    return o * g.so + p * g.sp;
}

// producer warp: TMA bulk copies of tokens [t_lo, t_hi) of image b into the slab
template <int POOL_T>
__device__ __forceinline__ void bk_issue_rows(const BlockArgs& a, int b, int t_lo, int t_hi, unsigned char* slab,
                                              uint64_t* bar, bool arm) {
    const int lane = threadIdx.x & 31;
    const uint32_t rowB = (uint32_t)a.g.D * 2u;
    if (arm && lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)a.g.L * rowB);
    __syncwarp();
    const bf16* xb = a.x + (int64_t)b * a.xzbs;
    for (int t = t_lo + lane; t < t_hi; t += 32)
        bulk_g2s(slab + (size_t)(t + 3) * rowB, xb + bk_row<POOL_T>(a.g, t) * a.ldxz, rowB, bar);
}

template <int POOL_T, bool NORM, bool EXP16>
__global__ void __launch_bounds__(BK_THREADS, 1) block_fwd_kernel(const BlockArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int N = BK_NSTATE;
    const Geom& g = a.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = g.D, L = g.L, outer = g.outer;
    const int P = POOL_T ? POOL_T : g.pool;
    const int R = a.R;
    const uint32_t rowB = (uint32_t)D * 2u;

    unsigned char* slab = smem;                                      // (L + 6) token rows of D bf16; token t at row t+3
    bf16* ubuf = reinterpret_cast<bf16*>(smem + a.off_u);            // [2][outer][uld]
    float2* psum = reinterpret_cast<float2*>(smem + a.off_u);        // overlay (gate pass): [BK_GT][pld]
    float* sbuf = reinterpret_cast<float*>(smem + a.off_s);          // [2][outer][D]: delta_pre, then s
    float* xd = reinterpret_cast<float*>(smem + a.off_xdbl);         // [2][outer][xld]
    float2* stat = reinterpret_cast<float2*>(smem + a.off_stat);     // [BK_GT]
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + a.off_bar);

    // ---- one-time init: zero halo pad rows, mbarrier, first image's loads
    for (int i = tid; i < (int)(3 * rowB / 4); i += BK_THREADS) {
        reinterpret_cast<uint32_t*>(slab)[i] = 0u;
        reinterpret_cast<uint32_t*>(slab + (size_t)(L + 3) * rowB)[i] = 0u;
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    int img = blockIdx.x;
    if (warp == BK_WARPS - 1 && img < g.B) bk_issue_rows<POOL_T>(a, img, 0, L, slab, bar, true);

    const int ntile = (L + BK_GT - 1) / BK_GT;
    const int half_d = D >> 1, quart_d = D >> 2;

    // ---- pass-1 mapping: (channel pair, quarter of the pooled rows)
    const bool p1_live = tid < half_d * 4;
    const int p1_q = tid / half_d, p1_c = tid - p1_q * half_d;
    const int rpq = (outer + 3) >> 2;
    const int r_begin = min(outer, p1_q * rpq), r_end = min(outer, r_begin + rpq);
    // ---- gate mapping: (channel quad, slot) -> 4 consecutive tokens of a 28-token tile
    const int g_q = tid / quart_d, g_c = tid - g_q * quart_d;
    const bool g_live = g_q < 7;
    const int gd0 = g_c * 4;

    for (uint32_t it_img = 0; img < g.B; img += gridDim.x, ++it_img) {
        const int img_next = img + gridDim.x;
        const bool has_next = img_next < g.B;
        mbar_wait(bar, it_img & 1u);

        // ================= pass 1: conv (both directions) + SiLU + mean pool, w in place ===============
        uint32_t hl0 = 0, hl1 = 0, hl2 = 0;
        const unsigned char* my = slab + (size_t)p1_c * 4;   // this thread's channel pair, token -3
        if (p1_live && r_begin < r_end) {
            const int t0 = r_begin * P;
            hl0 = *reinterpret_cast<const uint32_t*>(my + (size_t)(t0 + 0) * rowB);
            hl1 = *reinterpret_cast<const uint32_t*>(my + (size_t)(t0 + 1) * rowB);
            hl2 = *reinterpret_cast<const uint32_t*>(my + (size_t)(t0 + 2) * rowB);
        }
        for (int i = tid; i < 2 * outer * a.xld; i += BK_THREADS) xd[i] = 0.f;
        __syncthreads();  // S1: left halos are in registers, nobody has overwritten x yet
        uint32_t dfr0 = 0, dfr1 = 0, dfr2 = 0;  // w of the first 3 tokens of the segment: stored after S2
        if (p1_live && r_begin < r_end) {
            const int d0 = p1_c * 2;
            float2 wf[4], wb[4], bf_, bb_, Df2, Db2;
            {
                const float4 f0 = ld4(a.cw + (int64_t)d0 * 4), f1 = ld4(a.cw + (int64_t)d0 * 4 + 4);
                const float4 b0 = ld4(a.cw + ((int64_t)D + d0) * 4), b1 = ld4(a.cw + ((int64_t)D + d0) * 4 + 4);
                wf[0] = make_float2(0.5f * f0.x, 0.5f * f1.x); wf[1] = make_float2(0.5f * f0.y, 0.5f * f1.y);
                wf[2] = make_float2(0.5f * f0.z, 0.5f * f1.z); wf[3] = make_float2(0.5f * f0.w, 0.5f * f1.w);
                wb[0] = make_float2(0.5f * b0.x, 0.5f * b1.x); wb[1] = make_float2(0.5f * b0.y, 0.5f * b1.y);
                wb[2] = make_float2(0.5f * b0.z, 0.5f * b1.z); wb[3] = make_float2(0.5f * b0.w, 0.5f * b1.w);
                bf_ = a.cb ? make_float2(0.5f * a.cb[d0], 0.5f * a.cb[d0 + 1]) : make_float2(0.f, 0.f);
                bb_ = a.cb ? make_float2(0.5f * a.cb[D + d0], 0.5f * a.cb[D + d0 + 1]) : make_float2(0.f, 0.f);
                Df2 = make_float2(0.5f * a.Dskip[d0], 0.5f * a.Dskip[d0 + 1]);
                Db2 = make_float2(0.5f * a.Dskip[D + d0], 0.5f * a.Dskip[D + d0 + 1]);
            }
            const float2 sc2 = make_float2(a.scale, a.scale);
            float2 win[7];
            win[0] = unpack2(hl0); win[1] = unpack2(hl1); win[2] = unpack2(hl2);
            {
                const unsigned char* p0 = my + (size_t)(r_begin * P + 3) * rowB;
                win[3] = unpack2(*reinterpret_cast<const uint32_t*>(p0));
                win[4] = unpack2(*reinterpret_cast<const uint32_t*>(p0 + rowB));
                win[5] = unpack2(*reinterpret_cast<const uint32_t*>(p0 + 2 * (size_t)rowB));
            }
            for (int r = r_begin; r < r_end; ++r) {
                float2 sumf = make_float2(0.f, 0.f), sumb = sumf;
                const bool first_row = r == r_begin;
                unsigned char* tok = slab + (size_t)(r * P + 3) * rowB + (size_t)p1_c * 4;  // token r*P
#define BK_TOKEN(C_, X_)                                                                                   \
    {                                                                                                      \
        X_(6) = unpack2(*reinterpret_cast<const uint32_t*>(tok + (size_t)((C_) + 3) * rowB));             \
        float2 af = __ffma2_rn(wf[0], X_(0), bf_), ab = __ffma2_rn(wb[0], X_(6), bb_);                     \
        af = __ffma2_rn(wf[1], X_(1), af); ab = __ffma2_rn(wb[1], X_(5), ab);                              \
        af = __ffma2_rn(wf[2], X_(2), af); ab = __ffma2_rn(wb[2], X_(4), ab);                              \
        af = __ffma2_rn(wf[3], X_(3), af); ab = __ffma2_rn(wb[3], X_(3), ab);                              \
        silu_pair_from_half(af.x, af.y);                                                                   \
        silu_pair_from_half(ab.x, ab.y);                                                                   \
        sumf = __fadd2_rn(sumf, af);                                                                       \
        sumb = __fadd2_rn(sumb, ab);                                                                       \
        const uint32_t wv = pack2(__ffma2_rn(Db2, ab, __fmul2_rn(Df2, af)));                               \
        if (first_row && (C_) < 3) {                                                                       \
            if ((C_) == 0) dfr0 = wv; else if ((C_) == 1) dfr1 = wv; else dfr2 = wv;                        \
        } else {                                                                                           \
            *reinterpret_cast<uint32_t*>(tok + (size_t)(C_) * rowB) = wv;                                  \
        }                                                                                                  \
    }
                if (POOL_T) {
#define BK_XU(k_) win[(c + (k_)) % 7]
#pragma unroll
                    for (int c = 0; c < (POOL_T ? POOL_T : 1); ++c) BK_TOKEN(c, BK_XU)
#undef BK_XU
                } else {
#define BK_XS(k_) win[(k_)]
                    for (int c = 0; c < P; ++c) {
                        BK_TOKEN(c, BK_XS)
#pragma unroll
                        for (int k = 0; k < 6; ++k) win[k] = win[k + 1];
                    }
#undef BK_XS
                }
#undef BK_TOKEN
                bf16* urow = ubuf + (size_t)r * a.uld + d0;
                *reinterpret_cast<uint32_t*>(urow) = pack2(__fmul2_rn(sumf, sc2));
                *reinterpret_cast<uint32_t*>(urow + (size_t)outer * a.uld) = pack2(__fmul2_rn(sumb, sc2));
            }
        }
        __syncthreads();  // S2: u complete; every right halo has been read
        if (p1_live && r_begin < r_end) {
            unsigned char* tok = slab + (size_t)(r_begin * P + 3) * rowB + (size_t)p1_c * 4;
            *reinterpret_cast<uint32_t*>(tok) = dfr0;
            *reinterpret_cast<uint32_t*>(tok + rowB) = dfr1;
            *reinterpret_cast<uint32_t*>(tok + 2 * (size_t)rowB) = dfr2;
        }

        // ================= x_proj on tensor cores: xd[dir][j][c] = sum_d u[dir][j][d] W_x[dir][c][d] ===============
        {
            const int gq = lane >> 2, tq = lane & 3;
            const int n_mt = (outer + 15) >> 4, n_nt = (a.ncols + 7) >> 3, kper = D >> 1;
            const int nitem = 2 * n_mt * n_nt * 2;
            for (int item = warp; item < nitem; item += BK_WARPS) {
                const int kh = item & 1;
                int rest = item >> 1;
                const int nt = rest % n_nt;
                rest /= n_nt;
                const int mt = rest % n_mt, dir = rest / n_mt;
                const int row0 = mt * 16 + gq, row1 = row0 + 8;
                const bf16* ua0 = ubuf + ((size_t)dir * outer + min(row0, outer - 1)) * a.uld + kh * kper + tq * 2;
                const bf16* ua1 = ubuf + ((size_t)dir * outer + min(row1, outer - 1)) * a.uld + kh * kper + tq * 2;
                const int n = nt * 8 + gq;
                const bool nv = n < a.ncols;
                const bf16* wp = a.xw + ((size_t)dir * a.ncols + (nv ? n : 0)) * D + kh * kper + tq * 2;
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
                for (int k0 = 0; k0 < kper; k0 += 16) {
                    const uint32_t a0 = *reinterpret_cast<const uint32_t*>(ua0 + k0);
                    const uint32_t a1 = *reinterpret_cast<const uint32_t*>(ua1 + k0);
                    const uint32_t a2 = *reinterpret_cast<const uint32_t*>(ua0 + k0 + 8);
                    const uint32_t a3 = *reinterpret_cast<const uint32_t*>(ua1 + k0 + 8);
                    uint32_t b0 = 0, b1 = 0;
                    if (nv) {
                        b0 = __ldg(reinterpret_cast<const unsigned int*>(wp + k0));
                        b1 = __ldg(reinterpret_cast<const unsigned int*>(wp + k0 + 8));
                    }
                    mma16816(acc, a0, a1, a2, a3, b0, b1);
                }
                const int c0 = nt * 8 + tq * 2;
                float* o0 = xd + ((size_t)dir * outer + row0) * a.xld + c0;
                float* o1 = xd + ((size_t)dir * outer + row1) * a.xld + c0;
                if (row0 < outer) {
                    if (c0 < a.ncols) atomicAdd(o0, acc[0]);
                    if (c0 + 1 < a.ncols) atomicAdd(o0 + 1, acc[1]);
                }
                if (row1 < outer) {
                    if (c0 < a.ncols) atomicAdd(o1, acc[2]);
                    if (c0 + 1 < a.ncols) atomicAdd(o1 + 1, acc[3]);
                }
            }
        }
        __syncthreads();  // S3: xd complete

        // ================= dt_proj on tensor cores: sbuf[dir][j][d] = sum_r xd[dir][j][r] W_dt[dir][d][r] ===========
        {
            const int gq = lane >> 2, tq = lane & 3;
            const int n_mt = (outer + 15) >> 4, n_nt = D >> 3;
            const int nitem = 2 * n_mt * n_nt;
            for (int item = warp; item < nitem; item += BK_WARPS) {
                const int nt = item % n_nt;
                const int rest = item / n_nt;
                const int mt = rest % n_mt, dir = rest / n_mt;
                const int row0 = mt * 16 + gq, row1 = row0 + 8;
                const float* xa0 = xd + ((size_t)dir * outer + min(row0, outer - 1)) * a.xld;
                const float* xa1 = xd + ((size_t)dir * outer + min(row1, outer - 1)) * a.xld;
                const bf16* wp = a.dtw + ((size_t)dir * D + nt * 8 + gq) * R;
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                for (int k0 = 0; k0 < R; k0 += 16) {
                    const int ka = k0 + tq * 2, kb = ka + 8;
                    uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0, b0 = 0, b1 = 0;
                    if (ka < R) {
                        a0 = pack2(xa0[ka], xa0[ka + 1]);
                        a1 = pack2(xa1[ka], xa1[ka + 1]);
                        b0 = __ldg(reinterpret_cast<const unsigned int*>(wp + ka));
                    }
                    if (kb < R) {
                        a2 = pack2(xa0[kb], xa0[kb + 1]);
                        a3 = pack2(xa1[kb], xa1[kb + 1]);
                        b1 = __ldg(reinterpret_cast<const unsigned int*>(wp + kb));
                    }
                    mma16816(acc, a0, a1, a2, a3, b0, b1);
                }
                const int c0 = nt * 8 + tq * 2;
                if (row0 < outer)
                    *reinterpret_cast<float2*>(sbuf + ((size_t)dir * outer + row0) * D + c0) = make_float2(acc[0], acc[1]);
                if (row1 < outer)
                    *reinterpret_cast<float2*>(sbuf + ((size_t)dir * outer + row1) * D + c0) = make_float2(acc[2], acc[3]);
            }
            if (a.xdbl_out) {  // saved for backward: (2, B*Lp, ncols) bf16
                for (int i = tid; i < 2 * outer * a.ncols; i += BK_THREADS) {
                    const int c = i % a.ncols, rj = i / a.ncols;  // rj = dir*outer + j
                    const int dir = rj / outer, j = rj - dir * outer;
                    a.xdbl_out[(((int64_t)dir * g.B + img) * outer + j) * a.ncols + c] =
                        __float2bfloat16_rn(xd[(size_t)rj * a.xld + c]);
                }
            }
        }
        // z of the first gate tile: issued now, lands while the scan runs
        uint2 zc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) zc[k] = make_uint2(0u, 0u);
        const bf16* zb = a.z + (int64_t)img * a.xzbs + gd0;
        if (g_live) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int t = g_q * 4 + k;
                if (t < L) zc[k] = __ldg(reinterpret_cast<const uint2*>(zb + bk_row<POOL_T>(g, t) * a.ldxz));
            }
        }
        __syncthreads();  // S4: delta_pre complete

        // ================= bidirectional selective scan over the pooled rows =======================
        if (tid < 2 * D) {
            const int dir = tid >= D ? 1 : 0, d = tid - dir * D;
            constexpr float LOG2E = 1.4426950408889634f;
            float A2[N], h[N];
            {
                const float4* Ap = reinterpret_cast<const float4*>(a.A + ((int64_t)dir * D + d) * N);
#pragma unroll
                for (int n4 = 0; n4 < N / 4; ++n4) {
                    const float4 v = __ldg(Ap + n4);
                    A2[4 * n4 + 0] = v.x; A2[4 * n4 + 1] = v.y; A2[4 * n4 + 2] = v.z; A2[4 * n4 + 3] = v.w;
                }
#pragma unroll
                for (int n = 0; n < N; ++n) {
                    A2[n] = (a.a_is_log ? -__expf(A2[n]) : A2[n]) * LOG2E;
                    h[n] = 0.f;
                }
            }
            const float bias = a.dtb[(int64_t)dir * D + d];
            const int64_t gplane = ((int64_t)dir * g.B + img) * outer;
            for (int step = 0; step < outer; ++step) {
                const int j = dir ? outer - 1 - step : step;
                float* sp = sbuf + ((size_t)dir * outer + j) * D + d;
                const bf16 ub = ubuf[((size_t)dir * outer + j) * a.uld + d];
                const float delta = bk_softplus(*sp + bias);
                const float du = delta * __bfloat162float(ub);
                const float* row = xd + ((size_t)dir * outer + j) * a.xld + R;
                float2 y2 = make_float2(0.f, 0.f);
                const float2 dl2 = make_float2(delta, delta), du2 = make_float2(du, du);
#pragma unroll
                for (int n = 0; n < N; n += 4) {
                    const float4 Bq = *reinterpret_cast<const float4*>(row + n);
                    const float4 Cq = *reinterpret_cast<const float4*>(row + N + n);
                    float2 e0 = __fmul2_rn(dl2, make_float2(A2[n], A2[n + 1]));
                    float2 e1 = __fmul2_rn(dl2, make_float2(A2[n + 2], A2[n + 3]));
                    if (EXP16) {
                        e0 = bk_ex2_pair_f16(e0);
                        e1 = bk_ex2_pair_f16(e1);
                    } else {
                        e0 = make_float2(bk_ex2(e0.x), bk_ex2(e0.y));
                        e1 = make_float2(bk_ex2(e1.x), bk_ex2(e1.y));
                    }
                    float2 h0 = __ffma2_rn(e0, make_float2(h[n], h[n + 1]), __fmul2_rn(du2, make_float2(Bq.x, Bq.y)));
                    float2 h1 = __ffma2_rn(e1, make_float2(h[n + 2], h[n + 3]), __fmul2_rn(du2, make_float2(Bq.z, Bq.w)));
                    h[n] = h0.x; h[n + 1] = h0.y; h[n + 2] = h1.x; h[n + 3] = h1.y;
                    y2 = __ffma2_rn(h0, make_float2(Cq.x, Cq.y), y2);
                    y2 = __ffma2_rn(h1, make_float2(Cq.z, Cq.w), y2);
                }
                const float yv = y2.x + y2.y;
                *sp = yv;
                if (a.s_out) a.s_out[(gplane + j) * D + d] = yv;
                if (a.u_out) a.u_out[(gplane + j) * D + d] = ub;
            }
        }
        __syncthreads();  // S5: s complete; u is dead (psum may overlay it)

        // ================= gate pass: v = w + (s_f + s_b)/2, LayerNorm, * silu(z), store y ===========
        float4 gam = make_float4(1.f, 1.f, 1.f, 1.f), bet = zero4();
        if (NORM && g_live) {
            gam = ld4(a.lnw + gd0);
            if (a.lnb) bet = ld4(a.lnb + gd0);
        }
        const float invD = 1.f / (float)D;
        bf16* yb = a.y + (int64_t)img * a.ybs + gd0;
        for (int tile = 0; tile < ntile; ++tile) {
            const int tl0 = g_q * 4, t0 = tile * BK_GT + tl0;
            // z of the next tile (one tile ahead, through registers)
            uint2 zn[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) zn[k] = make_uint2(0u, 0u);
            if (g_live && tile + 1 < ntile) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int t = t0 + BK_GT + k;
                    if (t < L) zn[k] = __ldg(reinterpret_cast<const uint2*>(zb + bk_row<POOL_T>(g, t) * a.ldxz));
                }
            }
            float4 v[4];
            if (g_live) {
                int jprev = -1;
                float4 ss = zero4();
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int t = t0 + k;
                    v[k] = zero4();
                    if (t < L) {
                        const int j = t / P;
                        if (j != jprev) {
                            const float4 sf = *reinterpret_cast<const float4*>(sbuf + (size_t)j * D + gd0);
                            const float4 sb = *reinterpret_cast<const float4*>(sbuf + ((size_t)outer + j) * D + gd0);
                            ss = scale4(sf + sb, 0.5f);
                            jprev = j;
                        }
                        const uint2 wv = *reinterpret_cast<const uint2*>(slab + (size_t)(t + 3) * rowB + (size_t)gd0 * 2);
                        const float2 w01 = unpack2(wv.x), w23 = unpack2(wv.y);
                        v[k] = make_float4(w01.x + ss.x, w01.y + ss.y, w23.x + ss.z, w23.y + ss.w);
                        if (NORM)
                            psum[(size_t)(tl0 + k) * a.pld + g_c] =
                                make_float2((v[k].x + v[k].y) + (v[k].z + v[k].w),
                                            fmaf(v[k].x, v[k].x, fmaf(v[k].y, v[k].y, fmaf(v[k].z, v[k].z, v[k].w * v[k].w))));
                    }
                }
            }
            if (NORM) {
                __syncthreads();  // GA: partial sums of this tile are in smem; the previous tile's rows are free
                if (tid < BK_GT * 8) {
                    const int tl = tid >> 3, sub = tid & 7;
                    float sum = 0.f, sq = 0.f;
                    if (tile * BK_GT + tl < L)
                        for (int k = sub; k < quart_d; k += 8) {
                            const float2 q = psum[(size_t)tl * a.pld + k];
                            sum += q.x;
                            sq += q.y;
                        }
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) {
                        sum += __shfl_xor_sync(0xffffffffu, sum, o);
                        sq += __shfl_xor_sync(0xffffffffu, sq, o);
                    }
                    if (sub == 0) {
                        const float mean = sum * invD;
                        stat[tl] = make_float2(mean, rsqrtf(fmaxf(sq * invD - mean * mean, 0.f) + a.eps));
                    }
                }
            } else {
                __syncthreads();  // GA (no-norm form): only orders the slab refill below
            }
            if (warp == BK_WARPS - 1 && has_next) {  // every w row of this tile has been read: refill with the next image
                fence_proxy_async();
                bk_issue_rows<POOL_T>(a, img_next, tile * BK_GT, min(L, (tile + 1) * BK_GT), slab, bar, tile == 0);
            }
            if (NORM) __syncthreads();  // GB: statistics ready
            if (g_live) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int t = t0 + k;
                    if (t < L) {
                        float4 o = v[k];
                        if (NORM) {
                            const float2 ms = stat[tl0 + k];
                            const float4 gs = scale4(gam, ms.y);
                            o.x = fmaf(o.x - ms.x, gs.x, bet.x);
                            o.y = fmaf(o.y - ms.x, gs.y, bet.y);
                            o.z = fmaf(o.z - ms.x, gs.z, bet.z);
                            o.w = fmaf(o.w - ms.x, gs.w, bet.w);
                        }
                        const float2 z01 = unpack2(zc[k].x), z23 = unpack2(zc[k].y);
                        float4 zz = silu4<true>(make_float4(z01.x, z01.y, z23.x, z23.y));
                        uint2 pk;
                        pk.x = pack2(o.x * zz.x, o.y * zz.y);
                        pk.y = pack2(o.z * zz.z, o.w * zz.w);
                        *reinterpret_cast<uint2*>(yb + bk_row<POOL_T>(g, t) * a.ldy) = pk;
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) zc[k] = zn[k];
        }
    }
}

struct BlockPlan {
    int ok;
    size_t smem;
    int xld, uld, pld, off_u, off_s, off_xdbl, off_stat, off_bar;
};

static BlockPlan plan_block(const fv_geom* g, int dtype, int R, int N) {
    BlockPlan p;
    p.ok = 0;
    if (dtype != FV_BF16 || g->inner != 1 || N != BK_NSTATE) return p;
    const int D = g->dim, outer = g->outer, pool = g->pool;
    if (D % 32 != 0 || D > 384 || pool < 4 || R <= 0 || R % 4 != 0 || R > 64) return p;
    const int64_t L = (int64_t)outer * pool;
    const int ncols = R + 2 * N;
    p.xld = (ncols + 7) / 8 * 8;
    p.uld = D + 8;
    p.pld = D / 4 + 1;
    auto up = [](size_t v) { return (v + 127) / 128 * 128; };
    size_t off = up((size_t)(L + 6) * D * 2);
    p.off_u = (int)off;
    const size_t ub = (size_t)2 * outer * p.uld * 2, pb = (size_t)BK_GT * p.pld * 8;
    off = up(off + (ub > pb ? ub : pb));
    p.off_s = (int)off;
    off = up(off + (size_t)2 * outer * D * 4);
    p.off_xdbl = (int)off;
    off = up(off + (size_t)2 * outer * p.xld * 4);
    p.off_stat = (int)off;
    off = up(off + BK_GT * 8);
    p.off_bar = (int)off;
    off += 16;
    p.smem = off;
    if (off > 227 * 1024 || L * D * 2 >= (1 << 20)) return p;  // smem capacity; mbarrier tx-count range
    p.ok = 1;
    return p;
}

}  // namespace fv

extern "C" int fv_block_fwd_supported(const fv_geom* g, int dtype, int dt_rank, int dstate) {
    if (!g || g->batch <= 0 || g->dim <= 0 || g->outer <= 0 || g->pool <= 0) return 0;
    return fv::plan_block(g, dtype, dt_rank, dstate).ok;
}

extern "C" int fv_block_fwd(const fv_geom* g_, int dtype, const void* x, const void* z, int64_t ldxz,
                            int64_t xz_bstride, const float* conv_w, const float* conv_b, const void* xproj_w,
                            const void* dt_w, const float* dt_bias, const float* A, int a_is_log,
                            int dt_rank, int dstate, const float* Dskip, const float* ln_w, const float* ln_b,
                            float eps, float scale, int exp_mode, void* y, int64_t ldy, int64_t y_bstride,
                            void* u_out, void* xdbl_out, float* s_out, void* stream) {
    using namespace fv;
    if (int rc = check_geom(g_, "fv_block_fwd")) return rc;
    FV_REQUIRE(x && z && conv_w && xproj_w && dt_w && dt_bias && A && Dskip && y, "fv_block_fwd: null pointer");
    const BlockPlan p = plan_block(g_, dtype, dt_rank, dstate);
    FV_REQUIRE(p.ok, "fv_block_fwd: unsupported configuration (bf16, plain geometry, dim %% 32 == 0, dim <= 384, "
                     "d_state 16, dt_rank %% 4 == 0, slab must fit 227 KB); use the four-launch path");
    FV_REQUIRE(ldxz % 8 == 0 && xz_bstride % 8 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)z % 8) == 0,
               "fv_block_fwd: x rows must be 16-byte aligned (ldxz %lld)", (long long)ldxz);
    FV_REQUIRE(ldy % 4 == 0 && y_bstride % 4 == 0 && ((uintptr_t)y % 8) == 0, "fv_block_fwd: y rows must be 8-byte aligned");
    FV_REQUIRE(((uintptr_t)xproj_w % 4) == 0 && ((uintptr_t)dt_w % 4) == 0, "fv_block_fwd: weights must be 4-byte aligned");
    BlockArgs a;
    a.g = make_geom(g_);
    a.x = (const bf16*)x; a.z = (const bf16*)z; a.ldxz = ldxz; a.xzbs = xz_bstride;
    a.cw = conv_w; a.cb = conv_b; a.xw = (const bf16*)xproj_w; a.dtw = (const bf16*)dt_w; a.dtb = dt_bias;
    a.A = A; a.a_is_log = a_is_log; a.Dskip = Dskip; a.lnw = ln_w; a.lnb = ln_b; a.eps = eps;
    a.scale = scale / (float)g_->pool;
    a.y = (bf16*)y; a.ldy = ldy; a.ybs = y_bstride;
    a.u_out = (bf16*)u_out; a.xdbl_out = (bf16*)xdbl_out; a.s_out = s_out;
    a.R = dt_rank; a.ncols = dt_rank + 2 * dstate; a.xld = p.xld; a.uld = p.uld; a.pld = p.pld;
    a.off_u = p.off_u; a.off_s = p.off_s; a.off_xdbl = p.off_xdbl; a.off_stat = p.off_stat; a.off_bar = p.off_bar;

    void (*kern)(const BlockArgs);
    const bool e16 = exp_mode != 0;
#define FV_BK_PICK(P_)                                                                        \
    (ln_w ? (e16 ? block_fwd_kernel<P_, true, true> : block_fwd_kernel<P_, true, false>)      \
          : (e16 ? block_fwd_kernel<P_, false, true> : block_fwd_kernel<P_, false, false>))
    if (g_->pool == 14) kern = FV_BK_PICK(14);
    else kern = FV_BK_PICK(0);
#undef FV_BK_PICK
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    FV_REQUIRE(e == cudaSuccess, "fv_block_fwd: cudaFuncSetAttribute(%zu): %s", p.smem, cudaGetErrorString(e));
    const int grid = g_->batch < sm_count() ? g_->batch : sm_count();
    kern<<<grid, BK_THREADS, p.smem, (cudaStream_t)stream>>>(a);
    return finish_launch("block_fwd");
}
