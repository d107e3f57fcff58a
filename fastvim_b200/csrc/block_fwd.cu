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
//   load   x token rows -> smem slab with 16-byte cp.async (LDGSTS), issued by the compute warps themselves; any
//          token permutation: the odd-layer rotation of models/fastvim.py:192-210 is just the source row.
//          (Measured: TMA bulk copies cost ~54 cycles of serial engine time per contiguous 768-byte token row --
//          x and z interleave in the in_proj output, so rows cannot be merged -- 5 us per image; LDGSTS costs 8.)
//   pass 1 depthwise conv (both directions, sliding 7-token register window, packed f32x2 FMAs) + SiLU +
//          mean pool -> pooled u (bf16, smem); the D-skip term w = (D_f xc_f + D_b xc_b)/2 is written IN PLACE
//          over x (bf16), so nothing is recomputed later;
//   x_proj [dt|B|C] = u W_x^T: a tiny (28 x 44 x 384) GEMM on the tensor cores (mma.sync bf16, fp32
//          accumulate; one (direction, 8-column tile, k-half) item per warp), u from smem, W_x from L2;
//   scan   one thread per (channel, direction): dt_proj row + 16 fp32 states in registers, softplus + exp2
//          recurrence over the pooled rows, both directions concurrently (different warps), s in smem;
//   gate   one WARP per token (lane = 4-channel groups lane, lane+32, lane+64): v = w + (s_f + s_b)/2,
//          LayerNorm statistics by warp shuffles (no block barrier in the whole pass), * silu(z) with z
//          streamed from HBM one round ahead through registers, y written once.
// The 24 warps walk the slab round-robin (token = warp + 24*round); as soon as a warp holds its token's w row in
// registers it refills that row with the NEXT image's x (cp.async), so the load of image i+1 overlaps the
// epilogue of image i and the only wait is a cp.async.wait_all + barrier at the top of the next image.
// HBM traffic = x + z read once, y written once: the algorithmic 3*B*L*D*s.
// Restrictions (fv_block_fwd_supported): bf16, plain (outer, pool, 1) geometry, mean pooling, d_state 16,
// dim <= 384 and (L+6)*dim*2 + pooled buffers <= 227 KB; everything else uses the four-launch path.

#include <cstdlib>

#include "block_common.cuh"

namespace fv {

int sm_count();
int check_geom(const fv_geom* g, const char* who);

constexpr int BK_THREADS = 768;
constexpr int BK_WARPS = BK_THREADS / 32;
constexpr int BK_NWORK = BK_WARPS;      // gate pass: every warp owns one token per round
constexpr int BK_NCG = 3;               // 4-channel groups per lane in the gate pass: dim <= 384

struct BlockArgs {
    Geom g;
    const bf16* x;
    const bf16* z;
    int64_t ldxz, xzbs;
    const float* cw;
    const float* cb;
    const bf16* xw;   // (2, ncols, dim)  x_proj weights
    const uint4* xwp; // the same weights in MMA-fragment order (fv_block_pack_xproj), or null
    const float* dtw; // (2, dim, R)      dt_proj weights (fp32, held in registers by the scan threads)
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
    int R, ncols, xld, uld;
    int off_u, off_s, off_xdbl, off_tab;  // byte offsets into dynamic smem (slab at 0)
    int* done_flags;  // (B) or null: done_flags[img] = done_epoch once every y row of the image is written (release, gpu scope)
    int done_epoch;   //   -> a dependent kernel launched early (PDL) consumes images as they complete (fv_gemm_out_norm_flow)
};

template <int POOL_T, bool NORM, bool FULL, int RT>
__global__ void __launch_bounds__(BK_THREADS, 1) block_fwd_kernel(const BlockArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int N = BK_NSTATE;
    const Geom& g = a.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = g.D, L = g.L, outer = g.outer;
    const int P = POOL_T ? POOL_T : g.pool;
    const uint32_t rowB = (uint32_t)D * 2u;

    unsigned char* slab = smem;                                      // (L + 6) token rows of D bf16; token t at row t+3
    bf16* ubuf = reinterpret_cast<bf16*>(smem + a.off_u);            // [2][outer][uld]
    float* sbuf = reinterpret_cast<float*>(smem + a.off_s);          // [2][outer][D]: scan output s per direction
    float* xd = reinterpret_cast<float*>(smem + a.off_xdbl);         // [2][outer][xld]
    uint32_t* ztab = reinterpret_cast<uint32_t*>(smem + a.off_tab);  // [L] element offset of token t's row in x / z
    uint32_t* ytab = ztab + L;                                       // [L] ... in y

    // ---- one-time init: zero halo pad rows, row tables, first image's loads
    for (int i = tid; i < (int)(3 * rowB / 4); i += BK_THREADS) {
        reinterpret_cast<uint32_t*>(slab)[i] = 0u;
        reinterpret_cast<uint32_t*>(slab + (size_t)(L + 3) * rowB)[i] = 0u;
    }
    for (int t = tid; t < L; t += BK_THREADS) {
        const int64_t row = bk_row<POOL_T>(g, t);
        ztab[t] = (uint32_t)(row * a.ldxz);
        ytab[t] = (uint32_t)(row * a.ldy);
    }
    __syncthreads();
    pdl_wait();     // the pad rows and row tables above overlapped the in_proj GEMM's tail; x / z are read from here on
    pdl_trigger();
    int img = blockIdx.x;
    const int cpr = D >> 3;  // 16-byte chunks per token row
    if (img < g.B) {         // first image: every thread fetches its share of the slab
        const bf16* xb = a.x + (int64_t)img * a.xzbs;
        if (BK_THREADS % cpr == 0) {
            // a thread keeps its 16-byte column and walks the token rows with a fixed stride: no division in the loop
            // (the generic form below costs ~50 instructions per chunk, 3 % of the kernel's instructions at 224^2)
            const int rstep = BK_THREADS / cpr, c = tid % cpr;
            unsigned char* dst = slab + (uint32_t)(tid / cpr + 3) * rowB + c * 16;
            const bf16* src = xb + c * 8;
            for (int t = tid / cpr; t < L; t += rstep, dst += (uint32_t)rstep * rowB) cp_async16(dst, src + ztab[t], true);
        } else {
            for (int i = tid; i < L * cpr; i += BK_THREADS) {
                const int t = i / cpr, c = i - t * cpr;
                cp_async16(slab + (uint32_t)(t + 3) * rowB + c * 16, xb + ztab[t] + c * 8, true);
            }
        }
    }

    const int half_d = D >> 1, quart_d = D >> 2;
    // ---- pass-1 mapping: (channel pair, quarter of the pooled rows)
    const bool p1_live = tid < half_d * 4;
    const int p1_q = tid / half_d, p1_c = tid - p1_q * half_d;
    const int rpq = (outer + 3) >> 2;
    const int r_begin = min(outer, p1_q * rpq), r_end = min(outer, r_begin + rpq);
    const bool p1_work = p1_live && r_begin < r_end;
    // byte offset of this thread's channel pair inside a token row
    const uint32_t p1_off = (uint32_t)p1_c * 4u;

    for (uint32_t it_img = 0; img < g.B; img += gridDim.x, ++it_img) {
        const int img_next = img + gridDim.x;
        const bool has_next = img_next < g.B;
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();  // S0: this image's x rows (issued by all threads) have landed
        // Publish the previous image.  The barrier orders every warp's y stores before this thread (CTA scope); its
        // gpu-scope fence + release store is cumulative over them (PTX memory model: causality order is transitive), so the
        // consumer's ld.acquire of the flag observes the whole image.
        if (a.done_flags && it_img > 0 && tid == 0) {
            __threadfence();
            asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(a.done_flags + (img - (int)gridDim.x)), "r"(a.done_epoch) : "memory");
        }

        // ================= pass 1: conv (both directions) + SiLU + mean pool, w in place ===============
        uint32_t hl0 = 0, hl1 = 0, hl2 = 0;
        if (p1_work) {
            const uint32_t o = (uint32_t)(r_begin * P) * rowB + p1_off;  // token r_begin*P - 3
            hl0 = *reinterpret_cast<const uint32_t*>(smem + o);
            hl1 = *reinterpret_cast<const uint32_t*>(smem + o + rowB);
            hl2 = *reinterpret_cast<const uint32_t*>(smem + o + 2 * rowB);
        }
        for (int i = tid; i < 2 * outer * a.xld; i += BK_THREADS) xd[i] = 0.f;
        __syncthreads();  // S1: left halos are in registers, nobody has overwritten x yet
        uint32_t dfr0 = 0, dfr1 = 0, dfr2 = 0;  // w of the first 3 tokens of the segment: stored after S2
        if (p1_work) {
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
            uint32_t tok = (uint32_t)(r_begin * P + 3) * rowB + p1_off;  // byte offset of token r*P (this thread's pair)
            win[3] = unpack2(*reinterpret_cast<const uint32_t*>(smem + tok));
            win[4] = unpack2(*reinterpret_cast<const uint32_t*>(smem + tok + rowB));
            win[5] = unpack2(*reinterpret_cast<const uint32_t*>(smem + tok + 2 * rowB));
            uint32_t uo = ((uint32_t)r_begin * a.uld + d0) * 2u;  // byte offset in ubuf
            const uint32_t udir = (uint32_t)outer * a.uld * 2u;
            for (int r = r_begin; r < r_end; ++r, tok += (uint32_t)P * rowB, uo += (uint32_t)a.uld * 2u) {
                float2 sumf = make_float2(0.f, 0.f), sumb = sumf;
                const bool first_row = r == r_begin;
#define BK_TOKEN(C_, X_)                                                                                   \
    {                                                                                                      \
        X_(6) = unpack2(*reinterpret_cast<const uint32_t*>(smem + tok + (uint32_t)((C_) + 3) * rowB));    \
        float2 af = __ffma2_rn(wf[0], X_(0), bf_), ab = __ffma2_rn(wb[0], X_(6), bb_);                     \
        af = __ffma2_rn(wf[1], X_(1), af); ab = __ffma2_rn(wb[1], X_(5), ab);                              \
        af = __ffma2_rn(wf[2], X_(2), af); ab = __ffma2_rn(wb[2], X_(4), ab);                              \
        af = __ffma2_rn(wf[3], X_(3), af); ab = __ffma2_rn(wb[3], X_(3), ab);                              \
        af = silu2_from_half(af);                                                                          \
        ab = silu2_from_half(ab);                                                                          \
        sumf = __fadd2_rn(sumf, af);                                                                       \
        sumb = __fadd2_rn(sumb, ab);                                                                       \
        const uint32_t wv = pack2(__ffma2_rn(Db2, ab, __fmul2_rn(Df2, af)));                               \
        if (first_row && (C_) < 3) {                                                                       \
            if ((C_) == 0) dfr0 = wv; else if ((C_) == 1) dfr1 = wv; else dfr2 = wv;                        \
        } else {                                                                                           \
            *reinterpret_cast<uint32_t*>(smem + tok + (uint32_t)(C_) * rowB) = wv;                         \
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
                *reinterpret_cast<uint32_t*>(smem + a.off_u + uo) = pack2(__fmul2_rn(sumf, sc2));
                *reinterpret_cast<uint32_t*>(smem + a.off_u + uo + udir) = pack2(__fmul2_rn(sumb, sc2));
            }
        }
        // ---- x_proj work items (dir, row tile, column tile, k-half): 24 items at 224^2 -> one per warp.  The W_x
        // fragments of the warp's first item do not depend on the image: fetch them BEFORE the barrier so their
        // L2 latency overlaps the wait for the slower pass-1 threads.
        const int gq = lane >> 2, tq = lane & 3;
        const int n_mt = (outer + 15) >> 4;
        const int xp_nnt = (a.ncols + 7) >> 3, kper = D >> 1, xp_per_dm = 2 * xp_nnt, xp_total = 2 * n_mt * xp_per_dm;
        constexpr int KS_MAX = 64 * BK_NCG / 16;  // k-steps of one k-half at the widest supported dim
        uint32_t xb0[KS_MAX], xb1[KS_MAX];
        {
            const int dm = warp / xp_per_dm, item = warp - dm * xp_per_dm;
            const int dir = dm / n_mt, n = (item >> 1) * 8 + gq;
            if (a.xwp) {
                // fragment-order weights: one coalesced 16-byte load per lane covers two k-steps
                const int nj = kper >> 5;
                const uint4* pw = a.xwp + ((size_t)((dir * xp_nnt + (item >> 1)) * 2 + (item & 1)) * nj) * 32 + lane;
#pragma unroll
                for (int j = 0; j < KS_MAX / 2; ++j) {
                    uint4 q = make_uint4(0u, 0u, 0u, 0u);
                    if (warp < xp_total && j < nj) q = __ldg(pw + j * 32);
                    xb0[2 * j] = q.x; xb1[2 * j] = q.y; xb0[2 * j + 1] = q.z; xb1[2 * j + 1] = q.w;
                }
            } else {
                const bool nv = warp < xp_total && n < a.ncols;
                const bf16* wp = a.xw + ((size_t)dir * a.ncols + (nv ? n : 0)) * D + (item & 1) * kper + tq * 2;
#pragma unroll
                for (int ks = 0; ks < KS_MAX; ++ks) {
                    xb0[ks] = xb1[ks] = 0u;
                    if (nv && ks * 16 < kper) {
                        xb0[ks] = __ldg(reinterpret_cast<const unsigned int*>(wp + ks * 16));
                        xb1[ks] = __ldg(reinterpret_cast<const unsigned int*>(wp + ks * 16 + 8));
                    }
                }
            }
        }
        __syncthreads();  // S2: u complete; every right halo has been read
        if (p1_work) {
            const uint32_t tok = (uint32_t)(r_begin * P + 3) * rowB + p1_off;
            *reinterpret_cast<uint32_t*>(smem + tok) = dfr0;
            *reinterpret_cast<uint32_t*>(smem + tok + rowB) = dfr1;
            *reinterpret_cast<uint32_t*>(smem + tok + 2 * rowB) = dfr2;
        }

        // ================= x_proj on tensor cores: xd[dir][j][c] = sum_d u[dir][j][d] W_x[dir][c][d] ===============
        for (int idx = warp; idx < xp_total; idx += BK_WARPS) {
            const int dm = idx / xp_per_dm, item = idx - dm * xp_per_dm;
            const int dir = dm / n_mt, mt = dm - dir * n_mt;
            const int row0 = mt * 16 + gq, row1 = row0 + 8;
            const int kh = item & 1, nt = item >> 1;
            const bf16* pa0 = ubuf + ((uint32_t)dir * outer + min(row0, outer - 1)) * a.uld + tq * 2 + kh * kper;
            const bf16* pa1 = ubuf + ((uint32_t)dir * outer + min(row1, outer - 1)) * a.uld + tq * 2 + kh * kper;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            if (idx == warp) {  // first item: W_x fragments already in registers
#pragma unroll
                for (int ks = 0; ks < KS_MAX; ++ks) {
                    if (ks * 16 < kper) {
                        const uint32_t a0 = *reinterpret_cast<const uint32_t*>(pa0 + ks * 16);
                        const uint32_t a1 = *reinterpret_cast<const uint32_t*>(pa1 + ks * 16);
                        const uint32_t a2 = *reinterpret_cast<const uint32_t*>(pa0 + ks * 16 + 8);
                        const uint32_t a3 = *reinterpret_cast<const uint32_t*>(pa1 + ks * 16 + 8);
                        mma16816(acc, a0, a1, a2, a3, xb0[ks], xb1[ks]);
                    }
                }
            } else {
                const int n = nt * 8 + gq;
                const bool nv = n < a.ncols;
                const bf16* wp = a.xw + ((size_t)dir * a.ncols + (nv ? n : 0)) * D + kh * kper + tq * 2;
#pragma unroll 4
                for (int k0 = 0; k0 < kper; k0 += 16) {
                    const uint32_t a0 = *reinterpret_cast<const uint32_t*>(pa0 + k0);
                    const uint32_t a1 = *reinterpret_cast<const uint32_t*>(pa1 + k0);
                    const uint32_t a2 = *reinterpret_cast<const uint32_t*>(pa0 + k0 + 8);
                    const uint32_t a3 = *reinterpret_cast<const uint32_t*>(pa1 + k0 + 8);
                    uint32_t b0 = 0, b1 = 0;
                    if (nv) {
                        b0 = __ldg(reinterpret_cast<const unsigned int*>(wp + k0));
                        b1 = __ldg(reinterpret_cast<const unsigned int*>(wp + k0 + 8));
                    }
                    mma16816(acc, a0, a1, a2, a3, b0, b1);
                }
            }
            const int c0 = nt * 8 + tq * 2;
            float* o0 = xd + ((uint32_t)dir * outer + row0) * a.xld + c0;
            float* o1 = xd + ((uint32_t)dir * outer + row1) * a.xld + c0;
            if (row0 < outer) {
                if (c0 < a.ncols) atomicAdd(o0, acc[0]);
                if (c0 + 1 < a.ncols) atomicAdd(o0 + 1, acc[1]);
            }
            if (row1 < outer) {
                if (c0 < a.ncols) atomicAdd(o1, acc[2]);
                if (c0 + 1 < a.ncols) atomicAdd(o1 + 1, acc[3]);
            }
        }
        __syncthreads();  // S3: xd complete

        {
            if (a.xdbl_out) {  // saved for backward: (2, B*Lp, ncols) bf16
                for (int i = tid; i < 2 * outer * a.ncols; i += BK_THREADS) {
                    const int c = i % a.ncols, rj = i / a.ncols;  // rj = dir*outer + j
                    const int dir = rj / outer, j = rj - dir * outer;
                    a.xdbl_out[(((int64_t)dir * g.B + img) * outer + j) * a.ncols + c] =
                        __float2bfloat16_rn(xd[(size_t)rj * a.xld + c]);
                }
            }
        }
        // z of the first gate round: issued now, lands while the scan runs
        const bf16* zb = a.z + (int64_t)img * a.xzbs;
        uint2 zc[BK_NCG];
#pragma unroll
        for (int i = 0; i < BK_NCG; ++i) {
            zc[i] = make_uint2(0u, 0u);
            const int cg = lane + 32 * i;
            if (warp < BK_NWORK && warp < L && cg < quart_d)
                zc[i] = __ldg(reinterpret_cast<const uint2*>(zb + ztab[warp] + cg * 4));
        }

        // ================= bidirectional selective scan over the pooled rows =======================
        if (tid < 2 * D) {
            const int dir = tid >= D ? 1 : 0, d = tid - dir * D;
            constexpr float LOG2E = 1.4426950408889634f;
            float2 A2[N / 2], h[N / 2];
            {
                const float4* Ap = reinterpret_cast<const float4*>(a.A + ((int64_t)dir * D + d) * N);
#pragma unroll
                for (int n4 = 0; n4 < N / 4; ++n4) {
                    float4 v = __ldg(Ap + n4);
                    if (a.a_is_log) {
                        v.x = -__expf(v.x); v.y = -__expf(v.y); v.z = -__expf(v.z); v.w = -__expf(v.w);
                    }
                    A2[2 * n4] = make_float2(v.x * LOG2E, v.y * LOG2E);
                    A2[2 * n4 + 1] = make_float2(v.z * LOG2E, v.w * LOG2E);
                }
#pragma unroll
                for (int n = 0; n < N / 2; ++n) h[n] = make_float2(0.f, 0.f);
            }
            // dt_proj row of this channel in registers (R <= 16): delta_pre = bias + W_dt[d,:] . dt[j,:]
            float Wd[RT];
            {
                const float* Wp = a.dtw + ((int64_t)dir * D + d) * RT;
#pragma unroll
                for (int r = 0; r < RT; ++r) Wd[r] = __ldg(Wp + r);
            }
            const float bias = a.dtb[(int64_t)dir * D + d];
            const int64_t gplane = ((int64_t)dir * g.B + img) * outer;
            const uint32_t dbase = (uint32_t)dir * outer;
            for (int step = 0; step < outer; ++step) {
                const uint32_t j = dir ? outer - 1 - step : step;
                float* sp = sbuf + (dbase + j) * D + d;
                const bf16 ub = ubuf[(dbase + j) * a.uld + d];
                const float* xrow = xd + (dbase + j) * a.xld;
                float dpre = bias;
#pragma unroll
                for (int r4 = 0; r4 < RT / 4; ++r4) {
                    const float4 q = *reinterpret_cast<const float4*>(xrow + 4 * r4);
                    dpre = fmaf(Wd[4 * r4], q.x, dpre); dpre = fmaf(Wd[4 * r4 + 1], q.y, dpre);
                    dpre = fmaf(Wd[4 * r4 + 2], q.z, dpre); dpre = fmaf(Wd[4 * r4 + 3], q.w, dpre);
                }
                const float delta = bk_softplus(dpre);
                const float du = delta * __bfloat162float(ub);
                const float* row = xrow + RT;
                float2 y2 = make_float2(0.f, 0.f);
                const float2 dl2 = make_float2(delta, delta), du2 = make_float2(du, du);
#pragma unroll
                for (int n = 0; n < N / 2; n += 2) {
                    const float4 Bq = *reinterpret_cast<const float4*>(row + 2 * n);
                    const float4 Cq = *reinterpret_cast<const float4*>(row + N + 2 * n);
                    float2 e0 = __fmul2_rn(dl2, A2[n]), e1 = __fmul2_rn(dl2, A2[n + 1]);
                    e0 = make_float2(bk_ex2(e0.x), bk_ex2(e0.y));
                    e1 = make_float2(bk_ex2(e1.x), bk_ex2(e1.y));
                    h[n] = __ffma2_rn(e0, h[n], __fmul2_rn(du2, make_float2(Bq.x, Bq.y)));
                    h[n + 1] = __ffma2_rn(e1, h[n + 1], __fmul2_rn(du2, make_float2(Bq.z, Bq.w)));
                    y2 = __ffma2_rn(h[n], make_float2(Cq.x, Cq.y), y2);
                    y2 = __ffma2_rn(h[n + 1], make_float2(Cq.z, Cq.w), y2);
                }
                const float yv = y2.x + y2.y;
                *sp = yv;
                if (a.s_out) a.s_out[(gplane + j) * D + d] = yv;
                if (a.u_out) a.u_out[(gplane + j) * D + d] = ub;
            }
        }
        __syncthreads();  // S5: s complete

        // ================= gate pass: v = w + (s_f + s_b)/2, LayerNorm, * silu(z), store y ===========
        if (warp < BK_NWORK) {
            const float invD = 1.f / (float)D;
            // per-lane constants: channel groups lane, lane+32, lane+64 (4 channels each)
            float2 gam[BK_NCG][2], bet[BK_NCG][2];
#pragma unroll
            for (int i = 0; i < BK_NCG; ++i) {
                gam[i][0] = gam[i][1] = make_float2(1.f, 1.f);
                bet[i][0] = bet[i][1] = make_float2(0.f, 0.f);
                const int cg = lane + 32 * i;
                if (NORM && (FULL || cg < quart_d)) {
                    const float4 gq4 = __ldg(reinterpret_cast<const float4*>(a.lnw + cg * 4));
                    gam[i][0] = make_float2(gq4.x, gq4.y); gam[i][1] = make_float2(gq4.z, gq4.w);
                    if (a.lnb) {
                        const float4 bq4 = __ldg(reinterpret_cast<const float4*>(a.lnb + cg * 4));
                        bet[i][0] = make_float2(bq4.x, bq4.y); bet[i][1] = make_float2(bq4.z, bq4.w);
                    }
                }
            }
            const bf16* zlane = zb + lane * 4;
            bf16* ylane = a.y + (int64_t)img * a.ybs + lane * 4;
            const uint32_t lane_w = (uint32_t)lane * 8u;  // byte offset of the lane's first 4-channel group in a slab row
            const bf16* xnext = a.x + (int64_t)img_next * a.xzbs;
            const float* s_lane = sbuf + lane * 4;
            const uint32_t splane = (uint32_t)outer * D;
            const float2 half2c = make_float2(0.5f, 0.5f);
            for (int t = warp; t < L; t += BK_NWORK) {
                // z of the next round (one round ahead, through registers)
                uint2 zn[BK_NCG];
                const int tn = t + BK_NWORK;
                {
                    const bf16* zrow = zlane + ztab[tn < L ? tn : t];
#pragma unroll
                    for (int i = 0; i < BK_NCG; ++i) {
                        zn[i] = make_uint2(0u, 0u);
                        if (tn < L && (FULL || lane + 32 * i < quart_d))
                            zn[i] = __ldg(reinterpret_cast<const uint2*>(zrow + 128 * i));
                    }
                }
                const uint32_t j = (uint32_t)(t / P);
                unsigned char* wrow = smem + (uint32_t)(t + 3) * rowB;
                const float* srow = s_lane + j * D;
                float2 v[BK_NCG][2];
                float2 sum2 = make_float2(0.f, 0.f), sq2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < BK_NCG; ++i) {
                    v[i][0] = v[i][1] = make_float2(0.f, 0.f);
                    if (FULL || lane + 32 * i < quart_d) {
                        const uint2 wv = *reinterpret_cast<const uint2*>(wrow + lane_w + 256 * i);
                        const float4 sf = *reinterpret_cast<const float4*>(srow + 128 * i);
                        const float4 sb = *reinterpret_cast<const float4*>(srow + splane + 128 * i);
                        const float2 s01 = __fadd2_rn(make_float2(sf.x, sf.y), make_float2(sb.x, sb.y));
                        const float2 s23 = __fadd2_rn(make_float2(sf.z, sf.w), make_float2(sb.z, sb.w));
                        v[i][0] = __ffma2_rn(half2c, s01, unpack2(wv.x));
                        v[i][1] = __ffma2_rn(half2c, s23, unpack2(wv.y));
                        if (NORM) {
                            sum2 = __fadd2_rn(sum2, __fadd2_rn(v[i][0], v[i][1]));
                            sq2 = __ffma2_rn(v[i][0], v[i][0], sq2);
                            sq2 = __ffma2_rn(v[i][1], v[i][1], sq2);
                        }
                    }
                }
                // this token's w row is in registers (all lanes): refill it with the same token of the next image
                __syncwarp();
                if (has_next) {
                    const bf16* xsrc = xnext + ztab[t];
                    for (int c = lane; c < cpr; c += 32) cp_async16(wrow + c * 16, xsrc + c * 8, true);
                }
                float2 gsc = make_float2(1.f, 1.f), nmean = make_float2(0.f, 0.f);
                if (NORM) {
                    float sum = sum2.x + sum2.y, sq = sq2.x + sq2.y;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        sum += __shfl_xor_sync(0xffffffffu, sum, o);
                        sq += __shfl_xor_sync(0xffffffffu, sq, o);
                    }
                    const float mean = sum * invD;
                    const float rstd = rsqrtf(fmaxf(fmaf(sq, invD, -mean * mean), 0.f) + a.eps);
                    gsc = make_float2(rstd, rstd);
                    nmean = make_float2(-mean, -mean);
                }
                bf16* yrow = ylane + ytab[t];
#pragma unroll
                for (int i = 0; i < BK_NCG; ++i) {
                    if (FULL || lane + 32 * i < quart_d) {
                        float2 o0 = v[i][0], o1 = v[i][1];
                        if (NORM) {
                            o0 = __ffma2_rn(__fmul2_rn(__fadd2_rn(o0, nmean), gsc), gam[i][0], bet[i][0]);
                            o1 = __ffma2_rn(__fmul2_rn(__fadd2_rn(o1, nmean), gsc), gam[i][1], bet[i][1]);
                        }
                        float2 h0 = __fmul2_rn(unpack2(zc[i].x), half2c), h1 = __fmul2_rn(unpack2(zc[i].y), half2c);
                        h0 = silu2_from_half(h0);
                        h1 = silu2_from_half(h1);
                        uint2 pk;
                        pk.x = pack2(__fmul2_rn(o0, h0));
                        pk.y = pack2(__fmul2_rn(o1, h1));
                        *reinterpret_cast<uint2*>(yrow + 128 * i) = pk;
                    }
                }
#pragma unroll
                for (int i = 0; i < BK_NCG; ++i) zc[i] = zn[i];
            }
        }
    }
    if (a.done_flags) {  // last image of this CTA
        __syncthreads();
        const int last = img - (int)gridDim.x;
        if (tid == 0 && last >= 0 && last < g.B) {
            __threadfence();
            asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(a.done_flags + last), "r"(a.done_epoch) : "memory");
        }
    }
}

// x_proj weights (2, ncols, D) bf16 -> MMA B-fragment order: for work item (dir, column tile nt, k-half kh), uint4 j of
// lane l holds {b0, b1} of k-steps 2j and 2j+1 (b0 = W[n = nt*8 + l/4][k .. k+1], b1 = the same at k + 8, with
// k = kh*D/2 + ks*16 + (l%4)*2), so a warp reads its 12 k-steps with six fully coalesced 16-byte loads per lane
// instead of 24 scattered 4-byte ones.  Columns >= ncols are zero.
__global__ void __launch_bounds__(256)
pack_xproj_kernel(const bf16* __restrict__ w, int ncols, int D, int n_nt, int nj, uint4* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 2 * n_nt * 2 * nj * 32) return;
    const int lane = idx & 31, j = (idx >> 5) % nj, item = (idx >> 5) / nj;
    const int kh = item & 1, nt = (item >> 1) % n_nt, dir = (item >> 1) / n_nt;
    const int n = nt * 8 + (lane >> 2), kper = D >> 1;
    uint32_t r[4] = {0u, 0u, 0u, 0u};
    if (n < ncols) {
        const bf16* row = w + ((size_t)dir * ncols + n) * D + kh * kper + (lane & 3) * 2;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = (2 * j + h) * 16;
            r[2 * h] = *reinterpret_cast<const uint32_t*>(row + k);
            r[2 * h + 1] = *reinterpret_cast<const uint32_t*>(row + k + 8);
        }
    }
    out[idx] = make_uint4(r[0], r[1], r[2], r[3]);
}

struct BlockPlan {
    int ok;
    size_t smem;
    int xld, uld, off_u, off_s, off_xdbl, off_tab;
};

static BlockPlan plan_block(const fv_geom* g, int dtype, int R, int N, int64_t ldxz, int64_t ldy) {
    BlockPlan p;
    p.ok = 0;
    if (dtype != FV_BF16 || g->inner != 1 || N != BK_NSTATE) return p;
    const int D = g->dim, outer = g->outer, pool = g->pool;
    if (D % 32 != 0 || D > 128 * BK_NCG || pool < 4 || R <= 0 || R % 4 != 0 || R > 16) return p;
    const int64_t L = (int64_t)outer * pool;
    if (L * (ldxz > ldy ? ldxz : ldy) >= (1ll << 31)) return p;  // 32-bit row-offset tables
    const int ncols = R + 2 * N;
    p.xld = (ncols + 7) / 8 * 8;
    p.uld = D + 8;
    auto up = [](size_t v) { return (v + 127) / 128 * 128; };
    size_t off = up((size_t)(L + 6) * D * 2);
    p.off_u = (int)off;
    off = up(off + (size_t)2 * outer * p.uld * 2);
    p.off_s = (int)off;
    off = up(off + (size_t)2 * outer * D * 4);
    p.off_xdbl = (int)off;
    off = up(off + (size_t)2 * outer * p.xld * 4);
    p.off_tab = (int)off;
    off = up(off + (size_t)2 * L * 4);
    p.smem = off;
    if (off > 227 * 1024) return p;  // shared-memory capacity of one SM
    p.ok = 1;
    return p;
}

}  // namespace fv

// cluster form (block_cluster.cu): channels of one image split over the CTAs of a thread-block cluster
namespace fv {
struct ClusterPlan {
    int ok;
    size_t smem;
    int xld, uld, nnt, C, xdb, off_u, off_s, off_xp, off_xd, off_st;
};
ClusterPlan plan_cluster(const fv_geom* g, int dtype, int R, int N, int64_t ldxz, int64_t ldy);
int64_t pack_xproj_slab_bytes(int dim, int ncols);
int pack_xproj_slab(int dim, int ncols, const void* xproj_w, void* packed, cudaStream_t st);
int launch_block_cluster(const fv_geom* g_, const ClusterPlan& p, const void* x, const void* z, int64_t ldxz, int64_t xz_bstride,
                         const float* conv_w, const float* conv_b, const void* xw_slab_packed, const float* dt_w,
                         const float* dt_bias, const float* A, int a_is_log, int dt_rank, int dstate, const float* Dskip,
                         const float* ln_w, const float* ln_b, float eps, float scale, void* y, int64_t ldy, int64_t y_bstride,
                         void* u_out, void* xdbl_out, float* s_out, void* v_out, float* pre_out, cudaStream_t stream);
// Which kernel serves a configuration both can run (dim <= 384, i.e. FastVim-T): measured on B200 at batch 256 the
// one-CTA-per-image kernel is still ahead there (68 vs 76 us), so "auto" gives it the narrow models and the cluster
// kernel everything wider.  FASTVIM_BLOCK_CLUSTER=1 forces the cluster kernel wherever it applies, =0 disables it
// (A/B timing, tools/kbench.py).
static int cluster_mode() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("FASTVIM_BLOCK_CLUSTER");
        v = !e ? 2 : (e[0] == '0' ? 0 : (e[0] == '1' ? 1 : 2));
    }
    return v;
}
static bool cluster_enabled() { return cluster_mode() != 0; }
static bool prefer_cluster(const fv_geom* g, int dtype, int R, int N, int64_t ldxz, int64_t ldy) {
    if (cluster_mode() == 1) return true;
    return !plan_block(g, dtype, R, N, ldxz, ldy).ok;
}
static int64_t pack_old_bytes(int dim, int ncols) {
    if (dim <= 0 || ncols <= 0 || dim % 64 != 0) return 0;
    return (int64_t)2 * ((ncols + 7) / 8) * 2 * (dim / 64) * 32 * 16;
}
}  // namespace fv

extern "C" int fv_block_fwd_supported(const fv_geom* g, int dtype, int dt_rank, int dstate) {
    if (!g || g->batch <= 0 || g->dim <= 0 || g->outer <= 0 || g->pool <= 0) return 0;
    if (fv::cluster_enabled() && fv::plan_cluster(g, dtype, dt_rank, dstate, 2 * (int64_t)g->dim, g->dim).ok) return 1;
    return fv::plan_block(g, dtype, dt_rank, dstate, 2 * (int64_t)g->dim, g->dim).ok;
}

extern "C" int fv_block_fwd_saves_v(const fv_geom* g, int dtype, int dt_rank, int dstate) {
    if (!g || g->batch <= 0 || g->dim <= 0 || g->outer <= 0 || g->pool <= 0) return 0;
    return fv::cluster_enabled() && fv::plan_cluster(g, dtype, dt_rank, dstate, 2 * (int64_t)g->dim, g->dim).ok;
}

// packed x_proj weights = [one-CTA-per-image fragment order | per-slab fragment order of the cluster kernel]
extern "C" int64_t fv_block_pack_xproj_bytes(int dim, int ncols) {
    if (dim <= 0 || ncols <= 0) return 0;
    return fv::pack_old_bytes(dim, ncols) + fv::pack_xproj_slab_bytes(dim, ncols);
}

extern "C" int fv_block_pack_xproj(int dim, int ncols, const void* xproj_w, void* packed, void* stream) {
    using namespace fv;
    FV_REQUIRE(xproj_w && packed, "fv_block_pack_xproj: null pointer");
    FV_REQUIRE(dim > 0 && ncols > 0 && fv_block_pack_xproj_bytes(dim, ncols) > 0,
               "fv_block_pack_xproj: dim (%d) must be a positive multiple of 64 or 192", dim);
    FV_REQUIRE(((uintptr_t)xproj_w % 4) == 0 && ((uintptr_t)packed % 16) == 0, "fv_block_pack_xproj: misaligned pointer");
    const int64_t old_b = pack_old_bytes(dim, ncols);
    if (old_b) {
        const int n_nt = (ncols + 7) / 8, nj = dim / 64, total = 2 * n_nt * 2 * nj * 32;
        pack_xproj_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const bf16*)xproj_w, ncols, dim, n_nt, nj,
                                                                             (uint4*)packed);
        if (int rc = finish_launch("block_pack_xproj")) return rc;
    }
    if (pack_xproj_slab_bytes(dim, ncols))
        return pack_xproj_slab(dim, ncols, xproj_w, (unsigned char*)packed + old_b, (cudaStream_t)stream);
    return 0;
}

static int block_fwd_impl(const fv_geom* g_, int dtype, const void* x, const void* z, int64_t ldxz,
                            int64_t xz_bstride, const float* conv_w, const float* conv_b, const void* xproj_w,
                            const void* xproj_w_packed, const float* dt_w, const float* dt_bias, const float* A, int a_is_log,
                            int dt_rank, int dstate, const float* Dskip, const float* ln_w, const float* ln_b,
                            float eps, float scale, void* y, int64_t ldy, int64_t y_bstride,
                            void* u_out, void* xdbl_out, float* s_out, void* v_out, float* pre_out, int* done_flags,
                            int done_epoch, void* stream);

extern "C" int fv_block_fwd(const fv_geom* g_, int dtype, const void* x, const void* z, int64_t ldxz,
                            int64_t xz_bstride, const float* conv_w, const float* conv_b, const void* xproj_w,
                            const void* xproj_w_packed, const float* dt_w, const float* dt_bias, const float* A, int a_is_log,
                            int dt_rank, int dstate, const float* Dskip, const float* ln_w, const float* ln_b,
                            float eps, float scale, void* y, int64_t ldy, int64_t y_bstride,
                            void* u_out, void* xdbl_out, float* s_out, void* v_out, float* pre_out, void* stream) {
    return block_fwd_impl(g_, dtype, x, z, ldxz, xz_bstride, conv_w, conv_b, xproj_w, xproj_w_packed, dt_w, dt_bias, A, a_is_log,
                          dt_rank, dstate, Dskip, ln_w, ln_b, eps, scale, y, ldy, y_bstride, u_out, xdbl_out, s_out, v_out,
                          pre_out, nullptr, 0, stream);
}

// Inference form that also publishes per-image completion: done_flags[img] = done_epoch (release) once the image's y rows
// are written, so that fv_gemm_out_norm_flow -- launched programmatically dependent, resident on the SMs this kernel has
// already left -- starts on finished images while the rest are still being computed.  One-CTA-per-image kernel only.
extern "C" int fv_block_fwd_signal_supported(const fv_geom* g, int dtype, int dt_rank, int dstate) {
    if (!g || g->batch <= 0 || g->dim <= 0 || g->outer <= 0 || g->pool <= 0) return 0;
    const int64_t ld = 2 * (int64_t)g->dim;
    if (fv::cluster_enabled() && fv::prefer_cluster(g, dtype, dt_rank, dstate, ld, g->dim)) return 0;
    return fv::plan_block(g, dtype, dt_rank, dstate, ld, g->dim).ok;
}

extern "C" int fv_block_fwd_signal(const fv_geom* g_, int dtype, const void* x, const void* z, int64_t ldxz,
                                   int64_t xz_bstride, const float* conv_w, const float* conv_b, const void* xproj_w,
                                   const void* xproj_w_packed, const float* dt_w, const float* dt_bias, const float* A,
                                   int a_is_log, int dt_rank, int dstate, const float* Dskip, const float* ln_w,
                                   const float* ln_b, float eps, float scale, void* y, int64_t ldy, int64_t y_bstride,
                                   int* done_flags, int done_epoch, void* stream) {
    FV_REQUIRE(done_flags && done_epoch > 0, "fv_block_fwd_signal: done_flags must be given and done_epoch positive");
    FV_REQUIRE(g_ && fv_block_fwd_signal_supported(g_, dtype, dt_rank, dstate),
               "fv_block_fwd_signal: configuration is not served by the one-CTA-per-image kernel");
    return block_fwd_impl(g_, dtype, x, z, ldxz, xz_bstride, conv_w, conv_b, xproj_w, xproj_w_packed, dt_w, dt_bias, A, a_is_log,
                          dt_rank, dstate, Dskip, ln_w, ln_b, eps, scale, y, ldy, y_bstride, nullptr, nullptr, nullptr, nullptr,
                          nullptr, done_flags, done_epoch, stream);
}

static int block_fwd_impl(const fv_geom* g_, int dtype, const void* x, const void* z, int64_t ldxz,
                            int64_t xz_bstride, const float* conv_w, const float* conv_b, const void* xproj_w,
                            const void* xproj_w_packed, const float* dt_w, const float* dt_bias, const float* A, int a_is_log,
                            int dt_rank, int dstate, const float* Dskip, const float* ln_w, const float* ln_b,
                            float eps, float scale, void* y, int64_t ldy, int64_t y_bstride,
                            void* u_out, void* xdbl_out, float* s_out, void* v_out, float* pre_out, int* done_flags,
                            int done_epoch, void* stream) {
    using namespace fv;
    if (int rc = check_geom(g_, "fv_block_fwd")) return rc;
    FV_REQUIRE(x && z && conv_w && xproj_w && dt_w && dt_bias && A && Dskip && y, "fv_block_fwd: null pointer");
    if (cluster_enabled() && xproj_w_packed && !done_flags) {
        const ClusterPlan cp = plan_cluster(g_, dtype, dt_rank, dstate, ldxz, ldy);
        // the pre-norm value v (saved for the streaming gate backward) only exists in the cluster kernel
        if (cp.ok && (v_out || prefer_cluster(g_, dtype, dt_rank, dstate, ldxz, ldy)))
            return launch_block_cluster(g_, cp, x, z, ldxz, xz_bstride, conv_w, conv_b,
                                        (const unsigned char*)xproj_w_packed + pack_old_bytes(g_->dim, dt_rank + 2 * dstate), dt_w,
                                        dt_bias, A, a_is_log, dt_rank, dstate, Dskip, ln_w, ln_b, eps, scale, y, ldy, y_bstride,
                                        u_out, xdbl_out, s_out, v_out, pre_out, (cudaStream_t)stream);
    }
    FV_REQUIRE(!v_out && !pre_out, "fv_block_fwd: v_out / pre_out need the cluster kernel (fv_block_fwd_saves_v() == 1 and packed x_proj weights)");
    const BlockPlan p = plan_block(g_, dtype, dt_rank, dstate, ldxz, ldy);
    FV_REQUIRE(p.ok, "fv_block_fwd: unsupported configuration (bf16, plain geometry, d_state 16, and either dim %% 192 == 0 with "
                     "<= 16 pooled rows and packed x_proj weights [cluster kernel], or dim %% 32 == 0, dim <= 384, dt_rank in "
                     "{4, 8, 12, 16} and the slab fitting 227 KB); use the four-launch path");
    FV_REQUIRE(ldxz % 8 == 0 && xz_bstride % 8 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)z % 8) == 0,
               "fv_block_fwd: x rows must be 16-byte aligned (ldxz %lld)", (long long)ldxz);
    FV_REQUIRE(ldy % 4 == 0 && y_bstride % 4 == 0 && ((uintptr_t)y % 8) == 0, "fv_block_fwd: y rows must be 8-byte aligned");
    FV_REQUIRE(((uintptr_t)xproj_w % 4) == 0, "fv_block_fwd: x_proj weights must be 4-byte aligned");
    BlockArgs a;
    a.g = make_geom(g_);
    a.x = (const bf16*)x; a.z = (const bf16*)z; a.ldxz = ldxz; a.xzbs = xz_bstride;
    a.cw = conv_w; a.cb = conv_b; a.xw = (const bf16*)xproj_w; a.xwp = (g_->dim % 64 == 0) ? (const uint4*)xproj_w_packed : nullptr; a.dtw = dt_w; a.dtb = dt_bias;
    a.A = A; a.a_is_log = a_is_log; a.Dskip = Dskip; a.lnw = ln_w; a.lnb = ln_b; a.eps = eps;
    a.scale = scale / (float)g_->pool;
    a.y = (bf16*)y; a.ldy = ldy; a.ybs = y_bstride;
    a.u_out = (bf16*)u_out; a.xdbl_out = (bf16*)xdbl_out; a.s_out = s_out;
    a.R = dt_rank; a.ncols = dt_rank + 2 * dstate; a.xld = p.xld; a.uld = p.uld;
    a.off_u = p.off_u; a.off_s = p.off_s; a.off_xdbl = p.off_xdbl; a.off_tab = p.off_tab;
    a.done_flags = done_flags; a.done_epoch = done_epoch;

    void (*kern)(const BlockArgs) = nullptr;
    const bool full = g_->dim == 128 * BK_NCG;  // every lane owns exactly BK_NCG channel groups: no predicates
#define FV_BK_PICK3(P_, N_, F_)                                                    \
    (dt_rank == 4 ? block_fwd_kernel<P_, N_, F_, 4> : dt_rank == 8 ? block_fwd_kernel<P_, N_, F_, 8> \
     : dt_rank == 12 ? block_fwd_kernel<P_, N_, F_, 12> : block_fwd_kernel<P_, N_, F_, 16>)
#define FV_BK_PICK2(P_, N_) (full ? FV_BK_PICK3(P_, N_, true) : FV_BK_PICK3(P_, N_, false))
#define FV_BK_PICK(P_) (ln_w ? FV_BK_PICK2(P_, true) : FV_BK_PICK2(P_, false))
    if (g_->pool == 14) kern = FV_BK_PICK(14);
    else kern = FV_BK_PICK(0);
#undef FV_BK_PICK3
#undef FV_BK_PICK2
#undef FV_BK_PICK
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    FV_REQUIRE(e == cudaSuccess, "fv_block_fwd: cudaFuncSetAttribute(%zu): %s", p.smem, cudaGetErrorString(e));
    const int grid = g_->batch < sm_count() ? g_->batch : sm_count();
    e = launch_pdl(kern, dim3(grid), dim3(BK_THREADS), p.smem, (cudaStream_t)stream, (const BlockArgs)a);
    FV_REQUIRE(e == cudaSuccess, "fv_block_fwd: launch: %s", cudaGetErrorString(e));
    return finish_launch("block_fwd");
}
