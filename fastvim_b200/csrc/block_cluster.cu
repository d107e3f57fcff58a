// K-fused, cluster form -- the whole SSM-block interior of one image by a THREAD-BLOCK CLUSTER whose CTAs split the
// d_inner channels; the image's x stays resident in the (distributed) shared memory of the cluster.
//
// What it replaces in the reference (paths relative to /root/reference): everything between the in_proj output and
// the out_proj input of mamba_ssm/modules/mamba_simple_faster.py:269-453 (flip, 2x causal_conv1d_fn, 2x mean pool,
// 2x x_proj, 2x dt_proj, 2x selective_scan_fn, repeat_interleave + D skip, merge, LayerNorm, * silu(z)); scan kernel
// csrc/selective_scan/selective_scan_fwd_kernel.cuh:67-303.  Round 1's block_fwd.cu did this with ONE 768-thread CTA
// per SM owning a whole image (dim <= 384): ncu showed one resident CTA exposing every barrier and fixed-latency
// phase to the whole SM (issue slots 58 %, barrier stalls 50-60 % around x_proj), and FastVim-S/B did not fit.
//
// B200 mapping here.  Conv, pool, scan recurrence, D skip and the z gate are per-channel; the only couplings across
// channels are x_proj (a contraction over d_inner) and the LayerNorm over d_inner.  So an image is given to a cluster of
// C = d_inner / 192 CTAs (FastVim-T 2, -S 4, -B 8), CTA r owning channels [192 r, 192 r + 192):
//   * 384 threads and ~110 KB of shared memory per CTA  ->  TWO CTAs per SM, working on different images / phases, so
//     one CTA's barrier, DSMEM round trip or MUFU-bound scan overlaps the other's streaming phases;
//   * x_proj: each CTA multiplies its 192-channel slab of the pooled u on the tensor cores (mma.sync m16n8k16, one
//     (direction, 8-column tile) item per warp, full K-slab, no atomics) into a partial (2 x Lp x (R+2N)) fp32 tile;
//     after ONE cluster barrier every CTA sums the C partials through distributed shared memory in rank order
//     (deterministic, identical on every CTA);
//   * LayerNorm: per-token (sum, sum of squares) of each slab are exchanged the same way (second cluster barrier);
//   * grid = (C, images): the hardware scheduler hands clusters to SMs as slots free up (no persistent loop, no tail
//     imbalance beyond one half-image), and FastVim-S/B run the same kernel with 4 / 8 CTAs per image.
// Gate pass: 16 lanes per token (12 channels per lane, so the lane's LayerNorm weights live in 24 registers), two tokens
// per warp instruction, LayerNorm reductions are 4-step butterflies.  Scan: one thread per (channel, direction); the dt_proj
// pre-activations of all pooled rows are formed first (dt_proj rows pass through registers 12 at a time, so dt_rank 48 of
// FastVim-B needs no more registers than dt_rank 12), then the recurrence runs from registers with B / C broadcast
// from shared memory; both directions meet in ONE fp32 plane (forward stores, backward adds after a barrier -- shared fp32
// atomics are CAS loops in SASS).  HBM traffic stays the algorithmic 3 * B * L * D * 2 bytes.

#include <cooperative_groups.h>

#include "block_common.cuh"

namespace fv {

int sm_count();
int check_geom(const fv_geom* g, const char* who);

constexpr int BC_DC = 192;             // channels per CTA
constexpr int BC_THREADS = 2 * BC_DC;  // one scan thread per (channel, direction)
constexpr int BC_WARPS = BC_THREADS / 32;
constexpr int BC_MAXO = 16;            // pooled rows held in registers by the scan
constexpr int BC_NG = BC_DC / 64;      // 4-channel chunks per lane in the gate pass (16 lanes per token)

struct ClusterArgs {
    Geom g;
    const bf16* x;
    const bf16* z;
    int64_t ldxz, xzbs;
    const float* cw;
    const float* cb;
    const uint4* xwp;  // x_proj weights in slab fragment order (pack_xproj_slab_kernel)
    const float* dtw;
    const float* dtb;
    const float* A;
    int a_is_log;
    const float* Dskip;
    const float* lnw;
    const float* lnb;
    float eps, scale;
    bf16* y;
    int64_t ldy, ybs;
    bf16* u_out;
    bf16* xdbl_out;
    float* s_out;
    bf16* v_out;       // (B, L, dim) pre-norm merged value, memory token order (saved for fv_gate_bwd_v), or null
    float* pre_out;    // (2, B, Lp, dim) dt_proj pre-activation dt_bias + W_dt . dt (saved for fv_scan_bwd_short), or null
    int R, ncols, xld, uld, nnt, C;
    int off_u, off_s, off_xp, off_xd, off_st;
};

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t dsmem_addr(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ float4 ld_dsmem4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float2 ld_dsmem2(uint32_t a) {
    float2 v;
    asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// xd element access: fp32, or bf16 when the (2, Lp, R+2N) tile would not fit beside the slab (dt_rank > 16) -- the
// reference's x_dbl is bf16 as well (autocast GEMM output, mamba_simple_faster.py:321-323)
template <bool XDB>
__device__ __forceinline__ float4 xd_ld4(const unsigned char* xd, uint32_t idx) {
    if (XDB) return ld4(reinterpret_cast<const bf16*>(xd) + idx);
    return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(xd) + idx);
}

// RT: dt_rank (12 / 24 / 48 templated, 0 = any multiple of 4 <= 16 chunks);  NORM: LayerNorm after the SSM;
// F14: 14 x 14 token grid (pool 14, 14 pooled rows: every loop unrolled);  XDB: xd held as bf16.
template <int RT, bool NORM, bool F14, bool XDB>
__global__ void __launch_bounds__(BC_THREADS, 2) block_cluster_kernel(const ClusterArgs a) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int N = BK_NSTATE, DC = BC_DC, T = BC_THREADS;
    const Geom& g = a.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = g.L, D = g.D;
    const int outer = F14 ? 14 : g.outer;
    const int P = F14 ? 14 : g.pool;
    const int C = a.C;
    const int rank = blockIdx.x;  // == %cluster_ctarank: cluster dims are (C, 1, 1) and gridDim.x == C
    const int img = blockIdx.y;
    const int dbase = rank * DC;  // first channel of this CTA's slab
    constexpr uint32_t rowB = DC * 2u;

    unsigned char* slab = smem;                                       // (L + 6) rows of DC bf16; token t at row t + 3
    bf16* ubuf = reinterpret_cast<bf16*>(smem + a.off_u);             // [2][outer][uld]
    float* ssum = reinterpret_cast<float*>(smem + a.off_s);           // [outer][DC]: s_f + s_b
    float* xpart = reinterpret_cast<float*>(smem + a.off_xp);         // [2][outer][xld] fp32: this slab's x_proj partial
    unsigned char* xd = smem + a.off_xd;                              // [2][outer][xld]: full x_dbl (fp32 | bf16)
    // [L] this slab's per-token (sum, sum of squares): written after the scan, so it reuses the then-dead u buffer
    float2* stats = reinterpret_cast<float2*>(smem + a.off_st);

    const bf16* xb = a.x + (int64_t)img * a.xzbs + dbase;
    const bf16* zb = a.z + (int64_t)img * a.xzbs + dbase;

    // ---- load: the image's x slab by 16-byte cp.async; halo rows zeroed; z rows pulled towards L2.
    // A thread keeps its 16-byte column and walks the token rows with a stride of T / cpr rows: (o, p) advance
    // incrementally, no division in the loop.
    const int so32 = (int)g.so, sp32 = (int)g.sp;   // row strides fit 32 bits (plan_cluster checks L * ld < 2^31)
    {
        constexpr int cpr = DC / 8, RSTEP = T / cpr;   // 24 chunks per row, 16 rows per sweep
        static_assert(T % cpr == 0, "thread count must be a multiple of the chunks per row");
        const int c = tid % cpr;
        int t = tid / cpr;
        int o = t / P, p = t - o * P;
        const bf16* xcol = xb + c * 8;
        unsigned char* dst = slab + (uint32_t)(t + 3) * rowB + c * 16;
        for (; t < L; t += RSTEP, dst += RSTEP * rowB) {
            cp_async16(dst, xcol + (int64_t)((o * so32 + p * sp32)) * a.ldxz, true);
            p += RSTEP;
            while (p >= P) { p -= P; ++o; }
        }
        for (int i = tid; i < (int)(3 * rowB / 16); i += T) {
            reinterpret_cast<uint4*>(slab)[i] = make_uint4(0u, 0u, 0u, 0u);
            reinterpret_cast<uint4*>(slab + (size_t)(L + 3) * rowB)[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        // z: 3 x 128-byte lines per 384-byte row; thread = (line, row), rows strided by T / 3
        constexpr int ZSTEP = T / 3;
        const int zl = tid % 3;
        int tz = tid / 3;
        int oz = tz / P, pz = tz - oz * P;
        for (; tz < L; tz += ZSTEP) {
            prefetch_l2(zb + (int64_t)(oz * so32 + pz * sp32) * a.ldxz + zl * 64);
            pz += ZSTEP;
            while (pz >= P) { pz -= P; ++oz; }
        }
    }

    // ---- pass-1 mapping: (channel pair, quarter of the pooled rows)
    constexpr int half_d = DC / 2;
    const int p1_q = tid / half_d, p1_c = tid - p1_q * half_d;
    const int rpq = (outer + 3) >> 2;
    const int r_begin = min(outer, p1_q * rpq), r_end = min(outer, r_begin + rpq);
    const bool p1_work = r_begin < r_end;
    const uint32_t p1_off = (uint32_t)p1_c * 4u;
    float2 wf[4], wb[4], bf_, bb_, Df2, Db2;
    {
        const int d0 = dbase + p1_c * 2;
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
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();  // S0: the slab has landed

    // ================= pass 1: conv (both directions) + SiLU + mean pool; D-skip term w in place over x ===============
    uint32_t hl0 = 0, hl1 = 0, hl2 = 0;
    if (p1_work) {
        const uint32_t o = (uint32_t)(r_begin * P) * rowB + p1_off;  // token r_begin*P - 3
        hl0 = *reinterpret_cast<const uint32_t*>(smem + o);
        hl1 = *reinterpret_cast<const uint32_t*>(smem + o + rowB);
        hl2 = *reinterpret_cast<const uint32_t*>(smem + o + 2 * rowB);
    }
    __syncthreads();  // S1: left halos are in registers, nobody has overwritten x yet
    uint32_t dfr0 = 0, dfr1 = 0, dfr2 = 0;  // w of the first 3 tokens of the segment: stored after S2
    if (p1_work) {
        const float2 sc2 = make_float2(a.scale, a.scale);
        float2 win[7];
        win[0] = unpack2(hl0); win[1] = unpack2(hl1); win[2] = unpack2(hl2);
        uint32_t tok = (uint32_t)(r_begin * P + 3) * rowB + p1_off;  // byte offset of token r*P (this thread's pair)
        win[3] = unpack2(*reinterpret_cast<const uint32_t*>(smem + tok));
        win[4] = unpack2(*reinterpret_cast<const uint32_t*>(smem + tok + rowB));
        win[5] = unpack2(*reinterpret_cast<const uint32_t*>(smem + tok + 2 * rowB));
        uint32_t uo = ((uint32_t)r_begin * a.uld + p1_c * 2) * 2u;  // byte offset in ubuf
        const uint32_t udir = (uint32_t)outer * a.uld * 2u;
        for (int r = r_begin; r < r_end; ++r, tok += (uint32_t)P * rowB, uo += (uint32_t)a.uld * 2u) {
            float2 sumf = make_float2(0.f, 0.f), sumb = sumf;
            const bool first_row = r == r_begin;
#define BC_TOKEN(C_, X_)                                                                                   \
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
            if (F14) {
#define BC_XU(k_) win[(c + (k_)) % 7]
#pragma unroll
                for (int c = 0; c < 14; ++c) BC_TOKEN(c, BC_XU)
#undef BC_XU
            } else {
#define BC_XS(k_) win[(k_)]
                for (int c = 0; c < P; ++c) {
                    BC_TOKEN(c, BC_XS)
#pragma unroll
                    for (int k = 0; k < 6; ++k) win[k] = win[k + 1];
                }
#undef BC_XS
            }
#undef BC_TOKEN
            *reinterpret_cast<uint32_t*>(smem + a.off_u + uo) = pack2(__fmul2_rn(sumf, sc2));
            *reinterpret_cast<uint32_t*>(smem + a.off_u + uo + udir) = pack2(__fmul2_rn(sumb, sc2));
        }
    }
    // ---- x_proj work items (direction, 8-column tile), full K-slab each: the W_x fragments of the warp's first item do
    // not depend on the image -- fetch them BEFORE the barrier so their L2 latency overlaps the slower pass-1 threads.
    const int gq = lane >> 2, tq = lane & 3;
    const int nnt = a.nnt, xp_total = 2 * nnt;
    constexpr int KS = DC / 16;  // k-steps per item
    uint32_t xb0[KS], xb1[KS];
    {
        const int dir = warp / nnt, nt = warp - dir * nnt;
        const uint4* pw = a.xwp + ((size_t)((dir * nnt + nt) * C + rank) * (KS / 2)) * 32 + lane;
#pragma unroll
        for (int j = 0; j < KS / 2; ++j) {
            uint4 q = make_uint4(0u, 0u, 0u, 0u);
            if (warp < xp_total) q = __ldg(pw + j * 32);
            xb0[2 * j] = q.x; xb1[2 * j] = q.y; xb0[2 * j + 1] = q.z; xb1[2 * j + 1] = q.w;
        }
    }
    __syncthreads();  // S2: u complete; every right halo has been read
    if (p1_work) {
        const uint32_t tok = (uint32_t)(r_begin * P + 3) * rowB + p1_off;
        *reinterpret_cast<uint32_t*>(smem + tok) = dfr0;
        *reinterpret_cast<uint32_t*>(smem + tok + rowB) = dfr1;
        *reinterpret_cast<uint32_t*>(smem + tok + 2 * rowB) = dfr2;
    }

    // ================= x_proj partial on tensor cores: xpart[dir][j][c] = sum_{d in slab} u[dir][j][d] W_x[dir][c][d] =====
    for (int idx = warp; idx < xp_total; idx += BC_WARPS) {
        const int dir = idx / nnt, nt = idx - dir * nnt;
        const int row0 = gq, row1 = gq + 8;
        const bf16* pa0 = ubuf + ((uint32_t)dir * outer + min(row0, outer - 1)) * a.uld + tq * 2;
        const bf16* pa1 = ubuf + ((uint32_t)dir * outer + min(row1, outer - 1)) * a.uld + tq * 2;
        if (idx != warp) {  // later items (dt_rank > 16: more than 12 column tiles): fragments straight from L2
            const uint4* pw = a.xwp + ((size_t)((dir * nnt + nt) * C + rank) * (KS / 2)) * 32 + lane;
#pragma unroll
            for (int j = 0; j < KS / 2; ++j) {
                const uint4 q = __ldg(pw + j * 32);
                xb0[2 * j] = q.x; xb1[2 * j] = q.y; xb0[2 * j + 1] = q.z; xb1[2 * j + 1] = q.w;
            }
        }
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const uint32_t a0 = *reinterpret_cast<const uint32_t*>(pa0 + ks * 16);
            const uint32_t a1 = *reinterpret_cast<const uint32_t*>(pa1 + ks * 16);
            const uint32_t a2 = *reinterpret_cast<const uint32_t*>(pa0 + ks * 16 + 8);
            const uint32_t a3 = *reinterpret_cast<const uint32_t*>(pa1 + ks * 16 + 8);
            mma16816(acc, a0, a1, a2, a3, xb0[ks], xb1[ks]);
        }
        const int c0 = nt * 8 + tq * 2;
        if (row0 < outer) *reinterpret_cast<float2*>(xpart + ((uint32_t)dir * outer + row0) * a.xld + c0) = make_float2(acc[0], acc[1]);
        if (row1 < outer) *reinterpret_cast<float2*>(xpart + ((uint32_t)dir * outer + row1) * a.xld + c0) = make_float2(acc[2], acc[3]);
    }
    __syncthreads();  // S3: this CTA's partial is complete
    if (C > 1) {
        cluster_arrive();
        cluster_wait();  // every CTA's partial is visible cluster-wide
    }
    {
        // sum the C partials in rank order (same order on every CTA -> bit-identical x_dbl everywhere)
        const int n4 = 2 * outer * a.xld / 4;
        const uint32_t xp_addr = smem_u32(xpart);
        for (int i = tid; i < n4; i += T) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (C > 1) {
                for (int r = 0; r < C; ++r) {
                    const float4 v = ld_dsmem4(dsmem_addr(xp_addr + (uint32_t)i * 16u, (uint32_t)r));
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
            } else {
                acc = reinterpret_cast<const float4*>(xpart)[i];
            }
            if (XDB) st4(reinterpret_cast<bf16*>(xd) + 4 * i, acc);
            else reinterpret_cast<float4*>(xd)[i] = acc;
            if (a.xdbl_out && rank == 0) {  // saved for backward: (2, B*Lp, ncols) bf16
                const int e = 4 * i, rj = e / a.xld, c = e - rj * a.xld;  // rj = dir*outer + j
                const int dir = rj / outer, j = rj - dir * outer;
                bf16* o = a.xdbl_out + (((int64_t)dir * g.B + img) * outer + j) * a.ncols + c;
                const float vv[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (c + k < a.ncols) o[k] = __float2bfloat16_rn(vv[k]);
            }
        }
    }
    __syncthreads();  // S4: x_dbl complete

    // ================= bidirectional selective scan over the pooled rows =======================
    {
        const int dir = tid >= DC ? 1 : 0, d = tid - dir * DC;
        const int dg = dbase + d;  // global channel
        constexpr float LOG2E = 1.4426950408889634f;
        const uint32_t dirrow = (uint32_t)dir * outer;
        const int R = RT ? RT : a.R;
        // -- dt_proj pre-activations of every pooled row, in scan order; W_dt row passes through registers CH at a time
        float dpre[BC_MAXO];
        {
            const float bias = a.dtb[(int64_t)dir * D + dg];
#pragma unroll
            for (int s = 0; s < BC_MAXO; ++s) dpre[s] = bias;
            constexpr int CH = RT ? 12 : 4;
            const float* Wp = a.dtw + ((int64_t)dir * D + dg) * R;
            for (int c0 = 0; c0 < R; c0 += CH) {
                float Wd[CH];
#pragma unroll
                for (int k4 = 0; k4 < CH / 4; ++k4) {
                    const float4 q = __ldg(reinterpret_cast<const float4*>(Wp + c0) + k4);
                    Wd[4 * k4] = q.x; Wd[4 * k4 + 1] = q.y; Wd[4 * k4 + 2] = q.z; Wd[4 * k4 + 3] = q.w;
                }
#pragma unroll
                for (int s = 0; s < BC_MAXO; ++s) {
                    if (s < outer) {
                        const uint32_t j = dir ? outer - 1 - s : s;
                        const uint32_t xi = (dirrow + j) * a.xld + c0;
#pragma unroll
                        for (int k4 = 0; k4 < CH / 4; ++k4) {
                            const float4 q = xd_ld4<XDB>(xd, xi + 4 * k4);
                            dpre[s] = fmaf(Wd[4 * k4], q.x, dpre[s]); dpre[s] = fmaf(Wd[4 * k4 + 1], q.y, dpre[s]);
                            dpre[s] = fmaf(Wd[4 * k4 + 2], q.z, dpre[s]); dpre[s] = fmaf(Wd[4 * k4 + 3], q.w, dpre[s]);
                        }
                    }
                }
            }
        }
        float2 A2[N / 2], h[N / 2];
        {
            const float4* Ap = reinterpret_cast<const float4*>(a.A + ((int64_t)dir * D + dg) * N);
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
        const int64_t gplane = ((int64_t)dir * g.B + img) * outer;
#pragma unroll
        for (int s = 0; s < BC_MAXO; ++s) {
            if (s < outer) {
                const uint32_t j = dir ? outer - 1 - s : s;
                const bf16 ub = ubuf[(dirrow + j) * a.uld + d];
                const uint32_t xi = (dirrow + j) * a.xld + R;
                const float delta = bk_softplus(dpre[s]);
                if (a.pre_out) a.pre_out[(gplane + j) * D + dg] = dpre[s];
                const float du = delta * __bfloat162float(ub);
                float2 y2 = make_float2(0.f, 0.f);
                const float2 dl2 = make_float2(delta, delta), du2 = make_float2(du, du);
#pragma unroll
                for (int n = 0; n < N / 2; n += 2) {
                    const float4 Bq = xd_ld4<XDB>(xd, xi + 2 * n);
                    const float4 Cq = xd_ld4<XDB>(xd, xi + N + 2 * n);
                    float2 e0 = __fmul2_rn(dl2, A2[n]), e1 = __fmul2_rn(dl2, A2[n + 1]);
                    e0 = make_float2(bk_ex2(e0.x), bk_ex2(e0.y));
                    e1 = make_float2(bk_ex2(e1.x), bk_ex2(e1.y));
                    h[n] = __ffma2_rn(e0, h[n], __fmul2_rn(du2, make_float2(Bq.x, Bq.y)));
                    h[n + 1] = __ffma2_rn(e1, h[n + 1], __fmul2_rn(du2, make_float2(Bq.z, Bq.w)));
                    y2 = __ffma2_rn(h[n], make_float2(Cq.x, Cq.y), y2);
                    y2 = __ffma2_rn(h[n + 1], make_float2(Cq.z, Cq.w), y2);
                }
                const float yv = y2.x + y2.y;
                // the forward direction owns the plane during the scan; the backward direction parks its outputs in the
                // (consumed) dpre registers and adds them after a barrier -- a shared-memory fp32 atomic is a CAS loop
                if (dir == 0) ssum[j * DC + d] = yv;
                else dpre[s] = yv;
                if (a.s_out) a.s_out[(gplane + j) * D + dg] = yv;
                if (a.u_out) a.u_out[(gplane + j) * D + dg] = ub;
            }
        }
        __syncthreads();  // S5a: the forward direction's plane is complete
        if (dir == 1) {
#pragma unroll
            for (int s = 0; s < BC_MAXO; ++s)
                if (s < outer) ssum[(outer - 1 - s) * DC + d] += dpre[s];
        }
    }
    __syncthreads();  // S5: s = s_f + s_b complete

    // ================= gate: v = w + (s_f + s_b)/2, LayerNorm over ALL d_inner channels, * silu(z), store y ===========
    // 16 lanes per token (lane q of a half-warp owns the 4-channel chunks q, q+16, q+32 of the slab: 12 channels, so the
    // lane's LayerNorm weights stay in 24 registers), 2 tokens per warp instruction, 4-step butterflies.
    const int grp = lane >> 4, q16 = lane & 15;
    constexpr int TPR = BC_WARPS * 2;  // tokens per round of the CTA
    const int n_round = (L + TPR - 1) / TPR;
    const float2 half2c = make_float2(0.5f, 0.5f);
    if (NORM) {
        // ---- pass A: this slab's per-token (sum, sum of squares)
        int aj = (warp * 2 + grp) / P, ap = (warp * 2 + grp) - aj * P;
        for (int rd = 0; rd < n_round; ++rd) {
            const int t = (rd * BC_WARPS + warp) * 2 + grp;
            const uint32_t j = (uint32_t)aj;
            ap += TPR;
            while (ap >= P) { ap -= P; ++aj; }
            float sum = 0.f, sq = 0.f;
            if (t < L) {
                const unsigned char* wrow = smem + (uint32_t)(t + 3) * rowB + q16 * 8;
                const float* srow = ssum + j * DC + q16 * 4;
                float2 sum2 = make_float2(0.f, 0.f), sq2 = sum2;
#pragma unroll
                for (int i = 0; i < BC_NG; ++i) {
                    const uint2 wv = *reinterpret_cast<const uint2*>(wrow + 128 * i);
                    const float4 s0 = *reinterpret_cast<const float4*>(srow + 64 * i);
                    const float2 v0 = __ffma2_rn(half2c, make_float2(s0.x, s0.y), unpack2(wv.x));
                    const float2 v1 = __ffma2_rn(half2c, make_float2(s0.z, s0.w), unpack2(wv.y));
                    sum2 = __fadd2_rn(sum2, __fadd2_rn(v0, v1));
                    sq2 = __ffma2_rn(v0, v0, sq2);
                    sq2 = __ffma2_rn(v1, v1, sq2);
                }
                sum = sum2.x + sum2.y;
                sq = sq2.x + sq2.y;
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                sum += __shfl_xor_sync(0xffffffffu, sum, o);
                sq += __shfl_xor_sync(0xffffffffu, sq, o);
            }
            if (t < L && q16 == 0) stats[t] = make_float2(sum, sq);
        }
    }
    // z of the first round and the lane's LayerNorm weights: issued before the barriers so their latency overlaps them
    uint2 zc[BC_NG];
    int tj, tp;     // (pooled row, position in the row) of this lane's token, advanced incrementally (no division per round)
    {
        const int t = warp * 2 + grp;
        tj = t / P;
        tp = t - tj * P;
#pragma unroll
        for (int i = 0; i < BC_NG; ++i) zc[i] = make_uint2(0u, 0u);
        if (t < L) {
            const bf16* zrow = zb + (int64_t)(tj * so32 + tp * sp32) * a.ldxz + q16 * 4;
#pragma unroll
            for (int i = 0; i < BC_NG; ++i) zc[i] = __ldg(reinterpret_cast<const uint2*>(zrow + 64 * i));
        }
    }
    float2 gam[BC_NG][2], bet[BC_NG][2];
#pragma unroll
    for (int i = 0; i < BC_NG; ++i) {
        gam[i][0] = gam[i][1] = make_float2(1.f, 1.f);
        bet[i][0] = bet[i][1] = make_float2(0.f, 0.f);
        if (NORM) {
            const int c = dbase + (q16 + 16 * i) * 4;
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.lnw + c));
            gam[i][0] = make_float2(g0.x, g0.y); gam[i][1] = make_float2(g0.z, g0.w);
            if (a.lnb) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.lnb + c));
                bet[i][0] = make_float2(b0.x, b0.y); bet[i][1] = make_float2(b0.z, b0.w);
            }
        }
    }
    float2* mr = stats + L;   // [L] (-mean, rstd) over all d_inner channels (also in the dead u buffer)
    if (NORM) {
        __syncthreads();  // S6: this CTA's statistics are complete
        if (C > 1) {
            cluster_arrive();
            cluster_wait();
        }
        // every CTA gathers the C partials of every token ONCE (all DSMEM loads in flight together) and keeps the
        // token's (-mean, rstd) locally: the gate loop below then needs neither remote loads nor shuffles
        const float invD = 1.f / (float)D;
        const uint32_t st_addr = smem_u32(stats);
        for (int t = tid; t < L; t += T) {
            float sum = 0.f, sq = 0.f;
            if (C > 1) {
                float2 part[8];
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    part[r] = r < C ? ld_dsmem2(dsmem_addr(st_addr + (uint32_t)t * 8u, (uint32_t)r)) : make_float2(0.f, 0.f);
#pragma unroll
                for (int r = 0; r < 8; ++r) { sum += part[r].x; sq += part[r].y; }   // rank order: identical on every CTA
            } else {
                sum = stats[t].x; sq = stats[t].y;
            }
            const float mean = sum * invD;
            const float rstd = rsqrtf(fmaxf(fmaf(sq, invD, -mean * mean), 0.f) + a.eps);
            mr[t] = make_float2(-mean, rstd);
        }
        __syncthreads();  // S7
    }
    {
        bf16* yb = a.y + (int64_t)img * a.ybs + dbase + q16 * 4;
        constexpr int PSTEP = TPR;
        for (int rd = 0; rd < n_round; ++rd) {
            const int t = (rd * BC_WARPS + warp) * 2 + grp;
            const int tn = t + TPR;
            // (j, p) of the next round's token
            int nj = tj, np = tp + PSTEP;
            while (np >= P) { np -= P; ++nj; }
            // next round's z, one round ahead through registers
            uint2 zn[BC_NG];
#pragma unroll
            for (int i = 0; i < BC_NG; ++i) zn[i] = make_uint2(0u, 0u);
            if (tn < L) {
                const bf16* zrow = zb + (int64_t)(nj * so32 + np * sp32) * a.ldxz + q16 * 4;
#pragma unroll
                for (int i = 0; i < BC_NG; ++i) zn[i] = __ldg(reinterpret_cast<const uint2*>(zrow + 64 * i));
            }
            if (t < L) {
                float2 gsc = make_float2(1.f, 1.f), nmean = make_float2(0.f, 0.f);
                if (NORM) {
                    const float2 m = mr[t];
                    nmean = make_float2(m.x, m.x);
                    gsc = make_float2(m.y, m.y);
                }
                const unsigned char* wrow = smem + (uint32_t)(t + 3) * rowB + q16 * 8;
                const float* srow = ssum + tj * DC + q16 * 4;
                bf16* yrow = yb + (int64_t)(tj * so32 + tp * sp32) * a.ldy;
                bf16* vrow = a.v_out ? a.v_out + ((int64_t)img * L + (tj * so32 + tp * sp32)) * D + dbase + q16 * 4 : nullptr;
#pragma unroll
                for (int i = 0; i < BC_NG; ++i) {
                    const uint2 wv = *reinterpret_cast<const uint2*>(wrow + 128 * i);
                    const float4 s0 = *reinterpret_cast<const float4*>(srow + 64 * i);
                    float2 o0 = __ffma2_rn(half2c, make_float2(s0.x, s0.y), unpack2(wv.x));
                    float2 o1 = __ffma2_rn(half2c, make_float2(s0.z, s0.w), unpack2(wv.y));
                    if (vrow) *reinterpret_cast<uint2*>(vrow + 64 * i) = make_uint2(pack2(o0), pack2(o1));
                    if (NORM) {
                        o0 = __ffma2_rn(__fmul2_rn(__fadd2_rn(o0, nmean), gsc), gam[i][0], bet[i][0]);
                        o1 = __ffma2_rn(__fmul2_rn(__fadd2_rn(o1, nmean), gsc), gam[i][1], bet[i][1]);
                    }
                    const float2 h0 = silu2_from_half(__fmul2_rn(unpack2(zc[i].x), half2c));
                    const float2 h1 = silu2_from_half(__fmul2_rn(unpack2(zc[i].y), half2c));
                    uint2 pk;
                    pk.x = pack2(__fmul2_rn(o0, h0));
                    pk.y = pack2(__fmul2_rn(o1, h1));
                    *reinterpret_cast<uint2*>(yrow + 64 * i) = pk;
                }
            }
            tj = nj; tp = np;
#pragma unroll
            for (int i = 0; i < BC_NG; ++i) zc[i] = zn[i];
        }
    }
    if (C > 1) {  // nobody may exit while a peer can still read its shared memory
        cluster_arrive();
        cluster_wait();
    }
}

// x_proj weights (2, ncols, D) bf16 -> MMA B-fragment order per channel slab: for work item (dir, column tile nt, slab),
// uint4 j of lane l holds {b0, b1} of k-steps 2j and 2j+1 (b0 = W[n = nt*8 + l/4][k .. k+1], b1 = the same at k + 8, with
// k = slab*192 + ks*16 + (l%4)*2), so a warp reads its 12 k-steps with six coalesced 16-byte loads per lane.
__global__ void __launch_bounds__(256)
pack_xproj_slab_kernel(const bf16* __restrict__ w, int ncols, int D, int n_nt, int C, uint4* __restrict__ out) {
    constexpr int NJ = BC_DC / 32;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 2 * n_nt * C * NJ * 32) return;
    const int lane = idx & 31, j = (idx >> 5) % NJ, item = (idx >> 5) / NJ;
    const int slab = item % C, nt = (item / C) % n_nt, dir = item / C / n_nt;
    const int n = nt * 8 + (lane >> 2);
    uint32_t r[4] = {0u, 0u, 0u, 0u};
    if (n < ncols) {
        const bf16* row = w + ((size_t)dir * ncols + n) * D + slab * BC_DC + (lane & 3) * 2;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = (2 * j + h) * 16;
            r[2 * h] = *reinterpret_cast<const uint32_t*>(row + k);
            r[2 * h + 1] = *reinterpret_cast<const uint32_t*>(row + k + 8);
        }
    }
    out[idx] = make_uint4(r[0], r[1], r[2], r[3]);
}

struct ClusterPlan {
    int ok;
    size_t smem;
    int xld, uld, nnt, C, xdb, off_u, off_s, off_xp, off_xd, off_st;
};

ClusterPlan plan_cluster(const fv_geom* g, int dtype, int R, int N, int64_t ldxz, int64_t ldy) {
    ClusterPlan p;
    p.ok = 0;
    if (dtype != FV_BF16 || g->inner != 1 || N != BK_NSTATE) return p;
    const int D = g->dim, outer = g->outer, pool = g->pool;
    if (D % BC_DC != 0 || pool < 4 || outer > BC_MAXO || R <= 0 || R % 4 != 0) return p;
    p.C = D / BC_DC;
    if (p.C != 1 && p.C != 2 && p.C != 4 && p.C != 8) return p;
    const int64_t L = (int64_t)outer * pool;
    if (L * (ldxz > ldy ? ldxz : ldy) >= (1ll << 31)) return p;
    const int ncols = R + 2 * N;
    p.nnt = (ncols + 7) / 8;
    p.xld = p.nnt * 8;
    p.uld = BC_DC + 8;
    p.xdb = R > 16;
    auto up = [](size_t v) { return (v + 127) / 128 * 128; };
    size_t off = up((size_t)(L + 6) * BC_DC * 2);
    p.off_u = (int)off;
    off = up(off + (size_t)2 * outer * p.uld * 2);
    p.off_s = (int)off;
    off = up(off + (size_t)outer * BC_DC * 4);
    p.off_xp = (int)off;
    off = up(off + (size_t)2 * outer * p.xld * 4);
    p.off_xd = (int)off;
    off = up(off + (size_t)2 * outer * p.xld * (p.xdb ? 2 : 4));
    p.off_st = p.off_u;  // the statistics alias the u buffer (dead once the scan has finished)
    if ((size_t)L * 16 > (size_t)2 * outer * p.uld * 2) return p;   // per-token partial sums + (-mean, rstd)
    p.smem = off;
    // two CTAs per SM: 228 KB of shared memory per SM, 1 KB reserved per resident CTA
    if (off > (size_t)(233472 - 2 * 1024) / 2) return p;
    p.ok = 1;
    return p;
}

int64_t pack_xproj_slab_bytes(int dim, int ncols) {
    if (dim % BC_DC != 0) return 0;
    return (int64_t)2 * ((ncols + 7) / 8) * (dim / BC_DC) * (BC_DC / 32) * 32 * 16;
}

int pack_xproj_slab(int dim, int ncols, const void* xproj_w, void* packed, cudaStream_t st) {
    const int n_nt = (ncols + 7) / 8, C = dim / BC_DC, total = 2 * n_nt * C * (BC_DC / 32) * 32;
    pack_xproj_slab_kernel<<<(total + 255) / 256, 256, 0, st>>>((const bf16*)xproj_w, ncols, dim, n_nt, C, (uint4*)packed);
    return finish_launch("block_pack_xproj_slab");
}

int launch_block_cluster(const fv_geom* g_, const ClusterPlan& p, const void* x, const void* z, int64_t ldxz, int64_t xz_bstride,
                         const float* conv_w, const float* conv_b, const void* xw_slab_packed, const float* dt_w,
                         const float* dt_bias, const float* A, int a_is_log, int dt_rank, int dstate, const float* Dskip,
                         const float* ln_w, const float* ln_b, float eps, float scale, void* y, int64_t ldy, int64_t y_bstride,
                         void* u_out, void* xdbl_out, float* s_out, void* v_out, float* pre_out, cudaStream_t stream) {
    FV_REQUIRE(ldxz % 8 == 0 && xz_bstride % 8 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)z % 16) == 0,
               "fv_block_fwd: x / z rows must be 16-byte aligned (ldxz %lld)", (long long)ldxz);
    FV_REQUIRE(ldy % 8 == 0 && y_bstride % 8 == 0 && ((uintptr_t)y % 16) == 0, "fv_block_fwd: y rows must be 16-byte aligned");
    FV_REQUIRE(((uintptr_t)dt_w % 16) == 0 && ((uintptr_t)A % 16) == 0 && ((uintptr_t)conv_w % 16) == 0,
               "fv_block_fwd: fp32 parameter arrays must be 16-byte aligned");
    FV_REQUIRE(!ln_w || (((uintptr_t)ln_w % 16) == 0 && (!ln_b || ((uintptr_t)ln_b % 16) == 0)), "fv_block_fwd: LayerNorm parameters must be 16-byte aligned");
    ClusterArgs a;
    a.g = make_geom(g_);
    a.x = (const bf16*)x; a.z = (const bf16*)z; a.ldxz = ldxz; a.xzbs = xz_bstride;
    a.cw = conv_w; a.cb = conv_b; a.xwp = (const uint4*)xw_slab_packed; a.dtw = dt_w; a.dtb = dt_bias;
    a.A = A; a.a_is_log = a_is_log; a.Dskip = Dskip; a.lnw = ln_w; a.lnb = ln_b; a.eps = eps;
    a.scale = scale / (float)g_->pool;
    a.y = (bf16*)y; a.ldy = ldy; a.ybs = y_bstride;
    a.u_out = (bf16*)u_out; a.xdbl_out = (bf16*)xdbl_out; a.s_out = s_out; a.v_out = (bf16*)v_out; a.pre_out = pre_out;
    a.R = dt_rank; a.ncols = dt_rank + 2 * dstate; a.xld = p.xld; a.uld = p.uld; a.nnt = p.nnt; a.C = p.C;
    a.off_u = p.off_u; a.off_s = p.off_s; a.off_xp = p.off_xp; a.off_xd = p.off_xd; a.off_st = p.off_st;

    void (*kern)(const ClusterArgs) = nullptr;
    const bool f14 = g_->pool == 14 && g_->outer == 14;
#define FV_BC_PICK2(RT_, N_, XDB_) (f14 ? block_cluster_kernel<RT_, N_, true, XDB_> : block_cluster_kernel<RT_, N_, false, XDB_>)
#define FV_BC_PICK(RT_, XDB_) (ln_w ? FV_BC_PICK2(RT_, true, XDB_) : FV_BC_PICK2(RT_, false, XDB_))
    if (dt_rank == 12) kern = FV_BC_PICK(12, false);
    else if (dt_rank == 24) kern = FV_BC_PICK(24, true);
    else if (dt_rank == 48) kern = FV_BC_PICK(48, true);
    else if (p.xdb) kern = FV_BC_PICK(0, true);
    else kern = FV_BC_PICK(0, false);
#undef FV_BC_PICK2
#undef FV_BC_PICK
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    FV_REQUIRE(e == cudaSuccess, "fv_block_fwd: cudaFuncSetAttribute(%zu): %s", p.smem, cudaGetErrorString(e));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)p.C, (unsigned)g_->batch, 1);
    cfg.blockDim = dim3(BC_THREADS, 1, 1);
    cfg.dynamicSmemBytes = p.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    pdl_attr(&attr[1]);
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)p.C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    e = cudaLaunchKernelEx(&cfg, kern, a);
    FV_REQUIRE(e == cudaSuccess, "fv_block_fwd: cluster launch (%d CTAs / image): %s", p.C, cudaGetErrorString(e));
    return finish_launch("block_fwd[cluster]");
}

}  // namespace fv
