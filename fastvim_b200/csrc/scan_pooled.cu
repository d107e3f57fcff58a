// K2a -- bidirectional selective scan over the POOLED sequence, dt_proj + softplus fused.
//
// What it replaces in the reference (paths relative to /root/reference):
//   dt = dt_proj.weight @ dt.t(); rearrange; B/C .contiguous()       mamba_simple_faster.py:328-337, 384-394
//   selective_scan_fn(x_c, dt, A, B, C, D=None, z=None, delta_bias, delta_softplus=True)  :343-354, 397-410
//     -> selective_scan_fwd_kernel  csrc/selective_scan/selective_scan_fwd_kernel.cuh:67-303
// The reference launches grid (batch, dim) x 32 threads with 14 of 128 tile slots live at
// 224^2 and writes a (B, D, 1, 2N) fp32 checkpoint larger than its inputs.
// Here (v2): token-major pooled inputs; one thread per (image, channel, DIRECTION) holding the N
// fp32 states in registers -- the two directions are independent CTAs (grid.z), the backward one
// walking the un-flipped rows in descending order.  B/C and the low-rank dt rows of the image are
// staged per chunk of 16 pooled rows in shared memory and broadcast to the 128 channels of the
// CTA; the chunk's 16 u values are loaded as one batch into registers, so a chunk costs one
// memory round trip, not one per step.  Output: s[dir, b, j, d] in fp32 (2 planes; the epilogue
// kernel adds them), written with coalesced fire-and-forget stores -- no read-modify-write.
//
// Arithmetic per (b, d, j, dir): delta = softplus(bias + W_dt[d,:] . dt[j,:]);
//   h[n] = exp2(delta * A[d,n] * log2e) * h[n] + delta * B[j,n] * u[j,d];  y = sum_n h[n] C[j,n]
// (same exp2 formulation as fwd_kernel.cuh:169-171, 216).  MUFU-bound: N+2 SFU ops per pooled element.
#include <cstdlib>

#include "common.cuh"

namespace fv {

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// softplus with the reference threshold, 2 SFU ops: log(1 + e^x) = ln2 * log2(1 + 2^(x log2e)).
// dt_bias is initialised to softplus^-1([1e-3, 1e-1]) = [-6.9, -2.3] (mamba_simple_faster.py:111-130),
// where 1 + e^x loses the low bits of e^x: that range uses the series of log1p instead.
__device__ __forceinline__ float softplus_fast(float x) {
    constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
    const float e = ex2(x * LOG2E);
    // x < -5: log1p(e) = e (1 - e/2 + e^2/3) to 8e-8 relative; above, 1 + e is well conditioned
    const float sp = x < -5.f ? e * fmaf(e, fmaf(e, 0.33333334f, -0.5f), 1.f) : LN2 * lg2(1.f + e);
    return x <= 20.f ? sp : x;
}

// raw element -> fp32.  Prefetches keep the RAW loaded values in registers and convert at the point of use: with ld1()
// (load + convert) the compiler placed each convert right behind its load and re-used one register for successive loads, so
// the "independent" prefetch loads were issued one memory round trip apart (ncu SASS view: every SHF.L after an LDG.U16
// stalled on the long scoreboard, 47 % of all stall samples of scan_fwd at FastChannelVim-S).
__device__ __forceinline__ float raw2f(float v) { return v; }
__device__ __forceinline__ float raw2f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T raw_zero();
template <> __device__ __forceinline__ float raw_zero<float>() { return 0.f; }
template <> __device__ __forceinline__ bf16 raw_zero<bf16>() { return __ushort_as_bfloat16((unsigned short)0); }

constexpr int SCAN_THREADS = 128;
constexpr int SCAN_LC = 16;  // pooled rows per chunk

template <typename T, int RT, int N>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_fwd_kernel(Geom g, const T* __restrict__ u, const T* __restrict__ xdbl, int64_t ldxd, int R,
                const float* __restrict__ dtw, const float* __restrict__ dtb,
                const float* __restrict__ A, int a_is_log, float* __restrict__ s) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    constexpr int WROW = RT + 2 * N;
    __shared__ __align__(16) float tile[2][SCAN_LC][WROW];
    const int dir = blockIdx.z;
    const int b = blockIdx.y;
    const int d = blockIdx.x * SCAN_THREADS + threadIdx.x;
    const bool live = d < g.D;
    const int dd = live ? d : 0;
    const int nchunks = (g.Lp + SCAN_LC - 1) / SCAN_LC;
    const int64_t plane = (int64_t)g.B * g.Lp * g.D;
    constexpr float LOG2E = 1.4426950408889634f;

    float A2[N], W[RT], h[N];
    {
        const float* Ap = A + ((int64_t)dir * g.D + dd) * N;
#pragma unroll
        for (int n = 0; n < N; ++n) {
            float a = Ap[n];
            A2[n] = (a_is_log ? -expf(a) : a) * LOG2E;
            h[n] = 0.f;
        }
        const float* Wp = dtw + ((int64_t)dir * g.D + dd) * R;
#pragma unroll
        for (int j = 0; j < RT; ++j) W[j] = j < R ? Wp[j] : 0.f;
    }
    const float bias = dtb[(int64_t)dir * g.D + dd];
    const T* ub = u + dir * plane + (int64_t)b * g.Lp * g.D + dd;
    const T* xd = xdbl + ((int64_t)dir * g.B + b) * g.Lp * ldxd;
    float* sb = s + dir * plane + (int64_t)b * g.Lp * g.D + dd;
    const int step = dir == 0 ? 1 : -1;

    // The [dt | B | C] rows and the u values of chunk cc + 1 are fetched into registers BEFORE chunk cc is scanned and go to
    // shared memory after it: with the loads issued right before their use every chunk exposed a full global-memory round
    // trip (7 chunks on every CTA at FastChannelVim-S: 75 -> 73 us).  Measured dead ends at that shape (49 k chains of 112
    // steps, 10 warps per SM): splitting a chain's 16 states over 2 / 4 lanes (2x / 4x the warps, shared dt_proj dot product,
    // butterfly sums) ran 96 / 94 us, the two-pass chunk-parallel kernel below 151 us.
    constexpr int TPT = (SCAN_LC * WROW + SCAN_THREADS - 1) / SCAN_THREADS;   // tile elements per thread and chunk
    T tpre[TPT], upre[SCAN_LC];
    auto fetch = [&](int cc_) {
        const int chunk_ = dir == 0 ? cc_ : nchunks - 1 - cc_;
        const int rlo_ = chunk_ * SCAN_LC, rows_ = min(SCAN_LC, g.Lp - rlo_);
#pragma unroll
        for (int k = 0; k < TPT; ++k) {
            const int i = threadIdx.x + k * SCAN_THREADS;
            const int r = i / WROW, c = i - r * WROW;
            T v = raw_zero<T>();
            if (i < rows_ * WROW) {
                if (c < RT) {
                    if (c < R) v = __ldg(xd + (int64_t)(rlo_ + r) * ldxd + c);
                } else {
                    v = __ldg(xd + (int64_t)(rlo_ + r) * ldxd + R + (c - RT));
                }
            }
            tpre[k] = v;
        }
        const int r0_ = dir == 0 ? rlo_ : rlo_ + rows_ - 1;
#pragma unroll
        for (int rr = 0; rr < SCAN_LC; ++rr)
            upre[rr] = (rr < rows_ && live) ? __ldg(ub + (int64_t)(r0_ + rr * step) * g.D) : raw_zero<T>();
    };
    fetch(0);
#pragma unroll 1
    for (int cc = 0; cc < nchunks; ++cc) {
        const int chunk = dir == 0 ? cc : nchunks - 1 - cc;
        const int r_lo = chunk * SCAN_LC;
        const int rows = min(SCAN_LC, g.Lp - r_lo);
        float(*tl)[WROW] = tile[cc & 1];
#pragma unroll
        for (int k = 0; k < TPT; ++k) {
            const int i = threadIdx.x + k * SCAN_THREADS;
            if (i < SCAN_LC * WROW) (&tl[0][0])[i] = raw2f(tpre[k]);
        }
        const int r0 = dir == 0 ? r_lo : r_lo + rows - 1;
        float uu[SCAN_LC];
#pragma unroll
        for (int rr = 0; rr < SCAN_LC; ++rr) uu[rr] = raw2f(upre[rr]);
        if (cc + 1 < nchunks) fetch(cc + 1);
        __syncthreads();  // one barrier per chunk: tile[] is double-buffered
#pragma unroll
        for (int rr = 0; rr < SCAN_LC; ++rr) {
            if (rr < rows) {
                const int r = r0 + rr * step;
                const float* row = tl[r - r_lo];
                float dt = bias;
#pragma unroll
                for (int j = 0; j < RT; j += 4) {
                    float4 q = *reinterpret_cast<const float4*>(row + j);
                    dt = fmaf(W[j], q.x, dt); dt = fmaf(W[j + 1], q.y, dt);
                    dt = fmaf(W[j + 2], q.z, dt); dt = fmaf(W[j + 3], q.w, dt);
                }
                const float delta = softplus_fast(dt);
                const float du = delta * uu[rr];
                float y0 = 0.f, y1 = 0.f;
#pragma unroll
                for (int n = 0; n < N; n += 4) {
                    float4 Bq = *reinterpret_cast<const float4*>(row + RT + n);
                    float4 Cq = *reinterpret_cast<const float4*>(row + RT + N + n);
                    h[n] = fmaf(ex2(delta * A2[n]), h[n], du * Bq.x);             y0 = fmaf(h[n], Cq.x, y0);
                    h[n + 1] = fmaf(ex2(delta * A2[n + 1]), h[n + 1], du * Bq.y); y1 = fmaf(h[n + 1], Cq.y, y1);
                    h[n + 2] = fmaf(ex2(delta * A2[n + 2]), h[n + 2], du * Bq.z); y0 = fmaf(h[n + 2], Cq.z, y0);
                    h[n + 3] = fmaf(ex2(delta * A2[n + 3]), h[n + 3], du * Bq.w); y1 = fmaf(h[n + 3], Cq.w, y1);
                }
                if (live) sb[(int64_t)r * g.D] = y0 + y1;
            }
        }
    }
}

// ---- chunk-parallel variant for long pooled sequences with few images (2048^2: B = 1, Lp = 128) --------------
// The kernel above gives one thread the whole chain of Lp steps: at B = 1 that is 768 threads on 6 SMs and every
// step exposes the full LDS -> dt dot -> softplus -> ex2 -> FMA latency (72 us per launch, half of the 2048^2 step).
// Here the chain of each (channel, direction) is cut into NCH chunks of CL steps owned by different threads:
//   pass 1  every chunk scans from a zero state: S_k[n] (local end state) and sum_k(delta); since the decay of a
//           chunk is the product of exp2(delta*A2) = exp2(A2 * sum(delta)), no running product is kept;
//   combine every thread folds the (P, S) pairs of the chunks before it -> its start state (<= NCH-1 cheap steps);
//   pass 2  the chunk is scanned again from the true start state, producing y.
// Critical path 2*CL + NCH steps instead of Lp, and NCH times more warps to hide latency.  delta is computed once
// (pass 1) and kept in registers for pass 2.
constexpr int SCK_CL = 16;   // steps per chunk

template <typename T, int RT, int N, int CH>
__global__ void __launch_bounds__(CH * 16)
scan_fwd_chunked_kernel(Geom g, int nch, const T* __restrict__ u, const T* __restrict__ xdbl, int64_t ldxd, int R,
                        const float* __restrict__ dtw, const float* __restrict__ dtb,
                        const float* __restrict__ A, int a_is_log, float* __restrict__ s) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    constexpr int WROW = RT + 2 * N;
    extern __shared__ __align__(16) float sck_smem[];
    float* tile = sck_smem;                                    // [Lp][WROW]   dt | B | C rows of this image / direction
    float* carry = tile + (size_t)g.Lp * WROW;                 // [nch][CH][N]  local end states
    float* sumd = carry + (size_t)nch * CH * N;            // [nch][CH]     sum of delta per chunk
    const int dir = blockIdx.z, b = blockIdx.y;
    const int c = threadIdx.x & (CH - 1), k = threadIdx.x / CH;   // channel in CTA, chunk
    const int d = blockIdx.x * CH + c;
    const bool live = d < g.D;
    const int dd = live ? d : 0;
    const int64_t plane = (int64_t)g.B * g.Lp * g.D;
    constexpr float LOG2E = 1.4426950408889634f;
    const int Lp = g.Lp;

    // stage [dt | B | C] of the whole pooled sequence (fp32)
    const T* xd = xdbl + ((int64_t)dir * g.B + b) * Lp * ldxd;
    if (R % 4 == 0 && ldxd % 4 == 0 && ((uintptr_t)xd % (4 * sizeof(T))) == 0) {
        // 4-element groups, eight loads in flight per thread before the first shared-memory store (ncu on the scalar loop:
        // 54 % of the stalls on the global-load scoreboard -- 22 dependent load -> store round trips per thread)
        constexpr int GPR = WROW / 4, NB = 8;
        const int ngroups = Lp * GPR;
        for (int i0 = threadIdx.x; i0 < ngroups; i0 += blockDim.x * NB) {
            float4 v[NB];
#pragma unroll
            for (int q = 0; q < NB; ++q) {
                const int i = i0 + q * blockDim.x;
                v[q] = zero4();
                if (i < ngroups) {
                    const int r = i / GPR, gq = i - r * GPR;
                    const int col = gq < RT / 4 ? 4 * gq : R + 4 * (gq - RT / 4);
                    if (gq >= RT / 4 || 4 * gq < R) v[q] = ld4(xd + (int64_t)r * ldxd + col);
                }
            }
#pragma unroll
            for (int q = 0; q < NB; ++q) {
                const int i = i0 + q * blockDim.x;
                if (i < ngroups) *reinterpret_cast<float4*>(tile + (size_t)i * 4) = v[q];
            }
        }
    } else {
        for (int i = threadIdx.x; i < Lp * WROW; i += blockDim.x) {
            const int r = i / WROW, cc = i - r * WROW;
            float v = 0.f;
            if (cc < RT) {
                if (cc < R) v = ld1(xd + (int64_t)r * ldxd + cc);
            } else {
                v = ld1(xd + (int64_t)r * ldxd + R + (cc - RT));
            }
            tile[i] = v;
        }
    }
    float A2[N], h[N], W[RT];
    {
        const float* Ap = A + ((int64_t)dir * g.D + dd) * N;
#pragma unroll
        for (int n = 0; n < N; ++n) {
            const float a = Ap[n];
            A2[n] = (a_is_log ? -expf(a) : a) * LOG2E;
            h[n] = 0.f;
        }
        const float* Wp = dtw + ((int64_t)dir * g.D + dd) * R;
#pragma unroll
        for (int j = 0; j < RT; ++j) W[j] = j < R ? Wp[j] : 0.f;
    }
    const float bias = dtb[(int64_t)dir * g.D + dd];
    const T* ub = u + dir * plane + (int64_t)b * Lp * g.D + dd;
    float* sb = s + dir * plane + (int64_t)b * Lp * g.D + dd;
    // scan position p = k*CL + i  <->  pooled row j = dir ? Lp-1-p : p
    const int p0 = k * SCK_CL, steps = max(0, min(SCK_CL, Lp - p0));
    float uu[SCK_CL], dl[SCK_CL];
    {
        T uraw[SCK_CL];   // all loads first, conversions after (see raw2f)
#pragma unroll
        for (int i = 0; i < SCK_CL; ++i) {
            const int p = p0 + i, j = dir ? Lp - 1 - p : p;
            uraw[i] = (i < steps && live) ? __ldg(ub + (int64_t)j * g.D) : raw_zero<T>();
        }
#pragma unroll
        for (int i = 0; i < SCK_CL; ++i) uu[i] = raw2f(uraw[i]);
    }
    __syncthreads();
    // ---- pass 1: local scan from zero
    float sd = 0.f;
#pragma unroll
    for (int i = 0; i < SCK_CL; ++i) {
        dl[i] = 0.f;
        if (i < steps) {
            const int p = p0 + i, j = dir ? Lp - 1 - p : p;
            const float* row = tile + (size_t)j * WROW;
            float dt = bias;
#pragma unroll
            for (int q4 = 0; q4 < RT; q4 += 4) {
                const float4 q = *reinterpret_cast<const float4*>(row + q4);
                dt = fmaf(W[q4], q.x, dt); dt = fmaf(W[q4 + 1], q.y, dt);
                dt = fmaf(W[q4 + 2], q.z, dt); dt = fmaf(W[q4 + 3], q.w, dt);
            }
            const float delta = softplus_fast(dt);
            dl[i] = delta;
            sd += delta;
            const float du = delta * uu[i];
#pragma unroll
            for (int n = 0; n < N; n += 4) {
                const float4 Bq = *reinterpret_cast<const float4*>(row + RT + n);
                h[n] = fmaf(ex2(delta * A2[n]), h[n], du * Bq.x);
                h[n + 1] = fmaf(ex2(delta * A2[n + 1]), h[n + 1], du * Bq.y);
                h[n + 2] = fmaf(ex2(delta * A2[n + 2]), h[n + 2], du * Bq.z);
                h[n + 3] = fmaf(ex2(delta * A2[n + 3]), h[n + 3], du * Bq.w);
            }
        }
    }
    {
        float* cp = carry + ((size_t)k * CH + c) * N;
#pragma unroll
        for (int n = 0; n < N; n += 4) *reinterpret_cast<float4*>(cp + n) = make_float4(h[n], h[n + 1], h[n + 2], h[n + 3]);
        sumd[k * CH + c] = sd;
    }
    __syncthreads();
    // ---- combine: start state of chunk k = fold of chunks 0..k-1 (in scan order)
#pragma unroll
    for (int n = 0; n < N; ++n) h[n] = 0.f;
    for (int kk = 0; kk < k; ++kk) {
        const float sdk = sumd[kk * CH + c];
        const float* cp = carry + ((size_t)kk * CH + c) * N;
#pragma unroll
        for (int n = 0; n < N; n += 4) {
            const float4 Sq = *reinterpret_cast<const float4*>(cp + n);
            h[n] = fmaf(ex2(sdk * A2[n]), h[n], Sq.x);
            h[n + 1] = fmaf(ex2(sdk * A2[n + 1]), h[n + 1], Sq.y);
            h[n + 2] = fmaf(ex2(sdk * A2[n + 2]), h[n + 2], Sq.z);
            h[n + 3] = fmaf(ex2(sdk * A2[n + 3]), h[n + 3], Sq.w);
        }
    }
    // ---- pass 2: the chunk again, from the true start state
#pragma unroll
    for (int i = 0; i < SCK_CL; ++i) {
        if (i < steps) {
            const int p = p0 + i, j = dir ? Lp - 1 - p : p;
            const float* row = tile + (size_t)j * WROW;
            const float delta = dl[i], du = delta * uu[i];
            float y0 = 0.f, y1 = 0.f;
#pragma unroll
            for (int n = 0; n < N; n += 4) {
                const float4 Bq = *reinterpret_cast<const float4*>(row + RT + n);
                const float4 Cq = *reinterpret_cast<const float4*>(row + RT + N + n);
                h[n] = fmaf(ex2(delta * A2[n]), h[n], du * Bq.x);             y0 = fmaf(h[n], Cq.x, y0);
                h[n + 1] = fmaf(ex2(delta * A2[n + 1]), h[n + 1], du * Bq.y); y1 = fmaf(h[n + 1], Cq.y, y1);
                h[n + 2] = fmaf(ex2(delta * A2[n + 2]), h[n + 2], du * Bq.z); y0 = fmaf(h[n + 2], Cq.z, y0);
                h[n + 3] = fmaf(ex2(delta * A2[n + 3]), h[n + 3], du * Bq.w); y1 = fmaf(h[n + 3], Cq.w, y1);
            }
            if (live) sb[(int64_t)j * g.D] = y0 + y1;
        }
    }
}

int check_geom(const fv_geom* g, const char* who);
int sm_count();
template <typename T, int N>
static int launch_scan_chunked(const Geom& g, const T* u, const T* xdbl, int64_t ldxd, int R, const float* dtw,
                               const float* dtb, const float* A, int a_is_log, float* s, cudaStream_t st, bool* done) {
    *done = false;
    const int nch = ceil_div(g.Lp, SCK_CL);
    if (R > 24 || nch < 2 || nch > 16) return 0;   // <= 512 threads (128 registers each)
    const int RT = R <= 8 ? 8 : (R <= 12 ? 12 : (R <= 16 ? 16 : 24));
    // channels per CTA.  Fewer channels per CTA put more SMs to work (2048^2, 384 channels x 2 directions: 32 channels = 24
    // CTAs, 16 = 48, 8 = 96) but measured 15.8 / 16.0 / 19.2 us: the kernel is bound by each CTA's dependent chain (stage ->
    // pass 1 -> fold -> pass 2), not by the number of SMs, so 32 stays.
    int CH = 32;
    if (const char* e = getenv("FASTVIM_SCAN_CHUNK_CH")) CH = atoi(e) == 8 ? 8 : (atoi(e) == 16 ? 16 : 32);   // A/B timing
    const size_t smem = ((size_t)g.Lp * (RT + 2 * N) + (size_t)nch * CH * N + (size_t)nch * CH) * sizeof(float);
    if (smem > 200 * 1024) return 0;
    dim3 grid(ceil_div(g.D, CH), g.B, 2), block(CH * nch);
#define FV_SCK_LAUNCH(RT_, CH_)                                                                                       \
    {                                                                                                                 \
        auto kern = scan_fwd_chunked_kernel<T, RT_, N, CH_>;                                                          \
        if (smem > 48 * 1024) {                                                                                       \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
            FV_REQUIRE(e == cudaSuccess, "fv_scan_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));             \
        }                                                                                                             \
        FV_LAUNCH_PDL((kern), grid, block, smem, st, g, nch, u, xdbl, ldxd, R, dtw, dtb, A, a_is_log, s);             \
    }
#define FV_SCK_CASE(RT_)                         \
    if (RT == RT_) {                             \
        if (CH == 32) FV_SCK_LAUNCH(RT_, 32)     \
        else if (CH == 16) FV_SCK_LAUNCH(RT_, 16) \
        else FV_SCK_LAUNCH(RT_, 8)               \
    }
    FV_SCK_CASE(8) FV_SCK_CASE(12) FV_SCK_CASE(16) FV_SCK_CASE(24)
#undef FV_SCK_CASE
#undef FV_SCK_LAUNCH
    *done = true;
    return finish_launch("scan_fwd_chunked");
}

template <typename T, int N>
static int launch_scan(const Geom& g, const T* u, const T* xdbl, int64_t ldxd, int R, const float* dtw,
                       const float* dtb, const float* A, int a_is_log, float* s, cudaStream_t st) {
    // few images and a long pooled sequence: the one-thread-per-chain kernel would leave most SMs idle
    // (measured at FastChannelVim-S, 32 images x 112 pooled rows x 768 channels: chunked 151 us vs 74 us for the plain
    // kernel -- with enough chains to fill the SMs the two-pass form loses, so it stays reserved for the few-image case)
    if (g.Lp >= 2 * SCK_CL && (int64_t)ceil_div(g.D, SCAN_THREADS) * g.B * 2 < sm_count()) {
        bool done = false;
        const int rc = launch_scan_chunked<T, N>(g, u, xdbl, ldxd, R, dtw, dtb, A, a_is_log, s, st, &done);
        if (rc || done) return rc;
    }
    dim3 grid(ceil_div(g.D, SCAN_THREADS), g.B, 2), block(SCAN_THREADS);
#define FV_SCAN_CASE(RT_)                                                                          \
    if (R <= RT_) {                                                                                \
        FV_LAUNCH_PDL((scan_fwd_kernel<T, RT_, N>), grid, block, 0, st, g, u, xdbl, ldxd, R, dtw, dtb, A, a_is_log, s); \
        return finish_launch("scan_fwd");                                                          \
    }
    FV_SCAN_CASE(8) FV_SCAN_CASE(12) FV_SCAN_CASE(16) FV_SCAN_CASE(24) FV_SCAN_CASE(32) FV_SCAN_CASE(48) FV_SCAN_CASE(64)
#undef FV_SCAN_CASE
    return fail("fv_scan_fwd: dt_rank %d > 64 not supported", R);
}

}  // namespace fv

extern "C" int fv_scan_fwd(const fv_geom* g_, int dtype, const void* u, const void* xdbl,
                           int64_t ld_xdbl, int dt_rank, int dstate, const float* dt_w,
                           const float* dt_bias, const float* A, int a_is_log, float* s_out,
                           void* stream) {
    using namespace fv;
    if (int rc = check_geom(g_, "fv_scan_fwd")) return rc;
    FV_REQUIRE(u && xdbl && dt_w && dt_bias && A && s_out, "fv_scan_fwd: null pointer");
    FV_REQUIRE(dt_rank > 0 && ld_xdbl >= dt_rank + 2 * dstate, "fv_scan_fwd: ld_xdbl %lld < R+2N", (long long)ld_xdbl);
    FV_REQUIRE(g_->batch <= 65535, "fv_scan_fwd: batch > 65535");
    Geom g = make_geom(g_);
    cudaStream_t st = (cudaStream_t)stream;
#define FV_DISPATCH(T_)                                                                             \
    if (dstate == 16) return launch_scan<T_, 16>(g, (const T_*)u, (const T_*)xdbl, ld_xdbl, dt_rank, dt_w, dt_bias, A, a_is_log, s_out, st); \
    if (dstate == 8) return launch_scan<T_, 8>(g, (const T_*)u, (const T_*)xdbl, ld_xdbl, dt_rank, dt_w, dt_bias, A, a_is_log, s_out, st);   \
    return fail("fv_scan_fwd: d_state %d not supported (8 or 16)", dstate);
    if (dtype == FV_F32) { FV_DISPATCH(float) }
    if (dtype == FV_BF16) { FV_DISPATCH(bf16) }
#undef FV_DISPATCH
    return fail("fv_scan_fwd: unsupported dtype %d", dtype);
}
