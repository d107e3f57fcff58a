// K2a-bwd -- backward of the bidirectional pooled selective scan (dt_proj + softplus fused).
//
// What it replaces in the reference (paths relative to /root/reference):
//   SelectiveScanFn.backward -> selective_scan_cuda.bwd        mamba_ssm/ops/selective_scan_interface.py:59-102
//     -> selective_scan_bwd_kernel                             csrc/selective_scan/selective_scan_bwd_kernel.cuh:75-489
//   (math: SURVEY.md Appendix A "Backward of the scan"; bwd_kernel.cuh:244-296, 439-453)
// The reference walks 2048-element chunks last -> first, recomputes the forward states of a chunk
// from the checkpoint x[chunk-1] written by the forward kernel, runs a reverse block scan and
// pushes dB/dC through BlockExchange + fp32 atomics across the `dim` CTAs (:298-315).
//
// Here: same thread mapping as the forward (one thread per (image, channel, direction), 128
// channels per CTA, directions in grid.z).  Pass A (only when Lp > 16) re-runs the forward
// recurrence and checkpoints the 16 states at every 16-row chunk boundary in shared memory.
// Pass B walks the chunks in reverse scan order; per chunk and state it recomputes the 16 states
// of the chunk from the checkpoint (registers), then runs the reverse recurrence
//   g_i = C_i dy_i + a_{i+1} g_{i+1}
// accumulating du, d(delta), dA in registers.  dB/dC (sums over channels) are reduced across the
// 32 lanes with a recursive-halving butterfly (16 shuffles per 16 values instead of 80), summed
// over the CTA's 4 warps in shared memory and written as one partial plane per 128-channel CTA
// column -- no atomics, deterministic; fv_reduce_planes adds the planes.  dA_log and d(dt_bias)
// reduce over the batch with fp32 atomicAdd (as the reference does, :467-477).
//
// Outputs: du, dDelta (2, B, Lp, D) in the activation dtype; dBC planes (ncol, 2, B*Lp, 2N) fp32;
// dA_log (2, D, N), d_dt_bias (2, D) fp32 (accumulated: caller zero-fills).
// d(dt low-rank) = dDelta . W_dt and dW_dt = dDelta^T . dt are plain GEMMs done by the caller,
// as in the reference (selective_scan_interface.py:698-737).
#include "common.cuh"

namespace fv {

__device__ __forceinline__ float ex2b(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

constexpr int SB_THREADS = 128;
constexpr int SB_LC = 16;

// recursive-halving butterfly: on return lane L holds (in v[0]) the sum over all 32 lanes of
// element idx(L) = 8*b4 + 4*b3 + 2*b2 + b1 (b_k = bit k of L); lanes L and L^1 hold the same value.
__device__ __forceinline__ float butterfly16(float (&v)[16], int lane) {
    {
        const bool up = lane & 16;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float send = up ? v[k] : v[k + 8], keep = up ? v[k + 8] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    {
        const bool up = lane & 8;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float send = up ? v[k] : v[k + 4], keep = up ? v[k + 4] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    {
        const bool up = lane & 4;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float send = up ? v[k] : v[k + 2], keep = up ? v[k + 2] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    {
        const bool up = lane & 2;
        const float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}
__device__ __forceinline__ int butterfly_index(int lane) {
    return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

template <typename T, int RT, int N>
__global__ void __launch_bounds__(SB_THREADS, 2)
scan_bwd_kernel(Geom g, int nplanes_ds, const T* __restrict__ u, const T* __restrict__ xdbl, int64_t ldxd, int R,
                const float* __restrict__ dtw, const float* __restrict__ dtb, const float* __restrict__ A,
                int a_is_log, const float* __restrict__ ds, T* __restrict__ du, T* __restrict__ ddelta,
                float* __restrict__ dbc_planes, float* __restrict__ dA, float* __restrict__ dbias) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    static_assert(N == 16, "butterfly reduction is written for 16 states");
    constexpr int WROW = RT + 2 * N;
    extern __shared__ __align__(16) float smem[];
    // smem: tile[LC][WROW] | us[LC][128] | dys[LC][128] | red[4 warps][LC][2N] | ckpt[nchunks][N][128]
    float(*tile)[WROW] = reinterpret_cast<float(*)[WROW]>(smem);
    float* us = smem + SB_LC * WROW;
    float* dys = us + SB_LC * SB_THREADS;
    float* red = dys + SB_LC * SB_THREADS;
    float* ckpt = red + 4 * SB_LC * 2 * N;

    const int dir = blockIdx.z, b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = blockIdx.x * SB_THREADS + tid;
    const bool live = d < g.D;
    const int dd = live ? d : 0;
    const int Lp = g.Lp;
    const int nchunks = (Lp + SB_LC - 1) / SB_LC;
    const int64_t plane = (int64_t)g.B * Lp * g.D;
    constexpr float LOG2E = 1.4426950408889634f;

    float A2[N], Anat[N];
    {
        const float* Ap = A + ((int64_t)dir * g.D + dd) * N;
#pragma unroll
        for (int n = 0; n < N; ++n) {
            const float a = Ap[n];
            Anat[n] = a_is_log ? -expf(a) : a;
            A2[n] = Anat[n] * LOG2E;
        }
    }
    const float* Wp = dtw + ((int64_t)dir * g.D + dd) * R;
    const float bias = dtb[(int64_t)dir * g.D + dd];
    const T* ub = u + dir * plane + (int64_t)b * Lp * g.D + dd;
    const T* xd = xdbl + ((int64_t)dir * g.B + b) * Lp * ldxd;
    const float* dsb = ds + (int64_t)b * Lp * g.D + dd;
    T* dub = du + dir * plane + (int64_t)b * Lp * g.D + dd;
    T* ddb = ddelta + dir * plane + (int64_t)b * Lp * g.D + dd;
    const int step = dir == 0 ? 1 : -1;

    auto load_tile = [&](int r_lo, int rows) {
        for (int i = tid; i < rows * WROW; i += SB_THREADS) {
            const int r = i / WROW, c = i - r * WROW;
            float v = 0.f;
            if (c < RT) {
                if (c < R) v = ld1(xd + (int64_t)(r_lo + r) * ldxd + c);
            } else {
                v = ld1(xd + (int64_t)(r_lo + r) * ldxd + R + (c - RT));
            }
            tile[r][c] = v;
        }
    };
    // pre-activation of delta at tile row `tr`
    auto pre_of = [&](int tr) {
        const float* row = tile[tr];
        float dt = bias;
        for (int j = 0; j < R; ++j) dt = fmaf(__ldg(Wp + j), row[j], dt);
        return dt;
    };

    // ---------------- pass A: forward sweep, checkpoint states at chunk starts
    if (nchunks > 1) {
        float h[N];
#pragma unroll
        for (int n = 0; n < N; ++n) h[n] = 0.f;
        for (int cc = 0; cc < nchunks; ++cc) {
            const int chunk = dir == 0 ? cc : nchunks - 1 - cc;
            const int r_lo = chunk * SB_LC, rows = min(SB_LC, Lp - r_lo);
            __syncthreads();
            load_tile(r_lo, rows);
            __syncthreads();
#pragma unroll
            for (int n = 0; n < N; ++n) ckpt[(cc * N + n) * SB_THREADS + tid] = h[n];
            if (cc == nchunks - 1) break;
            const int r0 = dir == 0 ? r_lo : r_lo + rows - 1;
            for (int rr = 0; rr < rows; ++rr) {
                const int r = r0 + rr * step;
                const float* row = tile[r - r_lo];
                const float delta = softplus20(pre_of(r - r_lo));
                const float dlu = delta * (live ? ld1(ub + (int64_t)r * g.D) : 0.f);
#pragma unroll
                for (int n = 0; n < N; ++n) h[n] = fmaf(ex2b(delta * A2[n]), h[n], dlu * row[RT + n]);
            }
        }
    }

    // ---------------- pass B: reverse sweep
    float gcar[N], dAacc[N];
#pragma unroll
    for (int n = 0; n < N; ++n) gcar[n] = 0.f, dAacc[n] = 0.f;
    float dbias_acc = 0.f;
    for (int cc = nchunks - 1; cc >= 0; --cc) {
        const int chunk = dir == 0 ? cc : nchunks - 1 - cc;
        const int r_lo = chunk * SB_LC, rows = min(SB_LC, Lp - r_lo);
        const int r0 = dir == 0 ? r_lo : r_lo + rows - 1;  // first row of the chunk in scan order
        __syncthreads();
        load_tile(r_lo, rows);
        for (int i = tid; i < 4 * SB_LC * 2 * N; i += SB_THREADS) red[i] = 0.f;
        // per-thread chunk inputs in scan order: u, dy (sum of the incoming ds planes)
#pragma unroll
        for (int rr = 0; rr < SB_LC; ++rr) {
            float uv = 0.f, dyv = 0.f;
            if (rr < rows && live) {
                const int r = r0 + rr * step;
                uv = ld1(ub + (int64_t)r * g.D);
                for (int q = 0; q < nplanes_ds; ++q) dyv += dsb[q * plane + (int64_t)r * g.D];
            }
            us[rr * SB_THREADS + tid] = uv;
            dys[rr * SB_THREADS + tid] = dyv;
        }
        __syncthreads();
        float delta[SB_LC], ddl[SB_LC], dul[SB_LC];
#pragma unroll
        for (int rr = 0; rr < SB_LC; ++rr) {
            delta[rr] = rr < rows ? softplus20(pre_of((r0 + rr * step) - r_lo)) : 0.f;
            ddl[rr] = 0.f;
            dul[rr] = 0.f;
        }
#pragma unroll
        for (int n = 0; n < N; ++n) {
            float hl[SB_LC], al[SB_LC];
            float hprev = nchunks > 1 ? ckpt[(cc * N + n) * SB_THREADS + tid] : 0.f;
            const float h_in = hprev;
            // forward inside the chunk (delta = 0 past the end: a = 1, b = 0 keeps h)
#pragma unroll
            for (int rr = 0; rr < SB_LC; ++rr) {
                const int tr = rr < rows ? (r0 + rr * step) - r_lo : 0;
                al[rr] = ex2b(delta[rr] * A2[n]);
                hprev = fmaf(al[rr], hprev, delta[rr] * us[rr * SB_THREADS + tid] * tile[tr][RT + n]);
                hl[rr] = hprev;
            }
            // reverse
            float gn = gcar[n];  // = a_{i+1} g_{i+1} of the first row of the next chunk in scan order
            float dBv[SB_LC], dCv[SB_LC];
#pragma unroll
            for (int rr = SB_LC - 1; rr >= 0; --rr) {
                const int tr = rr < rows ? (r0 + rr * step) - r_lo : 0;
                const float Bn = tile[tr][RT + n], Cn = tile[tr][RT + N + n];
                const float dyv = dys[rr * SB_THREADS + tid], uv = us[rr * SB_THREADS + tid];
                const float gi = rr < rows ? fmaf(Cn, dyv, gn) : gn;
                const float hm1 = rr > 0 ? hl[rr - 1] : h_in;
                const float da_a = gi * hm1 * al[rr];  // dL/da * a
                dCv[rr] = rr < rows ? dyv * hl[rr] : 0.f;
                dBv[rr] = rr < rows ? gi * delta[rr] * uv : 0.f;
                dul[rr] = fmaf(gi * delta[rr], Bn, dul[rr]);
                ddl[rr] = fmaf(gi * Bn, uv, fmaf(da_a, Anat[n], ddl[rr]));
                dAacc[n] = fmaf(da_a, delta[rr], dAacc[n]);
                gn = rr < rows ? gi * al[rr] : gn;
            }
            gcar[n] = gn;
            // reduce dB / dC over the 32 channels of the warp; rows indexed in scan order
            const float sB = butterfly16(dBv, lane), sC = butterfly16(dCv, lane);
            if ((lane & 1) == 0) {
                const int rr = butterfly_index(lane);
                red[(warp * SB_LC + rr) * 2 * N + n] = sB;
                red[(warp * SB_LC + rr) * 2 * N + N + n] = sC;
            }
        }
        // per-row outputs
#pragma unroll
        for (int rr = 0; rr < SB_LC; ++rr) {
            if (rr < rows && live) {
                const int r = r0 + rr * step;
                const float pre = pre_of(r - r_lo);
                const float dpre = pre <= 20.f ? ddl[rr] * sigmoidf_(pre) : ddl[rr];
                dbias_acc += dpre;
                st1(dub + (int64_t)r * g.D, dul[rr]);
                st1(ddb + (int64_t)r * g.D, dpre);
            }
        }
        __syncthreads();
        // dB/dC partial plane of this 128-channel column: (ncol, 2, B*Lp, 2N)
        float* outp = dbc_planes + (((int64_t)blockIdx.x * 2 + dir) * g.B + b) * Lp * 2 * N;
        for (int i = tid; i < rows * 2 * N; i += SB_THREADS) {
            const int rr = i / (2 * N), c = i - rr * 2 * N;
            const float v = red[(0 * SB_LC + rr) * 2 * N + c] + red[(1 * SB_LC + rr) * 2 * N + c] +
                            red[(2 * SB_LC + rr) * 2 * N + c] + red[(3 * SB_LC + rr) * 2 * N + c];
            outp[(int64_t)(r0 + rr * step) * 2 * N + c] = v;
        }
    }
    if (live) {
        float* dAp = dA + ((int64_t)dir * g.D + d) * N;
#pragma unroll
        for (int n = 0; n < N; ++n) atomicAdd(dAp + n, a_is_log ? dAacc[n] * Anat[n] : dAacc[n]);
        atomicAdd(dbias + (int64_t)dir * g.D + d, dbias_acc);
    }
}

// ---- short pooled sequences (Lp <= 16: every 224^2 model) ----------------------------------------------------------
// The kernel above gives one thread a whole (channel, direction) chain and loops over the 16 states: 255 registers,
// 8 warps per SM, 32-step dependent chains -- 1.76 ms per launch at FastVim-B (half of the training step).
// Here a thread owns ONE state of one channel: 16 threads (a half-warp) per channel, 32 channels per CTA (512 threads).
//   rows   half-warp lane i prepares pooled row i of its channel: u, dy (sum of the ds planes), delta = softplus(bias +
//          W_dt . dt_i), shared with the other 15 lanes through shared memory (intra-warp: __syncwarp only);
//   fwd    h_s = a_s h_{s-1} + delta u B_n, keeping a_s and h_s of the <= 16 steps in registers (no recompute, no checkpoints);
//   rev    g_s = C_n dy + a_{s+1} g_{s+1}; per-step dB / dC are summed over the warp's two channels by one shuffle and
//          parked per warp in shared memory, then summed over the 16 warps -> one plane per 32-channel CTA (deterministic);
//          du / d(delta) partials are transposed through shared memory so that lane i sums the 16 states of row i;
//   out    du, d(delta) staged per CTA and written as 64-byte row segments.
constexpr int SBS_CH = 32, SBS_THREADS = SBS_CH * 16, SBS_LP = 16;

template <typename T, int RT>
__global__ void __launch_bounds__(SBS_THREADS, 2)
scan_bwd_small_kernel(Geom g, int nplanes_ds, const T* __restrict__ u, const T* __restrict__ xdbl, int64_t ldxd, int R,
                      const float* __restrict__ dtw, const float* __restrict__ dtb, const float* __restrict__ A,
                      int a_is_log, const float* __restrict__ ds, T* __restrict__ du, T* __restrict__ ddelta,
                      float* __restrict__ dbc_planes, float* __restrict__ dA, float* __restrict__ dbias) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    constexpr int N = 16, WROW = RT + 2 * N;
    extern __shared__ __align__(16) float sbs[];
    float* tile = sbs;                                   // [SBS_LP][WROW]      dt | B | C rows
    float* rowd = tile + SBS_LP * WROW;                  // [SBS_CH][4][SBS_LP] per channel: delta, pre, u, dy
    float* xch = rowd + SBS_CH * 4 * SBS_LP;             // [16 warps][2][SBS_LP][17]  per-warp exchange (du / ddelta partials)
    float* part = xch + 16 * 2 * SBS_LP * 17;            // [16 warps][SBS_LP][32]     dB | dC partials of each warp
    float* outs = part + 16 * SBS_LP * 32;               // [2][SBS_LP][SBS_CH]        staged du, ddelta
    const int dir = blockIdx.z, b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = tid & 15, c = tid >> 4;                // state, channel within the CTA
    const int hsel = (lane >> 4) & 1;                    // which of the warp's two channels
    const int d = blockIdx.x * SBS_CH + c;
    const bool live = d < g.D;
    const int dd = live ? d : 0;
    const int Lp = g.Lp;
    const int64_t plane = (int64_t)g.B * Lp * g.D;
    constexpr float LOG2E = 1.4426950408889634f;

    const T* xd = xdbl + ((int64_t)dir * g.B + b) * Lp * ldxd;
    for (int i = tid; i < Lp * WROW; i += SBS_THREADS) {
        const int r = i / WROW, cc = i - r * WROW;
        float v = 0.f;
        if (cc < RT) {
            if (cc < R) v = ld1(xd + (int64_t)r * ldxd + cc);
        } else {
            v = ld1(xd + (int64_t)r * ldxd + R + (cc - RT));
        }
        tile[i] = v;
    }
    const float a_raw = A[((int64_t)dir * g.D + dd) * N + n];
    const float Anat = a_is_log ? -expf(a_raw) : a_raw, A2 = Anat * LOG2E;
    __syncthreads();
    // ---- row data of this channel: lane n prepares pooled row n
    float* rd = rowd + c * 4 * SBS_LP;
    {
        float delta = 0.f, pre = 0.f, uv = 0.f, dyv = 0.f;
        if (n < Lp) {
            const float* row = tile + n * WROW;
            const float* Wp = dtw + ((int64_t)dir * g.D + dd) * R;
            pre = dtb[(int64_t)dir * g.D + dd];
            for (int j = 0; j < R; ++j) pre = fmaf(__ldg(Wp + j), row[j], pre);
            delta = softplus20(pre);
            if (live) {
                uv = ld1(u + dir * plane + ((int64_t)b * Lp + n) * g.D + dd);
                const float* dsb = ds + ((int64_t)b * Lp + n) * g.D + dd;
                for (int q = 0; q < nplanes_ds; ++q) dyv += dsb[q * plane];
            }
        }
        rd[0 * SBS_LP + n] = delta;
        rd[1 * SBS_LP + n] = pre;
        rd[2 * SBS_LP + n] = uv;
        rd[3 * SBS_LP + n] = dyv;
    }
    __syncwarp();
    // ---- forward: states and decays of every step stay in registers
    float hs[SBS_LP], as_[SBS_LP];
    {
        float h = 0.f;
#pragma unroll
        for (int s = 0; s < SBS_LP; ++s) {
            hs[s] = 0.f;
            as_[s] = 1.f;
            if (s < Lp) {
                const int i = dir ? Lp - 1 - s : s;
                const float delta = rd[i];
                as_[s] = ex2b(delta * A2);
                h = fmaf(as_[s], h, delta * rd[2 * SBS_LP + i] * tile[i * WROW + RT + n]);
                hs[s] = h;
            }
        }
    }
    // ---- reverse
    float gn = 0.f, dAacc = 0.f;
    float dul[SBS_LP], ddl[SBS_LP];
    float* pw = part + warp * SBS_LP * 32;
#pragma unroll
    for (int s = SBS_LP - 1; s >= 0; --s) {
        dul[s] = 0.f;
        ddl[s] = 0.f;
        if (s < Lp) {
            const int i = dir ? Lp - 1 - s : s;
            const float delta = rd[i], uv = rd[2 * SBS_LP + i], dyv = rd[3 * SBS_LP + i];
            const float Bn = tile[i * WROW + RT + n], Cn = tile[i * WROW + RT + N + n];
            const float gi = fmaf(Cn, dyv, gn);
            const float hm1 = s > 0 ? hs[s - 1] : 0.f;
            const float da_a = gi * hm1 * as_[s];
            float dCv = dyv * hs[s], dBv = gi * delta * uv;
            dul[s] = gi * delta * Bn;
            ddl[s] = fmaf(gi * Bn, uv, da_a * Anat);
            dAacc = fmaf(da_a, delta, dAacc);
            gn = gi * as_[s];
            dBv += __shfl_xor_sync(0xffffffffu, dBv, 16);   // the warp's two channels
            dCv += __shfl_xor_sync(0xffffffffu, dCv, 16);
            if (lane < 16) {
                pw[i * 32 + n] = dBv;
                pw[i * 32 + N + n] = dCv;
            }
        }
    }
    // ---- du / d(delta): sum over the 16 states of the channel (lane n ends up with row n)
    float* xw = xch + (warp * 2 + hsel) * SBS_LP * 17;
    float du_row = 0.f, dd_row = 0.f;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int s = 0; s < SBS_LP; ++s) {
            const int i = dir ? Lp - 1 - s : s;
            if (s < Lp) xw[i * 17 + n] = pass ? ddl[s] : dul[s];
        }
        __syncwarp();
        float acc = 0.f;
        if (n < Lp) {
#pragma unroll
            for (int k = 0; k < 16; ++k) acc += xw[n * 17 + k];
        }
        if (pass) dd_row = acc; else du_row = acc;
        __syncwarp();
    }
    float dpre = 0.f;
    if (n < Lp) {
        const float pre = rd[1 * SBS_LP + n];
        dpre = pre <= 20.f ? dd_row * sigmoidf_(pre) : dd_row;
        outs[(0 * SBS_LP + n) * SBS_CH + c] = du_row;
        outs[(1 * SBS_LP + n) * SBS_CH + c] = dpre;
    }
    // d(dt_bias) = sum over rows; dA (per state) over rows and images: fp32 atomics as in the reference (:467-477)
    float bsum = dpre;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
    if (live) {
        atomicAdd(dA + ((int64_t)dir * g.D + d) * N + n, a_is_log ? dAacc * Anat : dAacc);
        if (n == 0) atomicAdd(dbias + (int64_t)dir * g.D + d, bsum);
    }
    __syncthreads();
    // ---- coalesced outputs
    for (int t = tid; t < Lp * SBS_CH; t += SBS_THREADS) {
        const int i = t / SBS_CH, cc = t - i * SBS_CH;
        const int dch = blockIdx.x * SBS_CH + cc;
        if (dch < g.D) {
            const int64_t o = dir * plane + ((int64_t)b * Lp + i) * g.D + dch;
            st1(du + o, outs[(0 * SBS_LP + i) * SBS_CH + cc]);
            st1(ddelta + o, outs[(1 * SBS_LP + i) * SBS_CH + cc]);
        }
    }
    float* outp = dbc_planes + (((int64_t)blockIdx.x * 2 + dir) * g.B + b) * Lp * 2 * N;
    for (int t = tid; t < Lp * 32; t += SBS_THREADS) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < 16; ++w) acc += part[w * SBS_LP * 32 + t];
        outp[t] = acc;
    }
}

// ---- short pooled sequences, v2: two states per thread, dt_proj pre-activation supplied by the caller ----------------
// scan_bwd_small_kernel spends ~860 instructions per thread with 16 threads per (channel, direction): a third of them
// re-deriving delta (a 48-term dot product per row), the rest mostly shared-memory traffic -- 550 us per launch at
// FastVim-B.  Here
//   * the caller passes delta_pre = dt_bias + W_dt . dt (2, B, Lp, dim) fp32 -- the dt_proj GEMM, which the reference also
//     runs as a GEMM (mamba_simple_faster.py:328-334) -- so the kernel does no projection;
//   * a thread owns TWO states of one channel (8 threads per channel, 32 channels per 256-thread CTA) and all of its
//     recurrence arithmetic runs on the packed f32x2 pipe;
//   * per step a thread reads exactly two 16-byte shared-memory words: the channel's row record {delta, delta*u, dy, u}
//     and the state pair's {B, B', C, C'};
//   * h and the decays of the <= 16 steps stay in registers between the forward and the reverse sweep (no recompute);
//   * dB / dC partials go to shared memory as one 16-byte store per step and are summed over the CTA's 32 channels at
//     the end (one plane per CTA, deterministic); du / d(delta) partials are summed over a channel's 8 threads through
//     a warp-private shared-memory transpose.
// softplus in the form the forward kernels use (scan_pooled.cu softplus_fast / block_fwd.cu bk_softplus): 2 SFU ops
__device__ __forceinline__ float s2_softplus(float x) {
    constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
    const float e = ex2b(x * LOG2E);
    float l;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.f + e));
    const float sp = x < -5.f ? e * fmaf(e, fmaf(e, 0.33333334f, -0.5f), 1.f) : LN2 * l;
    return x <= 20.f ? sp : x;
}

constexpr int S2_CH = 32, S2_THREADS = S2_CH * 8, S2_LP = 16;

// dynamic shared memory, sized by the actual Lp (104 KB at Lp = 14: two CTAs per SM):
//   bc   [Lp][8] float4            per row, state pair: {B_2p, B_2p+1, C_2p, C_2p+1}
//   rowd [S2_CH][Lp] float4        per channel, row: {delta, delta*u, dy, u}
//   part [S2_CH][Lp][8] float4     per channel, row, state pair: {dB_2p, dB_2p+1, dC_2p, dC_2p+1}
//   xch  [S2_CH][Lp][9] float2     per channel, row, thread-of-channel: {du partial, d(delta) partial} (+1 pad)
//   prer [S2_CH][Lp] float         dt_proj pre-activation (for the softplus derivative)
//   outs [2][Lp][S2_CH] float      staged du, d(delta_pre)
static inline size_t s2_smem_bytes(int Lp) {
    return (size_t)Lp * (8 * 16 + S2_CH * 16 + S2_CH * 8 * 16 + S2_CH * 9 * 8 + S2_CH * 4 + 2 * S2_CH * 4);
}

// LPT: compile-time pooled length (14: every 224^2 model; 0: run-time g.Lp <= 16).  DIR: scan direction.  With both
// static every shared-memory address of the unrolled sweeps is base + constant (no index arithmetic, no guards).
template <typename T, int LPT, int DIR>
__device__ __forceinline__ void
scan_bwd_short_body(const Geom& g, int nplanes_ds, const T* __restrict__ u, const T* __restrict__ xdbl, int64_t ldxd, int R,
                    const float* __restrict__ pre, const float* __restrict__ A, int a_is_log,
                    const float* __restrict__ ds, T* __restrict__ du, T* __restrict__ ddelta,
                    float* __restrict__ dbc_planes, float* __restrict__ dA, float* __restrict__ dbias) {
    constexpr int N = 16;
    extern __shared__ __align__(16) unsigned char s2_raw[];
    const int Lp = LPT ? LPT : g.Lp;
    float4* s_bc = reinterpret_cast<float4*>(s2_raw);
    float4* s_rowd = s_bc + Lp * 8;
    float4* s_part = s_rowd + S2_CH * Lp;
    float2* s_xch = reinterpret_cast<float2*>(s_part + S2_CH * Lp * 8);
    float* s_prer = reinterpret_cast<float*>(s_xch + S2_CH * Lp * 9);
    float* s_outs = s_prer + S2_CH * Lp;
    constexpr int dir = DIR;
    const int b = blockIdx.y;
    const int tid = threadIdx.x, c = tid >> 3, p = tid & 7;   // channel within the CTA, state pair
    const int d = blockIdx.x * S2_CH + c;
    const bool live = d < g.D;
    const int dd = live ? d : 0;
    const int64_t plane = (int64_t)g.B * Lp * g.D;
    constexpr float LOG2E = 1.4426950408889634f;
    const float2 araw = *reinterpret_cast<const float2*>(A + ((int64_t)dir * g.D + dd) * N + 2 * p);   // latency hidden below

    // ---- B / C of the image's pooled rows and the row records.  Every global load of the prologue is issued BEFORE the first
    // conversion / store (raw values in registers): written as load -> convert -> store per item, each of the ~4 items of
    // a thread cost a full memory round trip and the prologue was half of the kernel (ncu SASS view: 31 % of all stall
    // samples on the converts / adds right behind these loads).
    const T* xd = xdbl + ((int64_t)dir * g.B + b) * Lp * ldxd + R;
    constexpr int NIT = (S2_LP * S2_CH + S2_THREADS - 1) / S2_THREADS;   // (row, channel) items per thread: 2
    T bcraw[4];
    const bool bc_live = tid < Lp * 8;
    {
        const int r = tid >> 3, pp = tid & 7;
        const T* row = xd + (int64_t)(bc_live ? r : 0) * ldxd;
#pragma unroll
        for (int q = 0; q < 4; ++q) bcraw[q] = __ldg(row + (q >> 1) * N + 2 * pp + (q & 1));
    }
    float pvr[NIT], dsr[NIT][4];
    T ur[NIT];
    int64_t off[NIT];
    bool it_live[NIT];
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
        const int i = tid + k * S2_THREADS;
        const int r = i >> 5, cc = i & 31;
        const int dch = blockIdx.x * S2_CH + cc;
        it_live[k] = i < Lp * S2_CH && dch < g.D;
        off[k] = it_live[k] ? ((int64_t)b * Lp + r) * g.D + dch : 0;
        pvr[k] = it_live[k] ? __ldg(pre + dir * plane + off[k]) : 0.f;
        ur[k] = __ldg(u + dir * plane + off[k]);
#pragma unroll
        for (int q = 0; q < 4; ++q) dsr[k][q] = (it_live[k] && q < nplanes_ds) ? __ldg(ds + q * plane + off[k]) : 0.f;
    }
    if (bc_live)
        s_bc[tid] = make_float4(ld1(&bcraw[0]), ld1(&bcraw[1]), ld1(&bcraw[2]), ld1(&bcraw[3]));
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
        const int i = tid + k * S2_THREADS;
        if (i < Lp * S2_CH) {
            const int r = i >> 5, cc = i & 31;
            float4 rec = make_float4(0.f, 0.f, 0.f, 0.f);
            float pv = 0.f;
            if (it_live[k]) {
                pv = pvr[k];
                const float uv = ld1(&ur[k]);
                float dyv = (dsr[k][0] + dsr[k][1]) + (dsr[k][2] + dsr[k][3]);
                for (int q0 = 4; q0 < nplanes_ds; q0 += 4) {   // more than four planes: not the training path's case
                    float v[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) v[q] = q0 + q < nplanes_ds ? ds[(q0 + q) * plane + off[k]] : 0.f;
                    dyv += (v[0] + v[1]) + (v[2] + v[3]);
                }
                const float delta = s2_softplus(pv);
                rec = make_float4(delta, delta * uv, dyv, uv);
            }
            s_rowd[cc * Lp + r] = rec;
            s_prer[cc * Lp + r] = pv;
        }
    }
    const float2 Anat = a_is_log ? make_float2(-expf(araw.x), -expf(araw.y)) : araw;
    const float2 A2 = make_float2(Anat.x * LOG2E, Anat.y * LOG2E);
    __syncthreads();

    // ---- forward sweep: states and decays of every step stay in registers
    const int i0 = DIR ? Lp - 1 : 0;
    constexpr int istep = DIR ? -1 : 1;
    float2 hs[S2_LP], as_[S2_LP];
    {
        float2 h = make_float2(0.f, 0.f);
#pragma unroll
        for (int s = 0; s < S2_LP; ++s) {
            hs[s] = make_float2(0.f, 0.f);
            as_[s] = make_float2(1.f, 1.f);
            if (s < Lp) {
                const int i = i0 + s * istep;
                const float4 rd = s_rowd[c * Lp + i];
                const float4 bcv = s_bc[i * 8 + p];
                const float2 ea = __fmul2_rn(make_float2(rd.x, rd.x), A2);
                as_[s] = make_float2(ex2b(ea.x), ex2b(ea.y));
                h = __ffma2_rn(as_[s], h, __fmul2_rn(make_float2(rd.y, rd.y), make_float2(bcv.x, bcv.y)));
                hs[s] = h;
            }
        }
    }
    // ---- reverse sweep
    float2 G = make_float2(0.f, 0.f), dAacc = make_float2(0.f, 0.f);
#pragma unroll
    for (int s = S2_LP - 1; s >= 0; --s) {
        if (s < Lp) {
            const int i = i0 + s * istep;
            const float4 rd = s_rowd[c * Lp + i];      // delta, delta*u, dy, u
            const float4 bcv = s_bc[i * 8 + p];        // B, B', C, C'
            const float2 dy2 = make_float2(rd.z, rd.z);
            const float2 gi = __ffma2_rn(make_float2(bcv.z, bcv.w), dy2, G);
            const float2 hm1 = s > 0 ? hs[s > 0 ? s - 1 : 0] : make_float2(0.f, 0.f);
            const float2 daa = __fmul2_rn(__fmul2_rn(gi, hm1), as_[s]);
            const float2 dC = __fmul2_rn(dy2, hs[s]);
            const float2 dB = __fmul2_rn(gi, make_float2(rd.y, rd.y));
            const float2 gB = __fmul2_rn(gi, make_float2(bcv.x, bcv.y));
            const float2 dul = __fmul2_rn(gB, make_float2(rd.x, rd.x));
            const float2 ddl = __ffma2_rn(gB, make_float2(rd.w, rd.w), __fmul2_rn(daa, Anat));
            dAacc = __ffma2_rn(daa, make_float2(rd.x, rd.x), dAacc);
            G = __fmul2_rn(gi, as_[s]);
            s_part[(c * Lp + i) * 8 + p] = make_float4(dB.x, dB.y, dC.x, dC.y);
            s_xch[(c * Lp + i) * 9 + p] = make_float2(dul.x + dul.y, ddl.x + ddl.y);
        }
    }
    __syncwarp();   // a channel's 8 threads sit in one warp
    // ---- du / d(delta): thread p sums the 8 partials of rows p and p + 8
    float bsum = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int r = p + 8 * h;
        if (r < Lp) {
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float2 v = s_xch[(c * Lp + r) * 9 + k];
                a0 += v.x;
                a1 += v.y;
            }
            const float pv = s_prer[c * Lp + r];
            const float dpre = pv <= 20.f ? a1 * sigmoidf_(pv) : a1;
            s_outs[r * S2_CH + c] = a0;
            s_outs[(Lp + r) * S2_CH + c] = dpre;
            bsum += dpre;
        }
    }
    // d(dt_bias) = sum over rows; dA per state over rows and images: fp32 atomics as in the reference (bwd_kernel.cuh:467-477)
    bsum += __shfl_xor_sync(0xffffffffu, bsum, 1);
    bsum += __shfl_xor_sync(0xffffffffu, bsum, 2);
    bsum += __shfl_xor_sync(0xffffffffu, bsum, 4);
    if (live) {
        float* dAp = dA + ((int64_t)dir * g.D + d) * N + 2 * p;
        atomicAdd(dAp, a_is_log ? dAacc.x * Anat.x : dAacc.x);
        atomicAdd(dAp + 1, a_is_log ? dAacc.y * Anat.y : dAacc.y);
        if (p == 0) atomicAdd(dbias + (int64_t)dir * g.D + d, bsum);
    }
    __syncthreads();
    // ---- coalesced outputs: du, d(delta_pre) as 32-channel row segments; [dB | dC] summed over the CTA's channels
    for (int t = tid; t < Lp * S2_CH; t += S2_THREADS) {
        const int i = t / S2_CH, cc = t - i * S2_CH;
        const int dch = blockIdx.x * S2_CH + cc;
        if (dch < g.D) {
            const int64_t o = dir * plane + ((int64_t)b * Lp + i) * g.D + dch;
            st1(du + o, s_outs[i * S2_CH + cc]);
            st1(ddelta + o, s_outs[(Lp + i) * S2_CH + cc]);
        }
    }
    // one thread per (row, state pair): 32 x LDS.128 + packed adds, then the plane's [dB | dC] row layout
    float* outp = dbc_planes + (((int64_t)blockIdx.x * 2 + dir) * g.B + b) * Lp * 2 * N;
    for (int t = tid; t < Lp * 8; t += S2_THREADS) {
        float2 accB = make_float2(0.f, 0.f), accC = make_float2(0.f, 0.f);
#pragma unroll 8
        for (int ch = 0; ch < S2_CH; ++ch) {
            const float4 v = s_part[ch * (Lp * 8) + t];
            accB = __fadd2_rn(accB, make_float2(v.x, v.y));
            accC = __fadd2_rn(accC, make_float2(v.z, v.w));
        }
        const int i = t >> 3, pp = t & 7;
        *reinterpret_cast<float2*>(outp + i * 32 + 2 * pp) = accB;
        *reinterpret_cast<float2*>(outp + i * 32 + N + 2 * pp) = accC;
    }
}

template <typename T, int LPT>
__global__ void __launch_bounds__(S2_THREADS, 2)
scan_bwd_short_kernel(Geom g, int nplanes_ds, const T* __restrict__ u, const T* __restrict__ xdbl, int64_t ldxd, int R,
                      const float* __restrict__ pre, const float* __restrict__ A, int a_is_log,
                      const float* __restrict__ ds, T* __restrict__ du, T* __restrict__ ddelta,
                      float* __restrict__ dbc_planes, float* __restrict__ dA, float* __restrict__ dbias) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    if (blockIdx.z)
        scan_bwd_short_body<T, LPT, 1>(g, nplanes_ds, u, xdbl, ldxd, R, pre, A, a_is_log, ds, du, ddelta, dbc_planes, dA, dbias);
    else
        scan_bwd_short_body<T, LPT, 0>(g, nplanes_ds, u, xdbl, ldxd, R, pre, A, a_is_log, ds, du, ddelta, dbc_planes, dA, dbias);
}

// sums `nplanes` planes of `n` floats: out[i] = sum_p in[p*n + i] (adds to the cast when `accumulate`)
template <typename TO>
__global__ void reduce_planes_kernel(const float* __restrict__ in, int nplanes, int64_t n, TO* __restrict__ out) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float acc = 0.f;
    for (int p = 0; p < nplanes; ++p) acc += in[p * n + i];
    st1(out + i, acc);
}

int check_geom(const fv_geom* g, const char* who);

template <typename T>
static int launch_scan_bwd(const Geom& g, int nplanes_ds, const T* u, const T* xdbl, int64_t ldxd, int R, int N,
                           const float* dtw, const float* dtb, const float* A, int a_is_log, const float* ds, T* du,
                           T* ddelta, float* dbc, float* dA, float* dbias, cudaStream_t st) {
    FV_REQUIRE(N == 16, "fv_scan_bwd: d_state %d not supported (16)", N);
    if (g.Lp <= SBS_LP) {  // short pooled sequences: one thread per (channel, state)
        dim3 grid(ceil_div(g.D, SBS_CH), g.B, 2), block(SBS_THREADS);
#define FV_SBS_CASE(RT_)                                                                                             \
    if (R <= RT_) {                                                                                                  \
        const size_t smem = sizeof(float) * ((size_t)SBS_LP * (RT_ + 32) + SBS_CH * 4 * SBS_LP + 16 * 2 * SBS_LP * 17 + \
                                             16 * SBS_LP * 32 + 2 * SBS_LP * SBS_CH);                                \
        auto kern = scan_bwd_small_kernel<T, RT_>;                                                                   \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
        FV_REQUIRE(e == cudaSuccess, "fv_scan_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));                \
        FV_LAUNCH_PDL((kern), grid, block, smem, st, g, nplanes_ds, u, xdbl, ldxd, R, dtw, dtb, A, a_is_log, ds, du, ddelta, dbc, \
                                        dA, dbias);                                                                  \
        return finish_launch("scan_bwd");                                                                            \
    }
        FV_SBS_CASE(12) FV_SBS_CASE(24) FV_SBS_CASE(48) FV_SBS_CASE(64)
#undef FV_SBS_CASE
        return fail("fv_scan_bwd: dt_rank %d > 64 not supported", R);
    }
    const int nchunks = ceil_div(g.Lp, SB_LC);
    dim3 grid(ceil_div(g.D, SB_THREADS), g.B, 2), block(SB_THREADS);
#define FV_SB_CASE(RT_)                                                                                              \
    if (R <= RT_) {                                                                                                  \
        const size_t smem = sizeof(float) * ((size_t)SB_LC * (RT_ + 32) + 2 * SB_LC * SB_THREADS + 4 * SB_LC * 32 +  \
                                             (nchunks > 1 ? (size_t)nchunks * 16 * SB_THREADS : 0));                 \
        FV_REQUIRE(smem <= 200 * 1024, "fv_scan_bwd: pooled length %d too long for the shared-memory checkpoints", g.Lp); \
        auto kern = scan_bwd_kernel<T, RT_, 16>;                                                                     \
        if (smem > 48 * 1024) {                                                                                      \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
            FV_REQUIRE(e == cudaSuccess, "fv_scan_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));            \
        }                                                                                                            \
        FV_LAUNCH_PDL((kern), grid, block, smem, st, g, nplanes_ds, u, xdbl, ldxd, R, dtw, dtb, A, a_is_log, ds, du, ddelta, dbc, \
                                        dA, dbias);                                                                  \
        return finish_launch("scan_bwd");                                                                            \
    }
    FV_SB_CASE(12) FV_SB_CASE(24) FV_SB_CASE(48) FV_SB_CASE(64)
#undef FV_SB_CASE
    return fail("fv_scan_bwd: dt_rank %d > 64 not supported", R);
}

}  // namespace fv

extern "C" int fv_scan_bwd(const fv_geom* g_, int dtype, int nplanes_ds, const void* u, const void* xdbl,
                           int64_t ld_xdbl, int dt_rank, int dstate, const float* dt_w, const float* dt_bias,
                           const float* A, int a_is_log, const float* ds, void* du, void* ddelta,
                           float* dbc_planes, float* dA, float* d_dt_bias, void* stream) {
    using namespace fv;
    if (int rc = check_geom(g_, "fv_scan_bwd")) return rc;
    FV_REQUIRE(u && xdbl && dt_w && dt_bias && A && ds && du && ddelta && dbc_planes && dA && d_dt_bias,
               "fv_scan_bwd: null pointer");
    FV_REQUIRE(dt_rank > 0 && ld_xdbl >= dt_rank + 2 * dstate, "fv_scan_bwd: ld_xdbl %lld < R+2N", (long long)ld_xdbl);
    FV_REQUIRE(g_->batch <= 65535 && nplanes_ds >= 1, "fv_scan_bwd: bad batch / plane count");
    Geom g = make_geom(g_);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FV_F32)
        return launch_scan_bwd<float>(g, nplanes_ds, (const float*)u, (const float*)xdbl, ld_xdbl, dt_rank, dstate, dt_w,
                                      dt_bias, A, a_is_log, ds, (float*)du, (float*)ddelta, dbc_planes, dA, d_dt_bias, st);
    if (dtype == FV_BF16)
        return launch_scan_bwd<bf16>(g, nplanes_ds, (const bf16*)u, (const bf16*)xdbl, ld_xdbl, dt_rank, dstate, dt_w,
                                     dt_bias, A, a_is_log, ds, (bf16*)du, (bf16*)ddelta, dbc_planes, dA, d_dt_bias, st);
    return fail("fv_scan_bwd: unsupported dtype %d", dtype);
}

/* number of [dB | dC] partial planes fv_scan_bwd writes for this geometry (the caller sizes dbc_planes with it) */
extern "C" int fv_scan_bwd_planes(const fv_geom* g) {
    if (!g || g->dim <= 0) return 0;
    const int Lp = g->outer * g->inner;
    return Lp <= fv::SBS_LP ? (g->dim + fv::SBS_CH - 1) / fv::SBS_CH : (g->dim + fv::SB_THREADS - 1) / fv::SB_THREADS;
}

extern "C" int fv_reduce_planes(int out_dtype, const float* in, int nplanes, int64_t n, void* out, void* stream) {
    using namespace fv;
    FV_REQUIRE(in && out && nplanes >= 1 && n > 0, "fv_reduce_planes: bad arguments");
    dim3 grid((unsigned)((n + 255) / 256)), block(256);
    cudaStream_t st = (cudaStream_t)stream;
    if (out_dtype == FV_F32) FV_LAUNCH_PDL((reduce_planes_kernel<float>), grid, block, 0, st, in, nplanes, n, (float*)out);
    else if (out_dtype == FV_BF16) FV_LAUNCH_PDL((reduce_planes_kernel<bf16>), grid, block, 0, st, in, nplanes, n, (bf16*)out);
    else return fail("fv_reduce_planes: unsupported dtype %d", out_dtype);
    return finish_launch("reduce_planes");
}

/* Short pooled sequences (Lp <= 16) with the dt_proj pre-activation supplied: see scan_bwd_short_kernel. */
extern "C" int fv_scan_bwd_short_supported(const fv_geom* g, int dstate) {
    return g && g->dim > 0 && g->outer > 0 && g->inner > 0 && g->outer * g->inner <= fv::S2_LP && dstate == 16;
}

extern "C" int fv_scan_bwd_short(const fv_geom* g_, int dtype, int nplanes_ds, const void* u, const void* xdbl,
                                 int64_t ld_xdbl, int dt_rank, int dstate, const float* delta_pre, const float* A,
                                 int a_is_log, const float* ds, void* du, void* ddelta, float* dbc_planes, float* dA,
                                 float* d_dt_bias, void* stream) {
    using namespace fv;
    if (int rc = check_geom(g_, "fv_scan_bwd_short")) return rc;
    FV_REQUIRE(u && xdbl && delta_pre && A && ds && du && ddelta && dbc_planes && dA && d_dt_bias,
               "fv_scan_bwd_short: null pointer");
    FV_REQUIRE(fv_scan_bwd_short_supported(g_, dstate), "fv_scan_bwd_short: needs Lp <= 16 and d_state 16");
    FV_REQUIRE(dt_rank > 0 && ld_xdbl >= dt_rank + 2 * dstate, "fv_scan_bwd_short: ld_xdbl %lld < R+2N", (long long)ld_xdbl);
    FV_REQUIRE(g_->batch <= 65535 && nplanes_ds >= 1, "fv_scan_bwd_short: bad batch / plane count");
    FV_REQUIRE(((uintptr_t)A % 8) == 0 && ((uintptr_t)dbc_planes % 8) == 0, "fv_scan_bwd_short: A / dbc_planes must be 8-byte aligned");
    Geom g = make_geom(g_);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(ceil_div(g.D, S2_CH), g.B, 2), block(S2_THREADS);
    const size_t smem = s2_smem_bytes(g.Lp);
    if (dtype == FV_F32) {
        auto kern = g.Lp == 14 ? scan_bwd_short_kernel<float, 14> : scan_bwd_short_kernel<float, 0>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        FV_REQUIRE(e == cudaSuccess, "fv_scan_bwd_short: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        FV_LAUNCH_PDL((kern), grid, block, smem, st, g, nplanes_ds, (const float*)u, (const float*)xdbl, ld_xdbl, dt_rank, delta_pre, A, a_is_log,
                                        ds, (float*)du, (float*)ddelta, dbc_planes, dA, d_dt_bias);
        return finish_launch("scan_bwd_short");
    }
    if (dtype == FV_BF16) {
        auto kern = g.Lp == 14 ? scan_bwd_short_kernel<bf16, 14> : scan_bwd_short_kernel<bf16, 0>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        FV_REQUIRE(e == cudaSuccess, "fv_scan_bwd_short: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        FV_LAUNCH_PDL((kern), grid, block, smem, st, g, nplanes_ds, (const bf16*)u, (const bf16*)xdbl, ld_xdbl, dt_rank, delta_pre, A, a_is_log,
                                        ds, (bf16*)du, (bf16*)ddelta, dbc_planes, dA, d_dt_bias);
        return finish_launch("scan_bwd_short");
    }
    return fail("fv_scan_bwd_short: unsupported dtype %d", dtype);
}
