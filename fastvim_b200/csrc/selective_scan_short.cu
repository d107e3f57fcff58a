// Operator-API selective scan for SHORT sequences (L <= 16, d_state 16): the shape FastVim actually runs, the pooled
// sequence of a 224 x 224 image has 14 steps.
//
// Replaces selective_scan_cuda.fwd / .bwd (mamba-1p1p1/csrc/selective_scan/selective_scan.cpp:226-336, 338-492; kernels
// selective_scan_fwd_kernel.cuh:67-303, selective_scan_bwd_kernel.cuh:75-489) behind selective_scan_fn
// (mamba_ssm/ops/selective_scan_interface.py:12-123) on the reference layout (batch, dim, L), L contiguous.
// The reference gives every (batch, channel) row its own CTA with 14 of 128 item slots live (344 us forward, 792 us
// backward at (256, 384, 14, 16) on this B200); this repo's general kernel (selective_scan.cu: one warp per row, 128-step
// chunks) is built for long sequences and only 1.3x better there, and its backward is slower.  For 14 steps there is
// nothing to scan in parallel -- the work is (rows x states), so:
//   forward   one THREAD per row, the 16 states in registers; a CTA owns 128 consecutive channels of one image, whose rows
//             are one contiguous (128 x L) block: loaded and stored coalesced through shared memory; B / C rows of the
//             image are staged once per CTA and read as broadcasts.
//   backward  8 threads per row (two states each, packed f32x2 recurrence), 32 rows per CTA: the forward states and decays
//             of all <= 16 steps stay in registers between the forward and the reverse sweep (no checkpoint tensor, no
//             recompute); du / d(delta) are summed over a row's 8 lanes by shuffles, dB / dC over the warp's 4 rows by
//             shuffles and over the CTA's warps in shared memory -> ONE atomic per (CTA, state, step) instead of one per
//             row (the reference: one per row, bwd_kernel.cuh:438-462).
// Same arithmetic as the general kernel (exp2 decays, softplus threshold 20, fp32 state).

#include <cstdlib>

#include "common.cuh"

namespace fv {

constexpr int SSS_N = 16, SSS_LMAX = 16;

__device__ __forceinline__ float sss_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---------------------------------------------------------------------------------------------------------------------
template <typename T, int ROWS>
__global__ void __launch_bounds__(ROWS)
ss_short_fwd_kernel(int dim, int L, int groups, const T* __restrict__ u, const T* __restrict__ delta,
                    const float* __restrict__ A, const T* __restrict__ Bm, const T* __restrict__ Cm,
                    const float* __restrict__ Dp, const T* __restrict__ z, const float* __restrict__ dbias, int softplus,
                    T* __restrict__ out, float* __restrict__ last_state) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    constexpr int N = SSS_N;
    __shared__ __align__(16) float sB[SSS_LMAX * N], sC[SSS_LMAX * N];   // [l][n]
    __shared__ T su[ROWS * SSS_LMAX], sd[ROWS * SSS_LMAX], sz[ROWS * SSS_LMAX];
    const int tid = threadIdx.x, b = blockIdx.y, d0 = blockIdx.x * ROWS;
    const int nrows = min(ROWS, dim - d0);
    const int64_t base = ((int64_t)b * dim + d0) * L;
    for (int i = tid; i < nrows * L; i += ROWS) {
        su[i] = u[base + i];
        sd[i] = delta[base + i];
        if (z) sz[i] = z[base + i];
    }
    const int grp = d0 / (dim / groups);
    const int64_t bc0 = ((int64_t)b * groups + grp) * N * L;
    for (int i = tid; i < N * L; i += ROWS) {
        const int n = i / L, l = i - n * L;
        sB[l * N + n] = ld1(Bm + bc0 + i);
        sC[l * N + n] = ld1(Cm + bc0 + i);
    }
    __syncthreads();
    if (tid < nrows) {
        const int d = d0 + tid;
        constexpr float LOG2E = 1.4426950408889634f;
        float A2[N], h[N];
#pragma unroll
        for (int q = 0; q < N / 4; ++q) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(A + (int64_t)d * N) + q);
            A2[4 * q] = v.x * LOG2E; A2[4 * q + 1] = v.y * LOG2E; A2[4 * q + 2] = v.z * LOG2E; A2[4 * q + 3] = v.w * LOG2E;
        }
#pragma unroll
        for (int n = 0; n < N; ++n) h[n] = 0.f;
        const float bias = dbias ? dbias[d] : 0.f, Dd = Dp ? Dp[d] : 0.f;
        for (int l = 0; l < L; ++l) {
            const float uv = ld1(su + tid * L + l);
            float x = ld1(sd + tid * L + l) + bias;
            if (softplus) x = softplus20(x);
            const float xu = x * uv;
            float y = 0.f;
#pragma unroll
            for (int q = 0; q < N / 4; ++q) {
                const float4 bq = *reinterpret_cast<const float4*>(sB + l * N + 4 * q);
                const float4 cq = *reinterpret_cast<const float4*>(sC + l * N + 4 * q);
                h[4 * q] = fmaf(sss_ex2(x * A2[4 * q]), h[4 * q], xu * bq.x);
                h[4 * q + 1] = fmaf(sss_ex2(x * A2[4 * q + 1]), h[4 * q + 1], xu * bq.y);
                h[4 * q + 2] = fmaf(sss_ex2(x * A2[4 * q + 2]), h[4 * q + 2], xu * bq.z);
                h[4 * q + 3] = fmaf(sss_ex2(x * A2[4 * q + 3]), h[4 * q + 3], xu * bq.w);
                y = fmaf(h[4 * q], cq.x, y); y = fmaf(h[4 * q + 1], cq.y, y);
                y = fmaf(h[4 * q + 2], cq.z, y); y = fmaf(h[4 * q + 3], cq.w, y);
            }
            float o = fmaf(Dd, uv, y);
            if (z) o *= silu_exact(ld1(sz + tid * L + l));
            st1(su + tid * L + l, o);   // the row's own u slot: nobody else reads it
        }
        if (last_state) {
            float4* ls = reinterpret_cast<float4*>(last_state + ((int64_t)b * dim + d) * N);
#pragma unroll
            for (int q = 0; q < N / 4; ++q) ls[q] = make_float4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
        }
    }
    __syncthreads();
    for (int i = tid; i < nrows * L; i += ROWS) out[base + i] = su[i];
}

// ---------------------------------------------------------------------------------------------------------------------
constexpr int SSB_ROWS = 32, SSB_THREADS = SSB_ROWS * 8;

template <typename T>
__global__ void __launch_bounds__(SSB_THREADS, 2)
ss_short_bwd_kernel(int dim, int L, int groups, const T* __restrict__ u, const T* __restrict__ delta,
                    const float* __restrict__ A, const T* __restrict__ Bm, const T* __restrict__ Cm,
                    const float* __restrict__ Dp, const T* __restrict__ z, const float* __restrict__ dbias, int softplus,
                    const T* __restrict__ dout, T* __restrict__ du, T* __restrict__ ddelta, float* __restrict__ dA,
                    float* __restrict__ dB, float* __restrict__ dC, float* __restrict__ dD, T* __restrict__ dz,
                    float* __restrict__ ddbias) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    constexpr int N = SSS_N, LM = SSS_LMAX, R = SSB_ROWS;
    __shared__ __align__(16) float sB[LM * N], sC[LM * N];            // [l][n]
    __shared__ float su[R * LM], sdl[R * LM], sraw[R * LM], sdy[R * LM], sgo[R * LM], szv[R * LM];   // [r][l] (stride LM)
    __shared__ float sdu[R * LM], sdd[R * LM], sdz[R * LM];
    __shared__ __align__(16) float part[(SSB_THREADS / 32) * LM * 2 * N];   // [warp][l][dB 16 | dC 16]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, d0 = blockIdx.x * R;
    const int nrows = min(R, dim - d0);
    const int64_t base = ((int64_t)b * dim + d0) * L;
    for (int i = tid; i < nrows * L; i += SSB_THREADS) {
        const int r = i / L, l = i - r * L;
        const float bias = dbias ? dbias[d0 + r] : 0.f;
        const float raw = ld1(delta + base + i) + bias;
        const float go = ld1(dout + base + i);
        const float zv = z ? ld1(z + base + i) : 0.f;
        su[r * LM + l] = ld1(u + base + i);
        sraw[r * LM + l] = raw;
        sdl[r * LM + l] = softplus ? softplus20(raw) : raw;
        sgo[r * LM + l] = go;
        szv[r * LM + l] = zv;
        sdy[r * LM + l] = z ? go * silu_exact(zv) : go;
    }
    const int grp = d0 / (dim / groups);
    const int64_t bc0 = ((int64_t)b * groups + grp) * N * L;
    for (int i = tid; i < N * L; i += SSB_THREADS) {
        const int n = i / L, l = i - n * L;
        sB[l * N + n] = ld1(Bm + bc0 + i);
        sC[l * N + n] = ld1(Cm + bc0 + i);
    }
    __syncthreads();
    const int r = tid >> 3, p = tid & 7;              // row within the CTA, state pair
    const bool rvalid = r < nrows;
    const int d = d0 + (rvalid ? r : 0);
    constexpr float LOG2E = 1.4426950408889634f;
    const float2 An = rvalid ? make_float2(A[(int64_t)d * N + 2 * p], A[(int64_t)d * N + 2 * p + 1]) : make_float2(0.f, 0.f);
    const float2 A2 = make_float2(An.x * LOG2E, An.y * LOG2E);
    const float Dd = (Dp && rvalid) ? Dp[d] : 0.f;
    const float* ur = su + r * LM;
    const float* dlr = sdl + r * LM;
    const float* dyr = sdy + r * LM;
    // ---- forward sweep: states and decays of every step stay in registers
    float2 hs[LM], as_[LM];
    {
        float2 h = make_float2(0.f, 0.f);
#pragma unroll
        for (int l = 0; l < LM; ++l) {
            hs[l] = make_float2(0.f, 0.f);
            as_[l] = make_float2(1.f, 1.f);
            if (l < L) {
                const float dl = rvalid ? dlr[l] : 0.f, uv = rvalid ? ur[l] : 0.f;
                const float2 Bq = *reinterpret_cast<const float2*>(sB + l * N + 2 * p);
                as_[l] = make_float2(sss_ex2(dl * A2.x), sss_ex2(dl * A2.y));
                h = __ffma2_rn(as_[l], h, __fmul2_rn(make_float2(dl * uv, dl * uv), Bq));
                hs[l] = h;
                if (z) {   // y (before the gate) is needed for dz: sum the 16 states over the row's 8 lanes
                    const float2 Cq = *reinterpret_cast<const float2*>(sC + l * N + 2 * p);
                    float yp = fmaf(h.x, Cq.x, h.y * Cq.y);
                    yp += __shfl_xor_sync(0xffffffffu, yp, 1);
                    yp += __shfl_xor_sync(0xffffffffu, yp, 2);
                    yp += __shfl_xor_sync(0xffffffffu, yp, 4);
                    if (p == 0 && rvalid) {
                        const float zv = szv[r * LM + l];
                        sdz[r * LM + l] = sgo[r * LM + l] * fmaf(Dd, uv, yp) * dsilu(zv);
                    }
                }
            }
        }
    }
    // ---- reverse sweep
    float2 G = make_float2(0.f, 0.f), dAacc = make_float2(0.f, 0.f);
    float dDacc = 0.f, dbacc = 0.f;
    float* pw = part + warp * LM * 2 * N;
#pragma unroll
    for (int l = LM - 1; l >= 0; --l) {
        if (l < L) {
            const float dl = rvalid ? dlr[l] : 0.f, uv = rvalid ? ur[l] : 0.f, dy = rvalid ? dyr[l] : 0.f;
            const float2 Bq = *reinterpret_cast<const float2*>(sB + l * N + 2 * p);
            const float2 Cq = *reinterpret_cast<const float2*>(sC + l * N + 2 * p);
            const float2 g = __ffma2_rn(Cq, make_float2(dy, dy), G);                  // gradient reaching h_l
            const float2 hm1 = l > 0 ? hs[l - 1] : make_float2(0.f, 0.f);
            const float2 tmp = __fmul2_rn(__fmul2_rn(g, hm1), as_[l]);               // g h_{l-1} a_l
            float ddp = fmaf(tmp.x, An.x, tmp.y * An.y) + (g.x * Bq.x + g.y * Bq.y) * uv;   // d delta_l (this pair)
            float dup = (g.x * Bq.x + g.y * Bq.y) * dl;                               // d u_l (this pair)
            dAacc = __ffma2_rn(tmp, make_float2(dl, dl), dAacc);
            float2 dBv = __fmul2_rn(g, make_float2(dl * uv, dl * uv));
            float2 dCv = __fmul2_rn(hs[l], make_float2(dy, dy));
            G = __fmul2_rn(as_[l], g);
            // row sums over the 8 lanes of the row
            ddp += __shfl_xor_sync(0xffffffffu, ddp, 1); dup += __shfl_xor_sync(0xffffffffu, dup, 1);
            ddp += __shfl_xor_sync(0xffffffffu, ddp, 2); dup += __shfl_xor_sync(0xffffffffu, dup, 2);
            ddp += __shfl_xor_sync(0xffffffffu, ddp, 4); dup += __shfl_xor_sync(0xffffffffu, dup, 4);
            if (p == 0 && rvalid) {
                const float raw = sraw[r * LM + l];
                const float dd = (softplus && raw <= 20.f) ? ddp * sigmoidf_(raw) : ddp;
                sdd[r * LM + l] = dd;
                sdu[r * LM + l] = fmaf(dy, Dd, dup);
                dbacc += dd;
                dDacc = fmaf(dy, uv, dDacc);
            }
            // dB / dC: sum over the warp's 4 rows (lanes with the same state pair)
            dBv.x += __shfl_xor_sync(0xffffffffu, dBv.x, 8); dBv.y += __shfl_xor_sync(0xffffffffu, dBv.y, 8);
            dCv.x += __shfl_xor_sync(0xffffffffu, dCv.x, 8); dCv.y += __shfl_xor_sync(0xffffffffu, dCv.y, 8);
            dBv.x += __shfl_xor_sync(0xffffffffu, dBv.x, 16); dBv.y += __shfl_xor_sync(0xffffffffu, dBv.y, 16);
            dCv.x += __shfl_xor_sync(0xffffffffu, dCv.x, 16); dCv.y += __shfl_xor_sync(0xffffffffu, dCv.y, 16);
            if (lane < 8) {
                *reinterpret_cast<float2*>(pw + l * 2 * N + 2 * p) = dBv;
                *reinterpret_cast<float2*>(pw + l * 2 * N + N + 2 * p) = dCv;
            }
        }
    }
    if (rvalid) {
        atomicAdd(dA + (int64_t)d * N + 2 * p, dAacc.x);
        atomicAdd(dA + (int64_t)d * N + 2 * p + 1, dAacc.y);
        if (p == 0) {
            if (dD) atomicAdd(dD + d, dDacc);
            if (ddbias) atomicAdd(ddbias + d, dbacc);
        }
    }
    __syncthreads();
    // ---- coalesced outputs
    for (int i = tid; i < nrows * L; i += SSB_THREADS) {
        const int rr = i / L, l = i - rr * L;
        st1(du + base + i, sdu[rr * LM + l]);
        st1(ddelta + base + i, sdd[rr * LM + l]);
        if (z) st1(dz + base + i, sdz[rr * LM + l]);
    }
    // dB / dC of this CTA's rows: sum the warps' partials, one atomic per (state, step)
    for (int i = tid; i < L * 2 * N; i += SSB_THREADS) {
        const int l = i / (2 * N), c = i - l * 2 * N;
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < SSB_THREADS / 32; ++w) acc += part[(w * LM + l) * 2 * N + c];
        const int n = c & (N - 1);
        atomicAdd((c < N ? dB : dC) + bc0 + (int64_t)n * L + l, acc);
    }
}

// FASTVIM_SCAN_SHORT=0 keeps the general warp-per-row kernels for every length (A/B timing)
static bool short_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("FASTVIM_SCAN_SHORT");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// rows per CTA of the forward kernel (0: not eligible)
int ss_short_fwd_rows(int dim, int64_t L, int dstate, int groups) {
    if (!short_enabled() || L > SSS_LMAX || dstate != SSS_N) return 0;
    const int per_group = dim / groups;
    if (groups == 1 || per_group % 128 == 0) return 128;
    if (per_group % 32 == 0) return 32;
    return 0;
}
bool ss_short_bwd_ok(int dim, int64_t L, int dstate, int groups) {
    if (!short_enabled() || L > SSS_LMAX || dstate != SSS_N) return false;
    return groups == 1 || (dim / groups) % SSB_ROWS == 0;
}

template <typename T>
int launch_ss_short_fwd(int rows, int batch, int dim, int L, int groups, const T* u, const T* delta, const float* A, const T* B,
                        const T* C, const float* D, const T* z, const float* dbias, int softplus, T* out, float* last,
                        cudaStream_t st) {
    dim3 grid((dim + rows - 1) / rows, batch);
    if (rows == 128)
        FV_LAUNCH_PDL((ss_short_fwd_kernel<T, 128>), grid, 128, 0, st, dim, L, groups, u, delta, A, B, C, D, z, dbias, softplus, out, last);
    else
        FV_LAUNCH_PDL((ss_short_fwd_kernel<T, 32>), grid, 32, 0, st, dim, L, groups, u, delta, A, B, C, D, z, dbias, softplus, out, last);
    return finish_launch("selective_scan_fwd[short]");
}
template <typename T>
int launch_ss_short_bwd(int batch, int dim, int L, int groups, const T* u, const T* delta, const float* A, const T* B, const T* C,
                        const float* D, const T* z, const float* dbias, int softplus, const T* dout, T* du, T* ddelta,
                        float* dA, float* dB, float* dC, float* dD, T* dz, float* ddbias, cudaStream_t st) {
    dim3 grid((dim + SSB_ROWS - 1) / SSB_ROWS, batch);
    FV_LAUNCH_PDL((ss_short_bwd_kernel<T>), grid, SSB_THREADS, 0, st, dim, L, groups, u, delta, A, B, C, D, z, dbias, softplus, dout, du,
                                                          ddelta, dA, dB, dC, dD, dz, ddbias);
    return finish_launch("selective_scan_bwd[short]");
}
template int launch_ss_short_fwd<float>(int, int, int, int, int, const float*, const float*, const float*, const float*, const float*,
                                        const float*, const float*, const float*, int, float*, float*, cudaStream_t);
template int launch_ss_short_fwd<bf16>(int, int, int, int, int, const bf16*, const bf16*, const float*, const bf16*, const bf16*,
                                       const float*, const bf16*, const float*, int, bf16*, float*, cudaStream_t);
template int launch_ss_short_bwd<float>(int, int, int, int, const float*, const float*, const float*, const float*, const float*,
                                        const float*, const float*, const float*, int, const float*, float*, float*, float*, float*,
                                        float*, float*, float*, float*, cudaStream_t);
template int launch_ss_short_bwd<bf16>(int, int, int, int, const bf16*, const bf16*, const float*, const bf16*, const bf16*,
                                       const float*, const bf16*, const float*, int, const bf16*, bf16*, bf16*, float*, float*,
                                       float*, float*, bf16*, float*, cudaStream_t);

}  // namespace fv
