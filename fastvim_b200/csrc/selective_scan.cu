// Operator-API selective scan on the reference layout (batch, dim, L) with L contiguous.
//
// Replaces selective_scan_cuda.fwd/.bwd  (mamba-1p1p1/csrc/selective_scan/selective_scan.cpp:226-336,
// 338-492; kernels selective_scan_fwd_kernel.cuh:67-303, selective_scan_bwd_kernel.cuh:75-489) behind
// selective_scan_fn (mamba_ssm/ops/selective_scan_interface.py:12-123).  Real A, variable B/C
// with groups, optional D, z, delta_bias, softplus, last_state.
//
// Design (not the reference's CUB BlockScan over one (b, d) row per CTA): one WARP per (b, d)
// row; the sequence is walked in chunks of 32 lanes x ITEMS consecutive timesteps.  For each
// state n a lane composes its ITEMS steps locally into an affine map (P, S): h_out = P*h_in + S,
// the 32 maps are combined with a 5-step warp-shuffle inclusive scan, the chunk carry is added,
// and the lane replays its steps to emit y += C*h.  The running carry of state n lives in lane n's
// register (N <= 32), so the kernel needs no shared memory and no __syncthreads; the 4 warps of a
// CTA work on 4 consecutive channels of one image so the B/C rows they share stay in L1.
// No per-chunk checkpoint tensor is written (the reference writes (B, D, nchunks, 2N) fp32 even at
// L = 14, selective_scan.cpp:307-313); the backward pass re-runs the forward recurrence instead.
#include "common.cuh"

namespace fv {

__device__ __forceinline__ float ex2f_(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

int sm_count();

constexpr int SS_WARPS = 4;
constexpr int SS_ITEMS = 4;

template <typename T>
__device__ __forceinline__ void load_items(const T* row, int64_t L, int64_t t0, bool vec, float (&v)[SS_ITEMS]) {
    if (vec && t0 + SS_ITEMS <= L) {
        float4 q = ld4(row + t0);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
#pragma unroll
        for (int k = 0; k < SS_ITEMS; ++k) v[k] = (t0 + k < L) ? ld1(row + t0 + k) : 0.f;
    }
}

// Four consecutive elements kept RAW in registers (converted at use): the B / C rows of a batch of states are fetched
// before the batch's scans start, so their L2 round trips overlap the shuffle scans instead of preceding each of them
// (with load_items -- load + convert -- inside the state loop every second state exposed a memory round trip).
template <typename T> struct Raw4;
template <> struct Raw4<bf16> {
    uint2 v;
    __device__ __forceinline__ void ld(const bf16* p) { v = __ldg(reinterpret_cast<const uint2*>(p)); }
    __device__ __forceinline__ void get(float (&o)[SS_ITEMS]) const {
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v.x), b = *reinterpret_cast<const __nv_bfloat162*>(&v.y);
        const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
        o[0] = fa.x; o[1] = fa.y; o[2] = fb.x; o[3] = fb.y;
    }
};
template <> struct Raw4<float> {
    float4 v;
    __device__ __forceinline__ void ld(const float* p) { v = __ldg(reinterpret_cast<const float4*>(p)); }
    __device__ __forceinline__ void get(float (&o)[SS_ITEMS]) const { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
};
constexpr int SS_NB = 8;   // states per prefetch batch

template <typename T, bool PRE>
__global__ void __launch_bounds__(SS_WARPS * 32)
selective_scan_fwd_kernel(int batch, int dim, int64_t L, int N, int groups, const T* __restrict__ u,
                          const T* __restrict__ delta, const float* __restrict__ A,
                          const T* __restrict__ Bm, const T* __restrict__ Cm,
                          const float* __restrict__ Dp, const T* __restrict__ z,
                          const float* __restrict__ dbias, int softplus, T* __restrict__ out,
                          float* __restrict__ last_state) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int d = blockIdx.x * SS_WARPS + (threadIdx.x >> 5);
    const int b = blockIdx.y;
    if (d >= dim) return;
    const int grp = d / (dim / groups);
    const int64_t row = ((int64_t)b * dim + d) * L;
    const T* ur = u + row;
    const T* dr = delta + row;
    const T* zr = z ? z + row : nullptr;
    const T* Br = Bm + ((int64_t)b * groups + grp) * N * L;
    const T* Cr = Cm + ((int64_t)b * groups + grp) * N * L;
    const bool vec = (L % 4 == 0);
    constexpr float LOG2E = 1.4426950408889634f;
    const float A2lane = lane < N ? A[(int64_t)d * N + lane] * LOG2E : 0.f;  // lane n holds A[d, n]
    float carry = 0.f;                                                        // lane n holds h_n carry
    const float bias = dbias ? dbias[d] : 0.f;
    const float Dd = Dp ? Dp[d] : 0.f;

    for (int64_t c0 = 0; c0 < L; c0 += 32 * SS_ITEMS) {
        const int64_t t0 = c0 + lane * SS_ITEMS;
        float uv[SS_ITEMS], dv[SS_ITEMS], y[SS_ITEMS];
        load_items(ur, L, t0, vec, uv);
        load_items(dr, L, t0, vec, dv);
#pragma unroll
        for (int k = 0; k < SS_ITEMS; ++k) {
            float x = dv[k] + bias;
            if (softplus) x = softplus20(x);
            dv[k] = (t0 + k < L) ? x : 0.f;  // delta = 0 past the end: a = 1, b = 0 (identity map)
            y[k] = 0.f;
        }
        // two states in flight: their shuffle scans are independent chains (latency-bound at small batch x dim)
        const bool pre4 = PRE && vec && t0 + SS_ITEMS <= L;   // this lane's four positions are inside the row: raw prefetch
        Raw4<T> bq[PRE ? SS_NB : 1], cq[PRE ? SS_NB : 1];
        for (int nb = 0; nb < N; nb += SS_NB) {
            if (PRE && pre4) {
#pragma unroll
                for (int j = 0; j < (PRE ? SS_NB : 1); ++j)
                    if (nb + j < N) {
                        bq[j].ld(Br + (int64_t)(nb + j) * L + t0);
                        cq[j].ld(Cr + (int64_t)(nb + j) * L + t0);
                    }
            }
#pragma unroll(PRE ? SS_NB : 2)
          for (int jn = 0; jn < SS_NB; ++jn) {
            const int n = nb + jn;
            if (n >= N) break;
            const float A2 = __shfl_sync(0xffffffffu, A2lane, n);
            const float hin = __shfl_sync(0xffffffffu, carry, n);
            float bv[SS_ITEMS], cv[SS_ITEMS], a[SS_ITEMS], sloc[SS_ITEMS];
            if (PRE && pre4) {
                bq[PRE ? jn : 0].get(bv);
                cq[PRE ? jn : 0].get(cv);
            } else {
                load_items(Br + (int64_t)n * L, L, t0, vec, bv);
                load_items(Cr + (int64_t)n * L, L, t0, vec, cv);
            }
            float P = 1.f, S = 0.f;
#pragma unroll
            for (int k = 0; k < SS_ITEMS; ++k) {
                a[k] = ex2f_(dv[k] * A2);
                S = fmaf(a[k], S, dv[k] * bv[k] * uv[k]);
                P *= a[k];
                sloc[k] = S;
                a[k] = P;  // cumulative product up to k
            }
            // inclusive warp scan of the affine maps (P, S): later o earlier
            float Pi = P, Si = S;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                float Pp = __shfl_up_sync(0xffffffffu, Pi, o);
                float Sp = __shfl_up_sync(0xffffffffu, Si, o);
                if (lane >= o) {
                    Si = fmaf(Pi, Sp, Si);
                    Pi *= Pp;
                }
            }
            // exclusive prefix for this lane, applied to the chunk carry
            float Pe = __shfl_up_sync(0xffffffffu, Pi, 1);
            float Se = __shfl_up_sync(0xffffffffu, Si, 1);
            const float h0 = lane == 0 ? hin : fmaf(Pe, hin, Se);
#pragma unroll
            for (int k = 0; k < SS_ITEMS; ++k) y[k] = fmaf(fmaf(a[k], h0, sloc[k]), cv[k], y[k]);
            // new carry = state after the last lane
            const float hend = fmaf(Pi, hin, Si);
            const float hlast = __shfl_sync(0xffffffffu, hend, 31);
            if (lane == n) carry = hlast;
          }
        }
        float zv[SS_ITEMS];
        if (zr) load_items(zr, L, t0, vec, zv);
#pragma unroll
        for (int k = 0; k < SS_ITEMS; ++k) {
            float o = fmaf(Dd, uv[k], y[k]);
            if (zr) o *= silu_exact(zv[k]);
            y[k] = o;
        }
        if (vec && t0 + SS_ITEMS <= L) {
            st4(out + row + t0, make_float4(y[0], y[1], y[2], y[3]));
        } else {
#pragma unroll
            for (int k = 0; k < SS_ITEMS; ++k)
                if (t0 + k < L) st1(out + row + t0 + k, y[k]);
        }
    }
    if (last_state && lane < N) last_state[((int64_t)b * dim + d) * N + lane] = carry;
}


// ---- backward ------------------------------------------------------------------------------------------------------
// Replaces selective_scan_cuda.bwd (selective_scan.cpp:338-492, selective_scan_bwd_kernel.cuh:75-489) behind
// SelectiveScanFn.backward (selective_scan_interface.py:59-102).  Same mapping as the forward: one warp per (b, d)
// row, 128-step chunks of 32 lanes x 4 items.
//   pass 1 (only when L > 128): forward sweep that records the state entering every chunk in `ws`
//          (batch*dim, nchunks, N) fp32 -- the role of the reference's `x` checkpoint tensor, but written here, by
//          the backward, so the inference forward never pays for it;
//   pass 2: chunks in reverse.  Per state n the lane recomputes h over its 4 steps (forward warp scan of the affine
//          maps, as in the forward), then runs the adjoint recurrence G_t = a_t (C_t dy_t + G_{t+1}) as a REVERSE
//          warp scan (shfl_down) of the maps (a_t, a_t C_t dy_t); G_{t+1} entering the chunk from the right is the
//          carry of state n in lane n.  With g_t = C_t dy_t + G_{t+1} (the gradient reaching h_t):
//              d delta_t += g_t (h_{t-1} a_t A_n + B_t u_t)     dA_n += g_t h_{t-1} a_t delta_t
//              dB_t[n]   += g_t delta_t u_t                     dC_t[n] += dy_t h_t
//              du_t      += g_t delta_t B_t  (+ dy_t D)         dD += dy_t u_t
//          and, with z: dy_t = dout_t silu(z_t), dz_t = dout_t y_t silu'(z_t) with y recomputed here (the reference
//          saves `out` for this, selective_scan_interface.py:58).
// dB / dC are summed over the channels of a group with fp32 vector atomics into caller-zeroed (batch, G, N, L)
// buffers (the reference does the same, bwd_kernel.cuh:438-462); dA, dD, d(delta_bias) by one atomic per row.
template <typename T, bool PRE>
__global__ void __launch_bounds__(SS_WARPS * 32)
selective_scan_bwd_kernel(int batch, int dim, int64_t L, int N, int groups, const T* __restrict__ u,
                          const T* __restrict__ delta, const float* __restrict__ A, const T* __restrict__ Bm,
                          const T* __restrict__ Cm, const float* __restrict__ Dp, const T* __restrict__ z,
                          const float* __restrict__ dbias, int softplus, const T* __restrict__ dout,
                          T* __restrict__ du, T* __restrict__ ddelta, float* __restrict__ dA, float* __restrict__ dB,
                          float* __restrict__ dC, float* __restrict__ dD, T* __restrict__ dz,
                          float* __restrict__ ddbias, float* __restrict__ ws) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    constexpr int CH = 32 * SS_ITEMS;
    const int lane = threadIdx.x & 31;
    const int d = blockIdx.x * SS_WARPS + (threadIdx.x >> 5);
    const int b = blockIdx.y;
    if (d >= dim) return;
    const int grp = d / (dim / groups);
    const int64_t rowi = (int64_t)b * dim + d, row = rowi * L;
    const T* ur = u + row;
    const T* dr = delta + row;
    const T* zr = z ? z + row : nullptr;
    const T* gr = dout + row;
    const int64_t bc0 = ((int64_t)b * groups + grp) * N * L;
    const T* Br = Bm + bc0;
    const T* Cr = Cm + bc0;
    const bool vec = (L % 4 == 0);
    constexpr float LOG2E = 1.4426950408889634f;
    const float Alane = lane < N ? A[(int64_t)d * N + lane] : 0.f;
    const float A2lane = Alane * LOG2E;
    const float bias = dbias ? dbias[d] : 0.f;
    const float Dd = Dp ? Dp[d] : 0.f;
    const int nchunks = (int)((L + CH - 1) / CH);
    float* wsr = ws ? ws + rowi * nchunks * N : nullptr;

    // ---- pass 1: state entering each chunk
    if (nchunks > 1) {
        float carry = 0.f;
        for (int c = 0; c < nchunks - 1; ++c) {
            if (lane < N) wsr[(int64_t)c * N + lane] = carry;
            const int64_t t0 = (int64_t)c * CH + lane * SS_ITEMS;
            float uv[SS_ITEMS], dv[SS_ITEMS];
            load_items(ur, L, t0, vec, uv);
            load_items(dr, L, t0, vec, dv);
#pragma unroll
            for (int k = 0; k < SS_ITEMS; ++k) {
                float x = dv[k] + bias;
                if (softplus) x = softplus20(x);
                dv[k] = x;  // every chunk but the last is full
            }
            for (int n = 0; n < N; ++n) {
                const float A2 = __shfl_sync(0xffffffffu, A2lane, n);
                const float hin = __shfl_sync(0xffffffffu, carry, n);
                float bv[SS_ITEMS];
                load_items(Br + (int64_t)n * L, L, t0, vec, bv);
                float P = 1.f, S = 0.f;
#pragma unroll
                for (int k = 0; k < SS_ITEMS; ++k) {
                    const float a = ex2f_(dv[k] * A2);
                    S = fmaf(a, S, dv[k] * bv[k] * uv[k]);
                    P *= a;
                }
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float Pp = __shfl_up_sync(0xffffffffu, P, o);
                    const float Sp = __shfl_up_sync(0xffffffffu, S, o);
                    if (lane >= o) {
                        S = fmaf(P, Sp, S);
                        P *= Pp;
                    }
                }
                const float hlast = __shfl_sync(0xffffffffu, fmaf(P, hin, S), 31);
                if (lane == n) carry = hlast;
            }
        }
        if (lane < N) wsr[(int64_t)(nchunks - 1) * N + lane] = carry;
        __syncwarp();
    }

    // ---- pass 2: chunks in reverse
    float Gcarry = 0.f, dAacc = 0.f, dDacc = 0.f, dbacc = 0.f;
    for (int c = nchunks - 1; c >= 0; --c) {
        const int64_t t0 = (int64_t)c * CH + lane * SS_ITEMS;
        const bool full4 = vec && t0 + SS_ITEMS <= L;
        float uv[SS_ITEMS], dv[SS_ITEMS], raw[SS_ITEMS], gv[SS_ITEMS], zv[SS_ITEMS], dy[SS_ITEMS];
        float ddl[SS_ITEMS], dul[SS_ITEMS], ys[SS_ITEMS];
        load_items(ur, L, t0, vec, uv);
        load_items(dr, L, t0, vec, raw);
        load_items(gr, L, t0, vec, gv);
        if (zr) load_items(zr, L, t0, vec, zv);
#pragma unroll
        for (int k = 0; k < SS_ITEMS; ++k) {
            raw[k] += bias;
            const float x = softplus ? softplus20(raw[k]) : raw[k];
            dv[k] = (t0 + k < L) ? x : 0.f;
            dy[k] = zr ? gv[k] * silu_exact(zv[k]) : gv[k];
            ddl[k] = 0.f;
            dul[k] = dy[k] * Dd;
            ys[k] = 0.f;
        }
        const float cin = (nchunks > 1 && lane < N) ? wsr[(int64_t)c * N + lane] : 0.f;
        Raw4<T> bq[PRE ? SS_NB : 1], cq[PRE ? SS_NB : 1];
        for (int nb = 0; nb < N; nb += SS_NB) {
            if (PRE && full4) {
#pragma unroll
                for (int j = 0; j < (PRE ? SS_NB : 1); ++j)
                    if (nb + j < N) {
                        bq[j].ld(Br + (int64_t)(nb + j) * L + t0);
                        cq[j].ld(Cr + (int64_t)(nb + j) * L + t0);
                    }
            }
#pragma unroll(PRE ? SS_NB : 2)
          for (int jn = 0; jn < SS_NB; ++jn) {
            const int n = nb + jn;
            if (n >= N) break;
            const float A2 = __shfl_sync(0xffffffffu, A2lane, n);
            const float An = __shfl_sync(0xffffffffu, Alane, n);
            const float hin = __shfl_sync(0xffffffffu, cin, n);
            const float Gin = __shfl_sync(0xffffffffu, Gcarry, n);
            float bv[SS_ITEMS], cv[SS_ITEMS], a[SS_ITEMS], ck[SS_ITEMS], hm1[SS_ITEMS];
            if (PRE && full4) {
                bq[PRE ? jn : 0].get(bv);
                cq[PRE ? jn : 0].get(cv);
            } else {
                load_items(Br + (int64_t)n * L, L, t0, vec, bv);
                load_items(Cr + (int64_t)n * L, L, t0, vec, cv);
            }
            // forward maps of the lane's steps, and the adjoint maps (reverse order)
            float P = 1.f, S = 0.f;
#pragma unroll
            for (int k = 0; k < SS_ITEMS; ++k) {
                a[k] = ex2f_(dv[k] * A2);
                S = fmaf(a[k], S, dv[k] * bv[k] * uv[k]);
                P *= a[k];
                ck[k] = cv[k] * dy[k];
            }
            float Sr = 0.f;
#pragma unroll
            for (int k = SS_ITEMS - 1; k >= 0; --k) Sr = a[k] * (ck[k] + Sr);
            float Pf = P, Sf = S, Pb = P, Sb = Sr;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float Pp = __shfl_up_sync(0xffffffffu, Pf, o);
                const float Sp = __shfl_up_sync(0xffffffffu, Sf, o);
                const float Pn = __shfl_down_sync(0xffffffffu, Pb, o);
                const float Sn = __shfl_down_sync(0xffffffffu, Sb, o);
                if (lane >= o) {
                    Sf = fmaf(Pf, Sp, Sf);
                    Pf *= Pp;
                }
                if (lane + o < 32) {
                    Sb = fmaf(Pb, Sn, Sb);
                    Pb *= Pn;
                }
            }
            const float Pe = __shfl_up_sync(0xffffffffu, Pf, 1), Se = __shfl_up_sync(0xffffffffu, Sf, 1);
            float h = lane == 0 ? hin : fmaf(Pe, hin, Se);            // state entering the lane's first step
            const float Pq = __shfl_down_sync(0xffffffffu, Pb, 1), Sq = __shfl_down_sync(0xffffffffu, Sb, 1);
            float G = lane == 31 ? Gin : fmaf(Pq, Gin, Sq);          // adjoint entering the lane's last step from the right
            const float Gnew = __shfl_sync(0xffffffffu, fmaf(Pb, Gin, Sb), 0);
            if (lane == n) Gcarry = Gnew;
            float hk[SS_ITEMS];
#pragma unroll
            for (int k = 0; k < SS_ITEMS; ++k) {
                hm1[k] = h;
                h = fmaf(a[k], h, dv[k] * bv[k] * uv[k]);
                hk[k] = h;
                ys[k] = fmaf(cv[k], h, ys[k]);
            }
            float dAp = 0.f, dBk[SS_ITEMS], dCk[SS_ITEMS];
#pragma unroll
            for (int k = SS_ITEMS - 1; k >= 0; --k) {
                const float g = ck[k] + G;
                const float tmp = g * hm1[k] * a[k];
                ddl[k] = fmaf(tmp, An, fmaf(g * bv[k], uv[k], ddl[k]));
                dAp = fmaf(tmp, dv[k], dAp);
                dBk[k] = g * dv[k] * uv[k];
                dul[k] = fmaf(g * dv[k], bv[k], dul[k]);
                dCk[k] = dy[k] * hk[k];
                G = a[k] * g;
            }
            dAp = warp_sum(dAp);
            if (lane == n) dAacc += dAp;
            float* dBr = dB + bc0 + (int64_t)n * L + t0;
            float* dCr = dC + bc0 + (int64_t)n * L + t0;
            if (full4) {
                atomicAdd(reinterpret_cast<float4*>(dBr), make_float4(dBk[0], dBk[1], dBk[2], dBk[3]));
                atomicAdd(reinterpret_cast<float4*>(dCr), make_float4(dCk[0], dCk[1], dCk[2], dCk[3]));
            } else {
#pragma unroll
                for (int k = 0; k < SS_ITEMS; ++k)
                    if (t0 + k < L) {
                        atomicAdd(dBr + k, dBk[k]);
                        atomicAdd(dCr + k, dCk[k]);
                    }
            }
          }
        }
        float dzv[SS_ITEMS];
#pragma unroll
        for (int k = 0; k < SS_ITEMS; ++k) {
            const bool in = t0 + k < L;
            if (zr) dzv[k] = gv[k] * fmaf(Dd, uv[k], ys[k]) * dsilu(zv[k]);
            if (in) dDacc = fmaf(dy[k], uv[k], dDacc);
            if (softplus && raw[k] <= 20.f) ddl[k] *= sigmoidf_(raw[k]);
            if (in) dbacc += ddl[k];
        }
        if (full4) {
            st4(du + row + t0, make_float4(dul[0], dul[1], dul[2], dul[3]));
            st4(ddelta + row + t0, make_float4(ddl[0], ddl[1], ddl[2], ddl[3]));
            if (zr) st4(dz + row + t0, make_float4(dzv[0], dzv[1], dzv[2], dzv[3]));
        } else {
#pragma unroll
            for (int k = 0; k < SS_ITEMS; ++k)
                if (t0 + k < L) {
                    st1(du + row + t0 + k, dul[k]);
                    st1(ddelta + row + t0 + k, ddl[k]);
                    if (zr) st1(dz + row + t0 + k, dzv[k]);
                }
        }
    }
    if (lane < N) atomicAdd(dA + (int64_t)d * N + lane, dAacc);
    dDacc = warp_sum(dDacc);
    dbacc = warp_sum(dbacc);
    if (lane == 0) {
        if (dD) atomicAdd(dD + d, dDacc);
        if (ddbias) atomicAdd(ddbias + d, dbacc);
    }
}

}  // namespace fv

// short sequences (L <= 16, d_state 16): selective_scan_short.cu
namespace fv {
int ss_short_fwd_rows(int dim, int64_t L, int dstate, int groups);
bool ss_short_bwd_ok(int dim, int64_t L, int dstate, int groups);
template <typename T>
int launch_ss_short_fwd(int rows, int batch, int dim, int L, int groups, const T* u, const T* delta, const float* A, const T* B,
                        const T* C, const float* D, const T* z, const float* dbias, int softplus, T* out, float* last,
                        cudaStream_t st);
template <typename T>
int launch_ss_short_bwd(int batch, int dim, int L, int groups, const T* u, const T* delta, const float* A, const T* B, const T* C,
                        const float* D, const T* z, const float* dbias, int softplus, const T* dout, T* du, T* ddelta,
                        float* dA, float* dB, float* dC, float* dD, T* dz, float* ddbias, cudaStream_t st);
}  // namespace fv

extern "C" int fv_selective_scan_fwd(int dtype, int batch, int dim, int64_t L, int dstate, int groups,
                                     const void* u, const void* delta, const float* A, const void* B,
                                     const void* C, const float* D, const void* z,
                                     const float* delta_bias, int delta_softplus, void* out,
                                     float* last_state, void* stream) {
    using namespace fv;
    FV_REQUIRE(u && delta && A && B && C && out, "fv_selective_scan_fwd: null pointer");
    FV_REQUIRE(batch > 0 && dim > 0 && L > 0, "fv_selective_scan_fwd: bad shape");
    FV_REQUIRE(dstate >= 1 && dstate <= 32, "fv_selective_scan_fwd: d_state %d not in [1, 32]", dstate);
    FV_REQUIRE(groups >= 1 && dim % groups == 0, "fv_selective_scan_fwd: dim %d not divisible by groups %d", dim, groups);
    FV_REQUIRE(batch <= 65535, "fv_selective_scan_fwd: batch > 65535");
    cudaStream_t st = (cudaStream_t)stream;
    if (const int rows = ss_short_fwd_rows(dim, L, dstate, groups)) {   // thread-per-row kernel for the pooled FastVim lengths
        if (dtype == FV_F32)
            return launch_ss_short_fwd<float>(rows, batch, dim, (int)L, groups, (const float*)u, (const float*)delta, A, (const float*)B,
                                              (const float*)C, D, (const float*)z, delta_bias, delta_softplus, (float*)out, last_state, st);
        if (dtype == FV_BF16)
            return launch_ss_short_fwd<bf16>(rows, batch, dim, (int)L, groups, (const bf16*)u, (const bf16*)delta, A, (const bf16*)B,
                                             (const bf16*)C, D, (const bf16*)z, delta_bias, delta_softplus, (bf16*)out, last_state, st);
        return fail("fv_selective_scan_fwd: unsupported dtype %d", dtype);
    }
    dim3 grid(ceil_div(dim, SS_WARPS), batch), block(SS_WARPS * 32);
    // few rows (latency-bound: 1 x 384 x 128 runs 8.2 -> 6.8 us forward, 24.6 -> 20.9 us backward with the raw B / C prefetch);
    // with many rows the kernel is throughput-bound and the extra registers cost occupancy (32 x 768 x 112: 81 -> 101 us)
    const bool pre = (int64_t)batch * dim <= (int64_t)sm_count() * 16;
    if (dtype == FV_F32)
        { decltype(&selective_scan_fwd_kernel<float, true>) kern = pre ? &selective_scan_fwd_kernel<float, true> : &selective_scan_fwd_kernel<float, false>;
        FV_LAUNCH_PDL((kern), grid, block, 0, st, batch, dim, L, dstate, groups, (const float*)u, (const float*)delta, A, (const float*)B, (const float*)C, D, (const float*)z, delta_bias, delta_softplus, (float*)out, last_state); }
    else if (dtype == FV_BF16)
        { decltype(&selective_scan_fwd_kernel<bf16, true>) kern = pre ? &selective_scan_fwd_kernel<bf16, true> : &selective_scan_fwd_kernel<bf16, false>;
        FV_LAUNCH_PDL((kern), grid, block, 0, st, batch, dim, L, dstate, groups, (const bf16*)u, (const bf16*)delta, A, (const bf16*)B, (const bf16*)C, D, (const bf16*)z, delta_bias, delta_softplus, (bf16*)out, last_state); }
    else
        return fail("fv_selective_scan_fwd: unsupported dtype %d", dtype);
    return finish_launch("selective_scan_fwd");
}

extern "C" int64_t fv_selective_scan_bwd_workspace_bytes(int batch, int dim, int64_t L, int dstate) {
    if (batch <= 0 || dim <= 0 || L <= 0 || dstate <= 0) return 0;
    const int64_t nchunks = (L + 32 * fv::SS_ITEMS - 1) / (32 * fv::SS_ITEMS);
    return nchunks > 1 ? (int64_t)batch * dim * nchunks * dstate * 4 : 0;
}

extern "C" int fv_selective_scan_bwd(int dtype, int batch, int dim, int64_t L, int dstate, int groups,
                                     const void* u, const void* delta, const float* A, const void* B,
                                     const void* C, const float* D, const void* z, const float* delta_bias,
                                     int delta_softplus, const void* dout, void* du, void* ddelta, float* dA,
                                     float* dB, float* dC, float* dD, void* dz, float* ddelta_bias,
                                     void* workspace, int64_t workspace_bytes, void* stream) {
    using namespace fv;
    FV_REQUIRE(u && delta && A && B && C && dout && du && ddelta && dA && dB && dC, "fv_selective_scan_bwd: null pointer");
    FV_REQUIRE(batch > 0 && dim > 0 && L > 0, "fv_selective_scan_bwd: bad shape");
    FV_REQUIRE(dstate >= 1 && dstate <= 32, "fv_selective_scan_bwd: d_state %d not in [1, 32]", dstate);
    FV_REQUIRE(groups >= 1 && dim % groups == 0, "fv_selective_scan_bwd: dim %d not divisible by groups %d", dim, groups);
    FV_REQUIRE(batch <= 65535, "fv_selective_scan_bwd: batch > 65535");
    FV_REQUIRE((!D || dD) && (!z || dz) && (!delta_bias || ddelta_bias), "fv_selective_scan_bwd: missing gradient buffer");
    const int64_t need = fv_selective_scan_bwd_workspace_bytes(batch, dim, L, dstate);
    FV_REQUIRE(need == 0 || (workspace && workspace_bytes >= need), "fv_selective_scan_bwd: workspace of %lld bytes required",
               (long long)need);
    FV_REQUIRE(L % 4 != 0 || (((uintptr_t)dB | (uintptr_t)dC) % 16) == 0, "fv_selective_scan_bwd: dB / dC must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (ss_short_bwd_ok(dim, L, dstate, groups)) {
        if (dtype == FV_F32)
            return launch_ss_short_bwd<float>(batch, dim, (int)L, groups, (const float*)u, (const float*)delta, A, (const float*)B,
                                              (const float*)C, D, (const float*)z, delta_bias, delta_softplus, (const float*)dout,
                                              (float*)du, (float*)ddelta, dA, dB, dC, dD, (float*)dz, ddelta_bias, st);
        if (dtype == FV_BF16)
            return launch_ss_short_bwd<bf16>(batch, dim, (int)L, groups, (const bf16*)u, (const bf16*)delta, A, (const bf16*)B,
                                             (const bf16*)C, D, (const bf16*)z, delta_bias, delta_softplus, (const bf16*)dout,
                                             (bf16*)du, (bf16*)ddelta, dA, dB, dC, dD, (bf16*)dz, ddelta_bias, st);
        return fail("fv_selective_scan_bwd: unsupported dtype %d", dtype);
    }
    dim3 grid(ceil_div(dim, SS_WARPS), batch), block(SS_WARPS * 32);
    // few rows (latency-bound: 1 x 384 x 128 runs 8.2 -> 6.8 us forward, 24.6 -> 20.9 us backward with the raw B / C prefetch);
    // with many rows the kernel is throughput-bound and the extra registers cost occupancy (32 x 768 x 112: 81 -> 101 us)
    const bool pre = (int64_t)batch * dim <= (int64_t)sm_count() * 16;
    if (dtype == FV_F32)
        { decltype(&selective_scan_bwd_kernel<float, true>) kern = pre ? &selective_scan_bwd_kernel<float, true> : &selective_scan_bwd_kernel<float, false>;
        FV_LAUNCH_PDL((kern), grid, block, 0, st, batch, dim, L, dstate, groups, (const float*)u, (const float*)delta, A, (const float*)B, (const float*)C, D, (const float*)z, delta_bias, delta_softplus, (const float*)dout, (float*)du, (float*)ddelta, dA, dB, dC, dD, (float*)dz, ddelta_bias, (float*)workspace); }
    else if (dtype == FV_BF16)
        { decltype(&selective_scan_bwd_kernel<bf16, true>) kern = pre ? &selective_scan_bwd_kernel<bf16, true> : &selective_scan_bwd_kernel<bf16, false>;
        FV_LAUNCH_PDL((kern), grid, block, 0, st, batch, dim, L, dstate, groups, (const bf16*)u, (const bf16*)delta, A, (const bf16*)B, (const bf16*)C, D, (const bf16*)z, delta_bias, delta_softplus, (const bf16*)dout, (bf16*)du, (bf16*)ddelta, dA, dB, dC, dD, (bf16*)dz, ddelta_bias, (float*)workspace); }
    else
        return fail("fv_selective_scan_bwd: unsupported dtype %d", dtype);
    return finish_launch("selective_scan_bwd");
}
