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

template <typename T>
__global__ void __launch_bounds__(SS_WARPS * 32)
selective_scan_fwd_kernel(int batch, int dim, int64_t L, int N, int groups, const T* __restrict__ u,
                          const T* __restrict__ delta, const float* __restrict__ A,
                          const T* __restrict__ Bm, const T* __restrict__ Cm,
                          const float* __restrict__ Dp, const T* __restrict__ z,
                          const float* __restrict__ dbias, int softplus, T* __restrict__ out,
                          float* __restrict__ last_state) {
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
        for (int n = 0; n < N; ++n) {
            const float A2 = __shfl_sync(0xffffffffu, A2lane, n);
            const float hin = __shfl_sync(0xffffffffu, carry, n);
            float bv[SS_ITEMS], cv[SS_ITEMS], a[SS_ITEMS], sloc[SS_ITEMS];
            load_items(Br + (int64_t)n * L, L, t0, vec, bv);
            load_items(Cr + (int64_t)n * L, L, t0, vec, cv);
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
    dim3 grid(ceil_div(dim, SS_WARPS), batch), block(SS_WARPS * 32);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FV_F32)
        selective_scan_fwd_kernel<float><<<grid, block, 0, st>>>(batch, dim, L, dstate, groups, (const float*)u, (const float*)delta, A, (const float*)B, (const float*)C, D, (const float*)z, delta_bias, delta_softplus, (float*)out, last_state);
    else if (dtype == FV_BF16)
        selective_scan_fwd_kernel<bf16><<<grid, block, 0, st>>>(batch, dim, L, dstate, groups, (const bf16*)u, (const bf16*)delta, A, (const bf16*)B, (const bf16*)C, D, (const bf16*)z, delta_bias, delta_softplus, (bf16*)out, last_state);
    else
        return fail("fv_selective_scan_fwd: unsupported dtype %d", dtype);
    return finish_launch("selective_scan_fwd");
}
