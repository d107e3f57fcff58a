// K2b-bwd, streaming form -- backward of broadcast + D skip + direction average + LayerNorm + SiLU gate
// (same contract as csrc/gate_bwd.cu; reference: mamba_ssm/modules/mamba_simple_faster.py:356-358, 412-416, 434-453).
//
// gate_bwd.cu keeps whole token rows (all channels) of a <= 8-token tile in shared memory because LayerNorm couples the
// channels: one 384-thread CTA per SM at dim 1536, three block-wide phases per tile, 386 us per launch at FastVim-B
// (5.5x its HBM time; shorter tiles are slower, profiles/r01_kbench_gate_bwd_tile_sweep.log).  The coupling, however,
// is only four numbers per token:
//     S1 = sum_d v      S2 = sum_d v^2      S3 = sum_d dxh      S4 = sum_d dxh v        (dxh = dy silu(z) gamma)
// from which mean = S1/D, rstd = rsqrt(S2/D - mean^2 + eps), c1 = S3/D and c2 = mean(dxh xhat) = rstd (S4 - mean S3)/D
// follow -- and dxh does not depend on the statistics.  So the work splits into two channel-local passes that both
// STREAM (a thread owns two channels and walks a run of tokens with a 7-token register window of x, no shared-memory
// data, no block barrier in the walk; see csrc/conv_pool_bwd.cu):
//   pass 0 (statistics): recompute v, form dxh, reduce the four partial sums over the warp with a 32-value transposed
//          butterfly (31 shuffles per 8 tokens) and add them to stats (B, L, 4) with one atomic per lane;
//   pass 1 (apply):      recompute v, read the token's four sums, emit dz and e = dv/2, accumulate ds (one plane,
//          flushed per pooled row), dD_f, dD_b, dgamma, dbeta in registers (one atomic per CTA and channel at the end).
// x, z, dy are read twice (8 T of traffic instead of 5 T) and the convs are evaluated twice; both passes run at the
// streaming kernels' occupancy instead of one CTA per SM.
#include <cstdlib>

#include "common.cuh"
#include "stream.cuh"

namespace fv {

int sm_count();
int check_geom(const fv_geom* g, const char* who);

constexpr int GBS_MAXRUN = 56;

template <bool FAST>
__device__ __forceinline__ float2 silu2_sel(float2 x) {
    if (FAST) {   // x * sigmoid(x) = h + h tanh(h), h = x / 2
        const float2 h = __fmul2_rn(x, make_float2(0.5f, 0.5f));
        float tx, ty;
        asm("tanh.approx.f32 %0, %1;" : "=f"(tx) : "f"(h.x));
        asm("tanh.approx.f32 %0, %1;" : "=f"(ty) : "f"(h.y));
        return __ffma2_rn(h, make_float2(tx, ty), h);
    }
    return make_float2(silu_exact(x.x), silu_exact(x.y));
}
// silu(z) and d silu(z) / dz of a pair
template <bool FAST>
__device__ __forceinline__ void silu_grad2(float2 zv, float2& sl, float2& ds) {
    float2 sg;
    if (FAST) {
        float tx, ty;
        asm("tanh.approx.f32 %0, %1;" : "=f"(tx) : "f"(0.5f * zv.x));
        asm("tanh.approx.f32 %0, %1;" : "=f"(ty) : "f"(0.5f * zv.y));
        sg = __ffma2_rn(make_float2(tx, ty), make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));
    } else {
        sg = make_float2(sigmoidf_(zv.x), sigmoidf_(zv.y));
    }
    sl = __fmul2_rn(zv, sg);
    const float2 oms = __ffma2_rn(sg, make_float2(-1.f, -1.f), make_float2(1.f, 1.f));
    ds = __fmul2_rn(sg, __ffma2_rn(zv, oms, make_float2(1.f, 1.f)));
}

// sum over the warp of 32 per-lane values: afterwards lane l holds the warp total of v[l]  (31 shuffles)
__device__ __forceinline__ float warp_transpose_reduce32(float (&v)[32], int lane) {
#define FV_TR_STAGE(OFF_, N_)                                                       \
    {                                                                               \
        const bool up = (lane & (OFF_)) != 0;                                       \
        _Pragma("unroll") for (int i = 0; i < (N_) / 2; ++i) {                      \
            const float send = up ? v[i] : v[i + (N_) / 2];                         \
            const float keep = up ? v[i + (N_) / 2] : v[i];                         \
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, (OFF_));               \
        }                                                                           \
    }
    FV_TR_STAGE(16, 32) FV_TR_STAGE(8, 16) FV_TR_STAGE(4, 8) FV_TR_STAGE(2, 4) FV_TR_STAGE(1, 2)
#undef FV_TR_STAGE
    return v[0];
}

template <typename T, bool NORM, int MODE>
__global__ void __launch_bounds__(256, 2)
gate_bwd_stream_kernel(Geom g, int nseg, int seg_len, int64_t nitems, int nslots, int chunks,
                       const T* __restrict__ x, const T* __restrict__ z, int64_t ldxz, int64_t xzbs,
                       const T* __restrict__ dy, int64_t lddy, int64_t dybs, const float* __restrict__ s,
                       const float* __restrict__ cw, const float* __restrict__ cb, const float* __restrict__ Dskip,
                       const float* __restrict__ lnw, const float* __restrict__ lnb, float eps,
                       float* __restrict__ stats, T* __restrict__ dz, T* __restrict__ e_out, float* __restrict__ ds,
                       float* __restrict__ dDskip, float* __restrict__ dlnw, float* __restrict__ dlnb) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    typedef Pair<T> P;
    typedef typename P::type PT;
    constexpr bool FAST = is_fast<T>::value;
    // per-run tables (padded so the unrolled walk never indexes past them); index i <-> token t0 - 3 + i for xoff,
    // index q <-> owned token t0 + q for the others.  -1: outside the image / the run.
    __shared__ int xoff[GBS_MAXRUN + 24], yoff[GBS_MAXRUN + 24], eoff[GBS_MAXRUN + 24], jtab[GBS_MAXRUN + 24];
    const int D = g.D, lane = threadIdx.x & 31;
    const int slot = blockIdx.x / chunks, chunk = blockIdx.x - slot * chunks;
    const int d0 = (chunk * blockDim.x + threadIdx.x) * 2;
    const int64_t splane = (int64_t)g.B * g.Lp * D;
    const float invD = 1.f / (float)D;
    const float2 z2 = make_float2(0.f, 0.f), half2 = make_float2(0.5f, 0.5f);
    float2 wf[4], wb[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        wf[k] = make_float2(cw[(int64_t)d0 * 4 + k], cw[(int64_t)(d0 + 1) * 4 + k]);
        wb[k] = make_float2(cw[((int64_t)D + d0) * 4 + k], cw[((int64_t)D + d0 + 1) * 4 + k]);
    }
    const float2 bf_ = cb ? make_float2(cb[d0], cb[d0 + 1]) : z2;
    const float2 bb_ = cb ? make_float2(cb[D + d0], cb[D + d0 + 1]) : z2;
    const float2 Dfh = make_float2(0.5f * Dskip[d0], 0.5f * Dskip[d0 + 1]);
    const float2 Dbh = make_float2(0.5f * Dskip[D + d0], 0.5f * Dskip[D + d0 + 1]);
    const float2 gam = NORM ? make_float2(lnw[d0], lnw[d0 + 1]) : make_float2(1.f, 1.f);
    const float2 bet = (NORM && lnb) ? make_float2(lnb[d0], lnb[d0 + 1]) : z2;
    float2 acc_dDf = z2, acc_dDb = z2, acc_dg = z2, acc_db = z2;

    for (int64_t item = slot; item < nitems; item += nslots) {
        const int b = (int)(item / nseg), sgi = (int)(item - (int64_t)b * nseg);
        const int t0 = sgi * seg_len, n = min(seg_len, g.L - t0), n6 = n + 6;
        __syncthreads();  // the previous run's tables are no longer read
        for (int i = threadIdx.x; i < n6 + 16; i += blockDim.x) {
            const int t = t0 - 3 + i;
            const bool in = i < n6 && t >= 0 && t < g.L;
            xoff[i] = in ? (int)seq_to_row(g, t) * (int)ldxz : -1;
            const int tq = t0 + i;                       // owned token q = i
            const bool own = i < n;
            const int row = own ? (int)seq_to_row(g, tq) : 0;
            yoff[i] = own ? row * (int)lddy : -1;
            eoff[i] = own ? row * D : -1;
            jtab[i] = min(tq, g.L - 1) / g.pool;
        }
        __syncthreads();
        const T* xb = x + (int64_t)b * xzbs + d0;
        const T* zb = z + (int64_t)b * xzbs + d0;
        const T* dyb = dy + (int64_t)b * dybs + d0;
        const float* sb_ = s + (int64_t)b * g.Lp * D + d0;
        const float* stb = stats + (int64_t)b * g.L * 4 + (int64_t)t0 * 4;
        PT qx[4], qz[4], qd[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int ox = xoff[k], oz = xoff[k + 3], oy = yoff[k];   // x of walk index k; z / dy of owned token k
            qx[k] = ox >= 0 ? P::ld(xb + ox) : P::zero();
            qz[k] = yoff[k] >= 0 ? P::ld(zb + oz) : P::zero();
            qd[k] = oy >= 0 ? P::ld(dyb + oy) : P::zero();
        }
        float2 ring[8] = {z2, z2, z2, z2, z2, z2, z2, z2};
        float2 sv = z2, ds_acc = z2;
        int jcur = -1;
        for (int base = 0; base < n6; base += 8) {
            float pv[32];
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 32; ++i) pv[i] = 0.f;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = base + k, q = i - 6;               // walk index (x row t0-3+i), owned token index
                ring[k] = P::up(qx[k & 3]);
                {
                    const int ox = xoff[i + 4];
                    qx[k & 3] = ox >= 0 ? P::ld(xb + ox) : P::zero();
                }
                const bool own = i >= 6 && i < n6;
                const int qc = own ? q : 0;
                const float2 zv = P::up(qz[(k + 2) & 3]);                                 // token q (slot q & 3)
                const float2 dyv = own ? P::up(qd[(k + 2) & 3]) : z2;                     // (slots hold later tokens while i < 6)
                if (i >= 2) {                                     // refill with owned token q + 4 = i - 2
                    const int oz = xoff[i + 1], oy = yoff[i - 2];
                    qz[(k + 2) & 3] = oy >= 0 ? P::ld(zb + oz) : P::zero();
                    qd[(k + 2) & 3] = oy >= 0 ? P::ld(dyb + oy) : P::zero();
                }
                const int j = jtab[qc];
                if (own && j != jcur) {
                    if (MODE == 1 && jcur >= 0) {                 // pooled row finished: flush its ds
                        float* dp = ds + ((int64_t)b * g.Lp + jcur) * D + d0;
                        atomicAdd(dp, ds_acc.x);
                        atomicAdd(dp + 1, ds_acc.y);
                        ds_acc = z2;
                    }
                    jcur = j;
                    const float2 a = __ldg(reinterpret_cast<const float2*>(sb_ + (int64_t)j * D));
                    const float2 c = __ldg(reinterpret_cast<const float2*>(sb_ + splane + (int64_t)j * D));
                    sv = __fmul2_rn(__fadd2_rn(a, c), half2);
                }
                // x[t-3 .. t+3] of token t = t0 + q:  x[t+m] sits in ring[(k + 5 + m) & 7]
                const float2 xm3 = ring[(k + 2) & 7], xm2 = ring[(k + 3) & 7], xm1 = ring[(k + 4) & 7], x00 = ring[(k + 5) & 7];
                const float2 xp1 = ring[(k + 6) & 7], xp2 = ring[(k + 7) & 7], xp3 = ring[k];
                const float2 cf = __ffma2_rn(wf[3], x00, __ffma2_rn(wf[2], xm1, __ffma2_rn(wf[1], xm2, __ffma2_rn(wf[0], xm3, bf_))));
                const float2 cbk = __ffma2_rn(wb[3], x00, __ffma2_rn(wb[2], xp1, __ffma2_rn(wb[1], xp2, __ffma2_rn(wb[0], xp3, bb_))));
                const float2 af = silu2_sel<FAST>(cf), ab = silu2_sel<FAST>(cbk);
                const float2 v = __ffma2_rn(Dbh, ab, __ffma2_rn(Dfh, af, sv));
                float2 zz, dsz;
                silu_grad2<FAST>(zv, zz, dsz);
                const float2 dln = __fmul2_rn(dyv, zz);            // 0 for tokens that are not owned (dy = 0)
                const float2 dxh = __fmul2_rn(dln, gam);
                if (MODE == 0) {
                    const float2 vo = own ? v : z2;
                    const float2 v2 = __fmul2_rn(vo, vo), dv_ = __fmul2_rn(dxh, vo);
                    pv[4 * k + 0] = vo.x + vo.y;
                    pv[4 * k + 1] = v2.x + v2.y;
                    pv[4 * k + 2] = dxh.x + dxh.y;
                    pv[4 * k + 3] = dv_.x + dv_.y;
                } else {
                    float2 xh = v, dv = dln, lnv = v;
                    if (NORM) {
                        float4 st = make_float4(0.f, 0.f, 1.f, 0.f);
                        if (own) st = __ldg(reinterpret_cast<const float4*>(stb + (int64_t)qc * 4));
                        const float mean = st.x * invD;
                        const float rstd = rsqrtf(fmaxf(fmaf(-mean, mean, st.y * invD), 0.f) + eps);
                        const float c1 = st.z * invD, c2 = rstd * (st.w - mean * st.z) * invD;
                        const float2 r2 = make_float2(rstd, rstd);
                        xh = __fmul2_rn(__fadd2_rn(v, make_float2(-mean, -mean)), r2);
                        lnv = __ffma2_rn(xh, gam, bet);
                        dv = __fmul2_rn(r2, __ffma2_rn(xh, make_float2(-c2, -c2), __fadd2_rn(dxh, make_float2(-c1, -c1))));
                        acc_dg = __ffma2_rn(dln, xh, acc_dg);     // dln = 0 when not owned
                        acc_db = __fadd2_rn(acc_db, dln);
                    }
                    const float2 e = own ? __fmul2_rn(dv, half2) : z2;
                    if (own) {
                        P::st(dz + (int64_t)b * xzbs + d0 + xoff[qc + 3], __fmul2_rn(__fmul2_rn(dyv, lnv), dsz));
                        P::st(e_out + (int64_t)b * g.L * D + d0 + eoff[qc], e);
                    }
                    ds_acc = __fadd2_rn(ds_acc, e);
                    acc_dDf = __ffma2_rn(e, af, acc_dDf);
                    acc_dDb = __ffma2_rn(e, ab, acc_dDb);
                }
            }
            if (MODE == 0) {
                // lane l ends up with the warp sum of component l & 3 of the block's token l >> 2 (walk index base + (l >> 2))
                const float tot = warp_transpose_reduce32(pv, lane);
                const int q = base + (lane >> 2) - 6;
                if (q >= 0 && q < n) atomicAdd(stats + ((int64_t)b * g.L + t0 + q) * 4 + (lane & 3), tot);
            }
        }
        if (MODE == 1 && jcur >= 0) {
            float* dp = ds + ((int64_t)b * g.Lp + jcur) * D + d0;
            atomicAdd(dp, ds_acc.x);
            atomicAdd(dp + 1, ds_acc.y);
        }
    }
    if (MODE == 1) {
        atomicAdd(dDskip + d0, acc_dDf.x); atomicAdd(dDskip + d0 + 1, acc_dDf.y);
        atomicAdd(dDskip + D + d0, acc_dDb.x); atomicAdd(dDskip + D + d0 + 1, acc_dDb.y);
        if (NORM) {
            atomicAdd(dlnw + d0, acc_dg.x); atomicAdd(dlnw + d0 + 1, acc_dg.y);
            atomicAdd(dlnb + d0, acc_db.x); atomicAdd(dlnb + d0 + 1, acc_db.y);
        }
    }
}

template <typename T, bool NORM, int MODE>
static int launch_gbs(const Geom& g, const T* x, const T* z, int64_t ldxz, int64_t xzbs, const T* dy, int64_t lddy,
                      int64_t dybs, const float* s, const float* cw, const float* cb, const float* Dskip, const float* lnw,
                      const float* lnb, float eps, float* stats, T* dz, T* e_out, float* ds, float* dDskip, float* dlnw,
                      float* dlnb, cudaStream_t st) {
    const int threads = stream_block(g.D), chunks = (g.D / 2) / threads;
    const int nseg = ceil_div(g.L, GBS_MAXRUN), seg_len = ceil_div(g.L, nseg);
    const int64_t nitems = (int64_t)g.B * nseg;
    auto kern = gate_bwd_stream_kernel<T, NORM, MODE>;
    int occ = 0;
    cudaError_t er = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, 0);
    FV_REQUIRE(er == cudaSuccess && occ > 0, "fv_gate_bwd_stream: occupancy query failed (%s)", cudaGetErrorString(er));
    int64_t nslots = ((int64_t)sm_count() * occ) / chunks;
    if (nslots < 1) nslots = 1;
    if (nslots > nitems) nslots = nitems;
    const int64_t rounds = (nitems + nslots - 1) / nslots;
    nslots = (nitems + rounds - 1) / rounds;
    FV_LAUNCH_PDL((kern), (unsigned)(nslots * chunks), threads, 0, st, g, nseg, seg_len, nitems, (int)nslots, chunks, x, z, ldxz, xzbs, dy, lddy,
                                                          dybs, s, cw, cb, Dskip, lnw, lnb, eps, stats, dz, e_out, ds, dDskip,
                                                          dlnw, dlnb);
    return finish_launch(MODE == 0 ? "gate_bwd_stats" : "gate_bwd_apply");
}

template <typename T>
static int run_gbs(const Geom& g, const T* x, const T* z, int64_t ldxz, int64_t xzbs, const T* dy, int64_t lddy, int64_t dybs,
                   const float* s, const float* cw, const float* cb, const float* Dskip, const float* lnw, const float* lnb,
                   float eps, float* stats, T* dz, T* e_out, float* ds, float* dDskip, float* dlnw, float* dlnb,
                   cudaStream_t st) {
    if (lnw) {
        if (int rc = launch_gbs<T, true, 0>(g, x, z, ldxz, xzbs, dy, lddy, dybs, s, cw, cb, Dskip, lnw, lnb, eps, stats, dz, e_out,
                                            ds, dDskip, dlnw, dlnb, st))
            return rc;
        return launch_gbs<T, true, 1>(g, x, z, ldxz, xzbs, dy, lddy, dybs, s, cw, cb, Dskip, lnw, lnb, eps, stats, dz, e_out, ds,
                                      dDskip, dlnw, dlnb, st);
    }
    return launch_gbs<T, false, 1>(g, x, z, ldxz, xzbs, dy, lddy, dybs, s, cw, cb, Dskip, lnw, lnb, eps, stats, dz, e_out, ds,
                                   dDskip, dlnw, dlnb, st);
}

}  // namespace fv

extern "C" int fv_gate_bwd_stream_supported(const fv_geom* g, int64_t ldxz, int64_t lddy) {
    if (!g || g->inner != 1 || g->dim <= 0 || g->pool <= 0) return 0;
    if (fv::stream_block(g->dim) == 0 || ldxz % 2 || lddy % 2) return 0;
    const int64_t L = (int64_t)g->outer * g->pool;
    return L * (ldxz > lddy ? ldxz : lddy) < (1ll << 31);   // 32-bit row-offset tables
}

extern "C" int fv_gate_bwd_stream(const fv_geom* g_, int dtype, const void* x, const void* z, int64_t ldxz,
                                  int64_t xz_bstride, const void* dy, int64_t lddy, int64_t dy_bstride, const float* s,
                                  const float* conv_w, const float* conv_b, const float* Dskip, const float* ln_w,
                                  const float* ln_b, float eps, float* stats, void* dz, void* e_out, float* ds,
                                  float* dDskip, float* dln_w, float* dln_b, void* stream) {
    using namespace fv;
    if (int rc = check_geom(g_, "fv_gate_bwd_stream")) return rc;
    FV_REQUIRE(x && z && dy && s && conv_w && Dskip && dz && e_out && ds && dDskip, "fv_gate_bwd_stream: null pointer");
    FV_REQUIRE(!ln_w || (dln_w && dln_b && stats), "fv_gate_bwd_stream: dln_w / dln_b / stats required with LayerNorm");
    FV_REQUIRE(fv_gate_bwd_stream_supported(g_, ldxz, lddy), "fv_gate_bwd_stream: unsupported configuration (plain geometry, "
               "dim %% 64 == 0, even strides)");
    FV_REQUIRE(xz_bstride % 2 == 0 && dy_bstride % 2 == 0 && g_->batch <= 65535, "fv_gate_bwd_stream: bad batch stride / batch");
    FV_REQUIRE(!stats || ((uintptr_t)stats % 16) == 0, "fv_gate_bwd_stream: stats must be 16-byte aligned");
    Geom g = make_geom(g_);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FV_F32)
        return run_gbs<float>(g, (const float*)x, (const float*)z, ldxz, xz_bstride, (const float*)dy, lddy, dy_bstride, s, conv_w,
                              conv_b, Dskip, ln_w, ln_b, eps, stats, (float*)dz, (float*)e_out, ds, dDskip, dln_w, dln_b, st);
    if (dtype == FV_BF16)
        return run_gbs<bf16>(g, (const bf16*)x, (const bf16*)z, ldxz, xz_bstride, (const bf16*)dy, lddy, dy_bstride, s, conv_w,
                             conv_b, Dskip, ln_w, ln_b, eps, stats, (bf16*)dz, (bf16*)e_out, ds, dDskip, dln_w, dln_b, st);
    return fail("fv_gate_bwd_stream: unsupported dtype %d", dtype);
}
