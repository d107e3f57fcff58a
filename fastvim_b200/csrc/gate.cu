// K2b -- scan epilogue: broadcast-back + D skip + direction average + LayerNorm(d_inner) + SiLU(z) gate.
//
// What it replaces in the reference (paths relative to /root/reference):
//   out.repeat_interleave(num_of_col, 2); out += D * x            mamba_simple_faster.py:356-358
//   the same for the b direction on x_flip                        :412-416
//   layernorm(rearrange(out + out_b.flip(-1)) / 2) * silu(z)      :434-441   (no-norm form :445-453)
// (>= 10 full-resolution passes incl. a flip and a (B,D,L)->(B,L,D) transpose copy) by ONE
// kernel that reads x and z once and writes the gated out_proj input once: 3 full-resolution
// tensors, the compulsory minimum.  The conv outputs xc_f / xc_b needed by the D skip are
// recomputed from x rather than stored by K1.
//
// Mapping (v4, persistent + pipelined): token-major; a tile is TT <= 8 consecutive SEQUENCE
// positions of one image with ALL d_inner channels (4 per thread); on plain grids a tile is an
// equal split of one pooled group (pool = 14 -> 2 x 7) so it reads a single pooled scan row.
// The grid is persistent (SM count x resident CTAs); each CTA walks tiles blockIdx.x, +gridDim.x...
// with a 2-stage cp.async pipeline: while tile i is computed, the TT+6 x rows (3 halo rows each
// side; rows outside the sequence zero-filled = the conv's padding) and TT z rows of tile i+1
// stream into the other shared-memory buffer as 16-byte LDGSTS, and its pooled scan row is
// prefetched into registers.  The per-channel parameters (8 taps, biases, D, gamma, beta) are
// loaded once per CTA.  Token -> memory-row tables (rotated layers, channel layouts) are built
// by one warp two tiles ahead, so the hot loops contain no integer division.
// Per tile: phase 1: sliding 7-row register window over the staged rows -> both convs + SiLU ->
// (s_f + s_b + D_f xc_f + D_b xc_b)/2 -> smem, per-token (sum, sumsq) partials by warp shuffle;
// one barrier; phase 2: LayerNorm * gamma + beta, * silu(z), 8-byte (bf16) / 16-byte (fp32)
// coalesced stores.
// In channel-sharded mode (2048^2 single image, d_inner split over GPUs) the CTA sees only a
// shard of the channels: it writes the pre-norm value and the shard's per-token partial
// statistics; fv_norm_gate_apply finishes after the all-reduce of the (B, L, 2) statistics.
#include "common.cuh"
#include "tiles.cuh"

namespace fv {

// Stages one tile with TMA bulk copies (warp 0, all lanes): lanes [0, np+6) own the x rows, lanes
// [TT+6, TT+6+np) the z rows; rows outside the sequence are zero-filled by their lane.
template <typename T, int TT>
__device__ __forceinline__ void bulk_stage_tile(int D, const TileTab<TT>* tab, const T* __restrict__ x,
                                                const T* __restrict__ z, int64_t ldxz, int64_t xzbs, T* buf,
                                                uint64_t* bar, bool with_z) {
    static_assert(2 * TT + 6 <= 32, "one lane per staged row");
    if (!tab->valid) return;
    const int lane = threadIdx.x & 31;
    const int np = tab->np;
    const uint32_t rowB = (uint32_t)D * sizeof(T);
    int row = -1;
    const T* src = nullptr;
    T* dst = nullptr;
    bool mine = false;
    if (lane < np + 6) {
        mine = true;
        row = tab->rows[lane];
        src = x;
        dst = buf + lane * D;
    } else if (with_z && lane >= TT + 6 && lane < TT + 6 + np) {
        mine = true;
        row = tab->rows[lane - (TT + 6) + 3];
        src = z;
        dst = buf + lane * D;  // z rows follow the TT+6 x rows
    }
    const unsigned copies = __ballot_sync(0xffffffffu, mine && row >= 0);
    if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)__popc(copies) * rowB);
    __syncwarp();
    if (mine) {
        if (row >= 0) {
            bulk_g2s(dst, src + (int64_t)tab->b * xzbs + (int64_t)row * ldxz, rowB, bar);
        } else {
            float4* d4 = reinterpret_cast<float4*>(dst);
            for (int i = 0; i < (int)(rowB / 16); ++i) d4[i] = zero4();
        }
    }
}

// MODE: 0 = LayerNorm + gate, 1 = gate only (use_norm_after_ssm=False), 2 = channel-sharded (pre-norm + stats)
template <typename T, int TT, int MODE, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
gate_fwd_kernel(Geom g, int64_t ntiles, int tiles_per_img, int tiles_per_group, int tile_len, int vec16, int nbuf,
                const T* __restrict__ x, const T* __restrict__ z, int64_t ldxz, int64_t xzbs,
                const float* __restrict__ s, const float* __restrict__ cw, const float* __restrict__ cb,
                const float* __restrict__ Dskip, const float* __restrict__ lnw, const float* __restrict__ lnb,
                float eps, T* __restrict__ y, int64_t ldy, int64_t ybs, float* __restrict__ stats) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    constexpr bool FAST = is_fast<T>::value;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nthreads = blockDim.x, nwarps = nthreads >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = g.D;
    // smem: vbuf[TT][D] fp32 | psum[TT][nthreads] float2 | stat[TT] float2 | nbuf x { xs[TT+6][D] T, zs[TT][D] T } |
    //       3 x TileTab | 2 mbarriers
    float* vbuf = reinterpret_cast<float*>(smem_raw);
    float2* psum = reinterpret_cast<float2*>(vbuf + TT * D);
    float2* stat = psum + TT * nthreads;
    T* stage0 = reinterpret_cast<T*>(stat + (TT + 1) / 2 * 2);
    const int stage_elems = (2 * TT + 6) * D;
    TileTab<TT>* tabs = reinterpret_cast<TileTab<TT>*>(stage0 + nbuf * stage_elems);
    uint64_t* bars = reinterpret_cast<uint64_t*>(tabs + 3);

    const int d0 = threadIdx.x * 4;
    const bool live = d0 < D;
    const int dd = live ? d0 : 0;
    const int64_t splane = (int64_t)g.B * g.Lp * D;
    const bool use_tma = vec16 != 0;

    // ---- per-CTA constants
    constexpr float PRE = FAST ? 0.5f : 1.f;
    const Taps tf = load_taps(cw, cb, D, 0, dd, PRE), tb = load_taps(cw, cb, D, 1, dd, PRE);
    // (s_f + s_b + D_f xc_f + D_b xc_b) / 2: the 1/2 is folded into D and s
    const float4 Df = scale4(ld4(Dskip + dd), 0.5f), Db = scale4(ld4(Dskip + D + dd), 0.5f);
    float4 gam = make_float4(1.f, 1.f, 1.f, 1.f), bet = zero4();
    if (MODE == 0) {
        gam = ld4(lnw + dd);
        if (lnb) bet = ld4(lnb + dd);
    }
    const float invD = 1.f / (float)D;

    auto issue_stage = [&](const TileTab<TT>* tab, int slot) {
        T* buf = stage0 + slot * stage_elems;
        if (use_tma) {
            if (warp == 0) bulk_stage_tile<T, TT>(D, tab, x, z, ldxz, xzbs, buf, &bars[slot], MODE != 2);
        } else {
            if (tab->valid) {
                const int np = tab->np;
                stage_rows(g, x + (int64_t)tab->b * xzbs, ldxz, tab->rows, np + 6, buf, false);
                if (MODE != 2)
                    stage_rows(g, z + (int64_t)tab->b * xzbs, ldxz, tab->rows + 3, np, buf + (TT + 6) * D, false);
            }
            cp_async_commit();
        }
    };
    auto wait_stage = [&](int slot, int use) {
        if (use_tma) mbar_wait(&bars[slot], (uint32_t)(use & 1));
        else cp_async_wait<0>();
    };

    // ---- prologue: barriers, tables of tiles 0 and 1, stage tile 0
    int64_t tile = blockIdx.x;
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        fill_tiletab<TT>(g, tile, ntiles, tiles_per_img, tiles_per_group, tile_len, ldy, ybs, &tabs[0]);
        fill_tiletab<TT>(g, tile + gridDim.x, ntiles, tiles_per_img, tiles_per_group, tile_len, ldy, ybs, &tabs[1]);
    }
    __syncthreads();
    issue_stage(&tabs[0], 0);
    float4 svf, svb;
    {
        const float* sp = s + ((int64_t)tabs[0].b * g.Lp + tabs[0].jt[0]) * D + dd;
        svf = ld4(sp);
        svb = ld4(sp + splane);
    }

    for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
        const TileTab<TT>* tab = &tabs[it % 3];
        const TileTab<TT>* tab1 = &tabs[(it + 1) % 3];
        const int slot = nbuf == 2 ? (it & 1) : 0;
        const T* xs = stage0 + slot * stage_elems;
        const T* zs = xs + (TT + 6) * D;
        if (nbuf == 1 && it > 0) {  // single buffer (very wide fp32 rows): no prefetch
            __syncthreads();
            issue_stage(tab, 0);
        }
        wait_stage(slot, nbuf == 2 ? (it >> 1) : it);
        __syncthreads();  // tile `it` has landed (and its zero-filled rows are visible); tile it-1 is done
        if (nbuf == 2) issue_stage(tab1, slot ^ 1);  // prefetch tile it+1 into the other buffer
        float4 sv = scale4(svf + svb, 0.5f);
        if (tab1->valid) {  // prefetch the pooled scan row of tile it+1
            const float* sp = s + ((int64_t)tab1->b * g.Lp + tab1->jt[0]) * D + dd;
            svf = ld4(sp);
            svb = ld4(sp + splane);
        }
        if (warp == 0)
            fill_tiletab<TT>(g, tile + 2 * (int64_t)gridDim.x, ntiles, tiles_per_img, tiles_per_group, tile_len, ldy,
                             ybs, &tabs[(it + 2) % 3]);
        const int np = tab->np;
        int jprev = tab->jt[0];

        // ---- phase 1: sliding 7-row register window -> pre-norm value v, per-thread LN partials.
        // Straight-line over all TT slots (slots >= np compute on stale rows and are never stored).
        float4 r[TT + 6];
        const T* xp = xs + dd;
        float* vp = vbuf + dd;
        float2* pp = psum + threadIdx.x;
        const T* zp = zs + dd;
#pragma unroll
        for (int k = 0; k < 6; ++k, xp += D) r[k] = ld4(xp);
#pragma unroll
        for (int p = 0; p < TT; ++p, xp += D, vp += D, pp += nthreads, zp += D) {
            r[p + 6] = ld4(xp);
            if (tiles_per_group == 0) {  // channel layouts: the pooled row may change inside a tile
                const int j = tab->jt[p];
                if (j != jprev) {
                    const float* sp = s + ((int64_t)tab->b * g.Lp + j) * D + dd;
                    sv = scale4(ld4(sp) + ld4(sp + splane), 0.5f);
                    jprev = j;
                }
            }
            float4 af, ab;
            conv_both_pre<FAST>(r[p], r[p + 1], r[p + 2], r[p + 3], r[p + 4], r[p + 5], r[p + 6], tf, tb, af, ab);
            float4 v = fma4(Db, ab, fma4(Df, af, sv));
            if (MODE == 1) {
                if (live && p < np) {
                    float4 zz = silu4<FAST>(ld4(zp));
                    st4(y + tab->yoff[p] + d0, make_float4(v.x * zz.x, v.y * zz.y, v.z * zz.z, v.w * zz.w));
                }
            } else {
                if (live) st4(vp, v);
                else v = zero4();
                *pp = make_float2((v.x + v.y) + (v.z + v.w), fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w))));
            }
        }
        if (MODE == 1) continue;
        __syncthreads();
        // ---- LN statistics: warp w reduces tokens w, w + nwarps, ... (independent shuffle chains)
        for (int p = warp; p < np; p += nwarps) {
            float sum = 0.f, sq = 0.f;
            for (int k = lane; k < nthreads; k += 32) {
                const float2 q = psum[p * nthreads + k];
                sum += q.x;
                sq += q.y;
            }
            sum = warp_sum(sum);
            sq = warp_sum(sq);
            if (lane == 0) {
                if (MODE == 2) {
                    stat[p] = make_float2(sum, sq);
                } else {
                    const float mean = sum * invD;
                    stat[p] = make_float2(mean, rsqrtf(fmaxf(sq * invD - mean * mean, 0.f) + eps));
                }
            }
        }
        __syncthreads();
        // ---- phase 2: normalise, gate, store
        vp = vbuf + dd;
        zp = zs + dd;
#pragma unroll
        for (int p = 0; p < TT; ++p, vp += D, zp += D) {
            if (p < np) {
                const float2 ms = stat[p];
                T* yo = y + tab->yoff[p] + d0;
                if (MODE == 2) {
                    if (threadIdx.x == 0) {
                        float* so = stats + ((int64_t)tab->b * g.L + tab->rows[p + 3]) * 2;
                        so[0] = ms.x;
                        so[1] = ms.y;
                    }
                    if (live) st4(yo, ld4(vp));
                } else if (live) {
                    const float4 v = ld4(vp);
                    const float4 zz = silu4<FAST>(ld4(zp));
                    const float4 gs = scale4(gam, ms.y);
                    float4 o;
                    o.x = fmaf(v.x - ms.x, gs.x, bet.x) * zz.x;
                    o.y = fmaf(v.y - ms.x, gs.y, bet.y) * zz.y;
                    o.z = fmaf(v.z - ms.x, gs.z, bet.z) * zz.z;
                    o.w = fmaf(v.w - ms.x, gs.w, bet.w) * zz.w;
                    st4(yo, o);
                }
            }
        }
    }
    if (!use_tma) cp_async_wait<0>();
}

// Finishes the channel-sharded path: y holds pre-norm values of this shard, stats the
// all-reduced per-token (sum, sumsq) over the full d_inner.
template <typename T>
__global__ void __launch_bounds__(256)
norm_gate_apply_kernel(Geom g, int full_dim, T* __restrict__ y, int64_t ldy, int64_t ybs,
                       const T* __restrict__ z, int64_t ldz, int64_t zbs,
                       const float* __restrict__ stats, const float* __restrict__ lnw,
                       const float* __restrict__ lnb, float eps) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    constexpr bool FAST = is_fast<T>::value;
    const int nvec = g.D >> 2;
    const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= (int64_t)g.B * g.L * nvec) return;
    const int d0 = (int)(item % nvec) * 4;
    const int64_t bt = item / nvec;
    const int64_t row = bt % g.L, b = bt / g.L;
    const float sum = stats[bt * 2], sq = stats[bt * 2 + 1];
    const float mean = sum / (float)full_dim;
    const float rstd = rsqrtf(fmaxf(sq / (float)full_dim - mean * mean, 0.f) + eps);
    float4 v = ld4(y + b * ybs + row * ldy + d0);
    float4 zz = silu4<FAST>(ld4(z + b * zbs + row * ldz + d0));
    float4 gam = lnw ? ld4(lnw + d0) : make_float4(1.f, 1.f, 1.f, 1.f), bet = lnb ? ld4(lnb + d0) : zero4();
    float4 o;
    if (lnw) {
        o.x = fmaf((v.x - mean) * rstd, gam.x, bet.x) * zz.x;
        o.y = fmaf((v.y - mean) * rstd, gam.y, bet.y) * zz.y;
        o.z = fmaf((v.z - mean) * rstd, gam.z, bet.z) * zz.z;
        o.w = fmaf((v.w - mean) * rstd, gam.w, bet.w) * zz.w;
    } else {
        o = make_float4(v.x * zz.x, v.y * zz.y, v.z * zz.z, v.w * zz.w);
    }
    st4(y + b * ybs + row * ldy + d0, o);
}

int check_geom(const fv_geom* g, const char* who);

template <typename T, int TT>
static size_t gate_smem(int D, int threads, int nbuf) {
    return (size_t)TT * D * 4 + (size_t)TT * threads * 8 + (size_t)((TT + 1) / 2 * 2) * 8 +
           (size_t)nbuf * (2 * TT + 6) * D * sizeof(T) + 3 * sizeof(TileTab<TT>) + 16;
}

template <typename T, int TT, int MODE>
static int launch_gate_tt(const Geom& g, int tpg, int tile_len, int nbuf, const T* x, const T* z, int64_t ldxz,
                          int64_t xzbs, const float* s, const float* cw, const float* cb, const float* Dskip,
                          const float* lnw, const float* lnb, float eps, T* y, int64_t ldy, int64_t ybs, float* stats,
                          cudaStream_t st) {
    const int threads = ((g.D / 4) + 31) / 32 * 32;
    const size_t smem = gate_smem<T, TT>(g.D, threads, nbuf);
    FV_REQUIRE(smem <= 227 * 1024, "fv_gate_fwd: shared memory %zu too large (dim %d)", smem, g.D);
    const int tiles_per_img = tpg > 0 ? g.Lp * tpg : ceil_div(g.L, TT);
    const int64_t ntiles = (int64_t)tiles_per_img * g.B;
    void (*kern)(Geom, int64_t, int, int, int, int, int, const T*, const T*, int64_t, int64_t, const float*,
                 const float*, const float*, const float*, const float*, const float*, float, T*, int64_t, int64_t,
                 float*);
    if (threads <= 128) kern = gate_fwd_kernel<T, TT, MODE, 128, 4>;
    else if (threads <= 256) kern = gate_fwd_kernel<T, TT, MODE, 256, 2>;
    else if (threads <= 512) kern = gate_fwd_kernel<T, TT, MODE, 512, 1>;
    else kern = gate_fwd_kernel<T, TT, MODE, 1024, 1>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        FV_REQUIRE(e == cudaSuccess, "fv_gate_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    int occ = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
    FV_REQUIRE(e == cudaSuccess && occ > 0, "fv_gate_fwd: occupancy query failed (%s)", cudaGetErrorString(e));
    const int64_t resident = (int64_t)sm_count() * occ;  // persistent grid: one wave
    dim3 grid((unsigned)(ntiles < resident ? ntiles : resident)), block(threads);
    const int vec16 = rows_vec16<T>(g.D, x, ldxz, xzbs) && ((uintptr_t)z % 16) == 0;
    FV_LAUNCH_PDL((kern), grid, block, smem, st, g, ntiles, tiles_per_img, tpg, tile_len, vec16, nbuf, x, z, ldxz, xzbs, s, cw, cb,
                                    Dskip, lnw, lnb, eps, y, ldy, ybs, stats);
    return finish_launch("gate_fwd");
}

template <typename T>
static int launch_gate(const Geom& g, const T* x, const T* z, int64_t ldxz, int64_t xzbs, const float* s,
                       const float* cw, const float* cb, const float* Dskip, const float* lnw,
                       const float* lnb, float eps, T* y, int64_t ldy, int64_t ybs, float* stats,
                       cudaStream_t st) {
    FV_REQUIRE(g.D <= 4096, "fv_gate_fwd: dim %d > 4096 not supported", g.D);
    // Tile length <= 8 tokens.  Plain grids: split each pooled group into equal tiles (pool = 14 -> 2 x 7)
    // so a tile reads one pooled scan row; channel layouts (inner > 1): runs of 8 sequence positions.
    const int threads = ((g.D / 4) + 31) / 32 * 32;
    const size_t budget = 200 * 1024;
    int maxlen = 8, nbuf = 2;
    if (gate_smem<T, 8>(g.D, threads, 2) > budget) {
        if (gate_smem<T, 4>(g.D, threads, 2) <= budget) maxlen = 4;
        else if (gate_smem<T, 8>(g.D, threads, 1) <= budget) nbuf = 1;
        else maxlen = 4, nbuf = 1;
    }
    int tpg = 0, tile_len = maxlen;
    if (g.inner == 1) {
        tpg = (g.pool + maxlen - 1) / maxlen;
        tile_len = (g.pool + tpg - 1) / tpg;
    }
#define FV_GATE_ARGS g, tpg, tile_len, nbuf, x, z, ldxz, xzbs, s, cw, cb, Dskip, lnw, lnb, eps, y, ldy, ybs, stats, st
#define FV_GATE_TT(TT_)                                                      \
    do {                                                                     \
        if (stats) return launch_gate_tt<T, TT_, 2>(FV_GATE_ARGS);           \
        if (lnw) return launch_gate_tt<T, TT_, 0>(FV_GATE_ARGS);             \
        return launch_gate_tt<T, TT_, 1>(FV_GATE_ARGS);                      \
    } while (0)
    if (tile_len <= 4) FV_GATE_TT(4);
    if (tile_len <= 7) FV_GATE_TT(7);
    FV_GATE_TT(8);
#undef FV_GATE_TT
#undef FV_GATE_ARGS
}

}  // namespace fv

extern "C" int fv_gate_fwd(const fv_geom* g_, int dtype, const void* x, const void* z, int64_t ldxz,
                           int64_t xz_bstride, const float* s, const float* conv_w,
                           const float* conv_b, const float* Dskip, const float* ln_w,
                           const float* ln_b, float eps, void* y, int64_t ldy, int64_t y_bstride,
                           float* stats, void* stream) {
    using namespace fv;
    if (int rc = check_geom(g_, "fv_gate_fwd")) return rc;
    FV_REQUIRE(x && z && s && conv_w && Dskip && y, "fv_gate_fwd: null pointer");
    FV_REQUIRE(ldxz % 4 == 0 && xz_bstride % 4 == 0 && ldy % 4 == 0 && y_bstride % 4 == 0,
               "fv_gate_fwd: strides must be multiples of 4 elements");
    Geom g = make_geom(g_);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FV_F32)
        return launch_gate<float>(g, (const float*)x, (const float*)z, ldxz, xz_bstride, s, conv_w, conv_b, Dskip, ln_w, ln_b, eps, (float*)y, ldy, y_bstride, stats, st);
    if (dtype == FV_BF16)
        return launch_gate<bf16>(g, (const bf16*)x, (const bf16*)z, ldxz, xz_bstride, s, conv_w, conv_b, Dskip, ln_w, ln_b, eps, (bf16*)y, ldy, y_bstride, stats, st);
    return fail("fv_gate_fwd: unsupported dtype %d", dtype);
}

extern "C" int fv_norm_gate_apply(const fv_geom* g_, int dtype, int full_dim, void* y, int64_t ldy,
                                  int64_t y_bstride, const void* z, int64_t ldz, int64_t z_bstride,
                                  const float* stats, const float* ln_w, const float* ln_b, float eps,
                                  void* stream) {
    using namespace fv;
    if (int rc = check_geom(g_, "fv_norm_gate_apply")) return rc;
    FV_REQUIRE(y && z && stats, "fv_norm_gate_apply: null pointer");
    FV_REQUIRE(full_dim >= g_->dim, "fv_norm_gate_apply: full_dim %d < dim %d", full_dim, g_->dim);
    Geom g = make_geom(g_);
    const int64_t items = (int64_t)g.B * g.L * (g.D / 4);
    dim3 grid((unsigned)((items + 255) / 256)), block(256);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FV_F32)
        FV_LAUNCH_PDL((norm_gate_apply_kernel<float>), grid, block, 0, st, g, full_dim, (float*)y, ldy, y_bstride, (const float*)z, ldz, z_bstride, stats, ln_w, ln_b, eps);
    else if (dtype == FV_BF16)
        FV_LAUNCH_PDL((norm_gate_apply_kernel<bf16>), grid, block, 0, st, g, full_dim, (bf16*)y, ldy, y_bstride, (const bf16*)z, ldz, z_bstride, stats, ln_w, ln_b, eps);
    else
        return fail("fv_norm_gate_apply: unsupported dtype %d", dtype);
    return finish_launch("norm_gate_apply");
}
