// K1-bwd -- backward of conv (both directions) + SiLU + mean pooling, merged with the D-skip path.
//
// Reference (paths relative to /root/reference): FastVim_MambaInnerFnNoOutProj_withoutZ.backward,
// mamba_ssm/ops/selective_scan_interface.py:662-668 (dconv_out = dout * D), :740-748 (un-pool:
// dconv_out += repeat_interleave(dconv_compressed) / num_of_col) and :751-753 (causal_conv1d_bwd of the
// un-vendored causal-conv1d package), once per direction on flipped tensors in the module
// (mamba_simple_faster.py:272-305).
//
// In original token coordinates, with e = dL/dv / 2 from the gate backward and
// du_f, du_b the gradients of the pooled conv outputs (scan backward + x_proj backward):
//   dxc_f[t] = e[t] D_f + du_f[j(t)] sf / pool          dxc_b[t] = e[t] D_b + du_b[j(t)] sf / pool
//   G_f[t]   = dxc_f[t] silu'(cf[t]),  cf[t] = b_f + sum_k w_f[k] x[t-3+k]
//   G_b[t]   = dxc_b[t] silu'(cb[t]),  cb[t] = b_b + sum_k w_b[k] x[t+3-k]
//   dx[t]    = sum_k w_f[k] G_f[t+3-k] + sum_k w_b[k] G_b[t-3+k]
//   dw_f[k]  = sum_t G_f[t] x[t-3+k],  dw_b[k] = sum_t G_b[t] x[t+3-k],  db_f = sum_t G_f[t],  db_b = sum_t G_b[t]
// One pass over x and e, one write of dx (next to dz in the d(xz) buffer).  Persistent grid; tile =
// TT <= 8 tokens of one pooled group, all channels (4 per thread); x and e rows (+3 halo rows each side)
// staged with cp.async; G_f / G_b live in thread-private shared-memory columns (no barrier between the
// two passes); weight / bias gradients accumulate in registers and leave with one atomicAdd per CTA.
#include <cstdlib>

#include "common.cuh"
#include "tiles.cuh"
#include "stream.cuh"

namespace fv {

template <typename T, int TT, int MAXT>
__global__ void __launch_bounds__(MAXT)
conv_pool_bwd_kernel(Geom g, int64_t ntiles, int tiles_per_img, int tiles_per_group, int tile_len, int vec16,
                     const T* __restrict__ x, int64_t ldx, int64_t xbs, const T* __restrict__ e,
                     const T* __restrict__ du, const float* __restrict__ cw, const float* __restrict__ cb,
                     const float* __restrict__ Dskip, float scale, T* __restrict__ dx,
                     float* __restrict__ dcw, float* __restrict__ dcb) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = g.D;
    constexpr int NR = TT + 6;
    // smem: Gf[NR][D] fp32 | Gb[NR][D] fp32 | xs[NR][D] T | es[NR][D] T | tab | jh[NR]
    float* Gf = reinterpret_cast<float*>(smem_raw);
    float* Gb = Gf + NR * D;
    T* xs = reinterpret_cast<T*>(Gb + NR * D);
    T* es = xs + NR * D;
    TileTab<TT>* tab = reinterpret_cast<TileTab<TT>*>(es + NR * D);
    int* jh = reinterpret_cast<int*>(tab + 1);

    const int d0 = threadIdx.x * 4;
    const bool live = d0 < D;
    const int dd = live ? d0 : 0;
    const int64_t uplane = (int64_t)g.B * g.Lp * D;
    const Taps tf = load_taps(cw, cb, D, 0, dd), tb = load_taps(cw, cb, D, 1, dd);
    const float4 Df = ld4(Dskip + dd), Db = ld4(Dskip + D + dd);
    const float pscale = scale / (float)g.pool;
    float4 awf[4], awb[4], abf = zero4(), abb = zero4();
#pragma unroll
    for (int k = 0; k < 4; ++k) awf[k] = zero4(), awb[k] = zero4();

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < 32) {
            fill_tiletab<TT>(g, tile, ntiles, tiles_per_img, tiles_per_group, tile_len, ldx, xbs, tab);
            // pooled position of every staged position (halo included)
            const int b_ = (int)(tile / tiles_per_img), rem = (int)(tile - (int64_t)b_ * tiles_per_img);
            const int jj = rem / tiles_per_group, q = rem - jj * tiles_per_group;
            const int t = jj * g.pool + q * tile_len - 3 + (int)threadIdx.x;
            if (threadIdx.x < NR) jh[threadIdx.x] = (t >= 0 && t < g.L) ? t / g.pool : 0;
        }
        __syncthreads();
        const int np = tab->np, b = tab->b;
        stage_rows(g, x + (int64_t)b * xbs, ldx, tab->rows, np + 6, xs, vec16 & 1);
        stage_rows(g, e + (int64_t)b * g.L * D, (int64_t)D, tab->rows, np + 6, es, (vec16 >> 1) & 1);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        if (!live) continue;  // (all barriers of the next iteration are reached by every thread)

        // ---- pass 1: G_f on positions [3, np+6), G_b on [0, np+3)
        const T* dub = du + (int64_t)b * g.Lp * D + d0;
        int jprev = -1;
        float4 duf = zero4(), dub4 = zero4();
        for (int i = 0; i < np + 6; ++i) {
            float4 gf = zero4(), gb = zero4();
            if (tab->rows[i] >= 0) {
                const int j = jh[i];
                if (j != jprev) {
                    duf = scale4(ld4(dub + (int64_t)j * D), pscale);
                    dub4 = scale4(ld4(dub + uplane + (int64_t)j * D), pscale);
                    jprev = j;
                }
                const float4 ev = ld4(es + i * D + d0);
                if (i >= 3) {
                    float4 c = tf.b;
#pragma unroll
                    for (int k = 0; k < 4; ++k) c = fma4(tf.w[k], ld4(xs + (i - 3 + k) * D + d0), c);
                    const float4 dxc = fma4(ev, Df, duf);
                    gf = make_float4(dxc.x * dsilu(c.x), dxc.y * dsilu(c.y), dxc.z * dsilu(c.z), dxc.w * dsilu(c.w));
                }
                if (i < np + 3) {
                    float4 c = tb.b;
#pragma unroll
                    for (int k = 0; k < 4; ++k) c = fma4(tb.w[k], ld4(xs + (i + 3 - k) * D + d0), c);
                    const float4 dxc = fma4(ev, Db, dub4);
                    gb = make_float4(dxc.x * dsilu(c.x), dxc.y * dsilu(c.y), dxc.z * dsilu(c.z), dxc.w * dsilu(c.w));
                }
            }
            st4(Gf + i * D + d0, gf);
            st4(Gb + i * D + d0, gb);
        }
        // ---- pass 2: dx and the weight / bias gradients of the owned positions [3, np+3)
        for (int p = 0; p < np; ++p) {
            const int i = p + 3;
            float4 acc = zero4();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                acc = fma4(tf.w[k], ld4(Gf + (i + 3 - k) * D + d0), acc);
                acc = fma4(tb.w[k], ld4(Gb + (i - 3 + k) * D + d0), acc);
            }
            st4(dx + tab->yoff[p] + d0, acc);
            const float4 gf = ld4(Gf + i * D + d0), gb = ld4(Gb + i * D + d0);
            abf = abf + gf;
            abb = abb + gb;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                awf[k] = fma4(gf, ld4(xs + (i - 3 + k) * D + d0), awf[k]);
                awb[k] = fma4(gb, ld4(xs + (i + 3 - k) * D + d0), awb[k]);
            }
        }
    }
    if (live) {
        // conv_w layout (2, D, 4): channel-major, tap fastest
        const float f[4][4] = {{awf[0].x, awf[1].x, awf[2].x, awf[3].x}, {awf[0].y, awf[1].y, awf[2].y, awf[3].y},
                               {awf[0].z, awf[1].z, awf[2].z, awf[3].z}, {awf[0].w, awf[1].w, awf[2].w, awf[3].w}};
        const float r[4][4] = {{awb[0].x, awb[1].x, awb[2].x, awb[3].x}, {awb[0].y, awb[1].y, awb[2].y, awb[3].y},
                               {awb[0].z, awb[1].z, awb[2].z, awb[3].z}, {awb[0].w, awb[1].w, awb[2].w, awb[3].w}};
        const float bf_[4] = {abf.x, abf.y, abf.z, abf.w}, bb_[4] = {abb.x, abb.y, abb.z, abb.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                atomicAdd(dcw + ((int64_t)(d0 + c)) * 4 + k, f[c][k]);
                atomicAdd(dcw + ((int64_t)D + d0 + c) * 4 + k, r[c][k]);
            }
            if (dcb) {
                atomicAdd(dcb + d0 + c, bf_[c]);
                atomicAdd(dcb + D + d0 + c, bb_[c]);
            }
        }
    }
}

// ---- streaming variant (default) -------------------------------------------------------------------------------------
// The tiled kernel above stages TT + 6 token rows of ALL channels per tile: at dim 1536 that is 184 KB of shared memory
// for TT = 4 -- one 384-thread CTA per SM, 2.5x halo re-reads and a full stage -> wait -> compute serialisation
// (465 us per launch at FastVim-B, 12x its HBM time).  But the conv only couples tokens, never channels, so a thread
// that owns TWO channels can simply walk a run of consecutive tokens with everything in registers:
//   at front position f (x[f], e[f] just arrived) it forms G_f[f] from x[f-3..f], G_b[f-3] from the SAME four x rows,
//   dx[f-3] from G_f[f-3..f] and G_b[f-6..f-3], and the weight-gradient products of both -- four 4-deep register windows
//   (x, dxc_b, G_f, G_b), statically indexed by unrolling the walk by 4; no shared-memory data, no barrier in the walk.
// Loads run 4 tokens ahead of their use through a register queue (8 independent 4-byte loads in flight per thread).
// A CTA (<= 256 channel pairs) walks runs of <= 56 tokens (3-token halo each side: ~12 % extra L2 reads and SiLU
// derivatives; the walk itself is branch-free, ownership of halo positions is applied with selects) and keeps the conv weight / bias gradient partials in registers across all its runs: one atomic per CTA
// and parameter at the end.  The only shared memory is the run's row table (token -> memory row, pooled index).
// d silu / dx; FAST (bf16 I/O): sigmoid through one tanh.approx MUFU op
template <bool FAST>
__device__ __forceinline__ float dsilu_sel(float x) {
    if (FAST) {
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
        const float s = fmaf(0.5f, t, 0.5f);
        return s * fmaf(x, 1.f - s, 1.f);
    }
    return dsilu(x);
}

constexpr int CBS_MAXRUN = 64;   // longest token run of one work item (row table size - 8)

// d silu / dx of a pair (see dsilu_sel), on the packed f32x2 pipe except for the two MUFU ops
template <bool FAST>
__device__ __forceinline__ float2 dsilu2(float2 x) {
    if (FAST) {
        const float2 h = __fmul2_rn(x, make_float2(0.5f, 0.5f));
        float tx, ty;
        asm("tanh.approx.f32 %0, %1;" : "=f"(tx) : "f"(h.x));
        asm("tanh.approx.f32 %0, %1;" : "=f"(ty) : "f"(h.y));
        const float2 s = __ffma2_rn(make_float2(tx, ty), make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));
        const float2 oms = __ffma2_rn(s, make_float2(-1.f, -1.f), make_float2(1.f, 1.f));
        return __fmul2_rn(s, __ffma2_rn(x, oms, make_float2(1.f, 1.f)));
    }
    return make_float2(dsilu(x.x), dsilu(x.y));
}
// the same with silu(x) itself as a second result (needed for the D-skip gradient dD = sum e * silu(conv))
template <bool FAST>
__device__ __forceinline__ float2 dsilu2_val(float2 x, float2& sl) {
    if (FAST) {
        const float2 h = __fmul2_rn(x, make_float2(0.5f, 0.5f));
        float tx, ty;
        asm("tanh.approx.f32 %0, %1;" : "=f"(tx) : "f"(h.x));
        asm("tanh.approx.f32 %0, %1;" : "=f"(ty) : "f"(h.y));
        const float2 s = __ffma2_rn(make_float2(tx, ty), make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));
        const float2 oms = __ffma2_rn(s, make_float2(-1.f, -1.f), make_float2(1.f, 1.f));
        sl = __fmul2_rn(x, s);
        return __fmul2_rn(s, __ffma2_rn(x, oms, make_float2(1.f, 1.f)));
    }
    sl = make_float2(silu_exact(x.x), silu_exact(x.y));
    return make_float2(dsilu(x.x), dsilu(x.y));
}

// WITH_DD: also accumulate the D-skip gradients dD_f = sum e * xc_f, dD_b = sum e * xc_b (the conv outputs are formed here
// anyway; the streaming gate backward fv_gate_bwd_v works from the saved pre-norm value and never sees them)
template <typename T, bool WITH_DD>
__global__ void __launch_bounds__(256, 2)
conv_pool_bwd_stream_kernel(Geom g, int nseg, int seg_len, int64_t nitems, int nslots, int chunks,
                            const T* __restrict__ x, int64_t ldx, int64_t xbs, const T* __restrict__ e,
                            const T* __restrict__ du, const float* __restrict__ cw, const float* __restrict__ cb,
                            const float* __restrict__ Dskip, float scale, T* __restrict__ dx,
                            float* __restrict__ dcw, float* __restrict__ dcb, float* __restrict__ dDs) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    typedef Pair<T> P;
    typedef typename P::type PT;
    constexpr bool FAST = is_fast<T>::value;
    // per-run tables, padded by 8 entries so the unrolled walk never indexes past them:
    // element offset of the token's row in x / dx and in e (-1: outside the image), pooled index
    __shared__ int xoff[CBS_MAXRUN + 16], eoff[CBS_MAXRUN + 16], jtab[CBS_MAXRUN + 16];
    const int D = g.D;
    const int slot = blockIdx.x / chunks, chunk = blockIdx.x - slot * chunks;
    const int d0 = (chunk * blockDim.x + threadIdx.x) * 2;
    const int64_t uplane = (int64_t)g.B * g.Lp * D;
    const float pscale = scale / (float)g.pool;
    const float2 ps2 = make_float2(pscale, pscale);
    float2 wf[4], wb[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        wf[k] = make_float2(cw[(int64_t)d0 * 4 + k], cw[(int64_t)(d0 + 1) * 4 + k]);
        wb[k] = make_float2(cw[((int64_t)D + d0) * 4 + k], cw[((int64_t)D + d0 + 1) * 4 + k]);
    }
    const float2 bf_ = cb ? make_float2(cb[d0], cb[d0 + 1]) : make_float2(0.f, 0.f);
    const float2 bb_ = cb ? make_float2(cb[D + d0], cb[D + d0 + 1]) : make_float2(0.f, 0.f);
    const float2 Df = make_float2(Dskip[d0], Dskip[d0 + 1]), Db = make_float2(Dskip[D + d0], Dskip[D + d0 + 1]);
    const float2 z2 = make_float2(0.f, 0.f);
    float2 awf[4] = {z2, z2, z2, z2}, awb[4] = {z2, z2, z2, z2}, abf = z2, abb = z2;
    float2 aDf = z2, aDb = z2;

    for (int64_t item = slot; item < nitems; item += nslots) {
        const int b = (int)(item / nseg), sgi = (int)(item - (int64_t)b * nseg);
        const int t0 = sgi * seg_len, n = min(seg_len, g.L - t0), n6 = n + 6;
        __syncthreads();  // the previous run's tables are no longer read
        for (int i = threadIdx.x; i < n6 + 8; i += blockDim.x) {
            const int t = t0 - 3 + i;
            const bool in = i < n6 && t >= 0 && t < g.L;
            const int row = in ? (int)seq_to_row(g, t) : 0;
            xoff[i] = in ? row * (int)ldx : -1;
            eoff[i] = in ? row * D : -1;
            jtab[i] = min(max(t, 0), g.L - 1) / g.pool;   // clamped: no spurious du reload outside the image
        }
        __syncthreads();
        const T* xb = x + (int64_t)b * xbs + d0;
        const T* eb = e + (int64_t)b * g.L * D + d0;
        T* dxb = dx + (int64_t)b * xbs + d0;
        const T* duf_b = du + (int64_t)b * g.Lp * D + d0;
        PT qx[4], qe[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int ox = xoff[k], oe = eoff[k];
            qx[k] = ox >= 0 ? P::ld(xb + ox) : P::zero();
            qe[k] = oe >= 0 ? P::ld(eb + oe) : P::zero();
        }
        float2 xw[4] = {z2, z2, z2, z2}, cbw[4] = {z2, z2, z2, z2}, gfw[4] = {z2, z2, z2, z2}, gbw[4] = {z2, z2, z2, z2};
        float2 ew[4] = {z2, z2, z2, z2};   // e[f-m] in slot (k - m) & 3 (WITH_DD only)
        float2 duf = z2, dub = z2;
        int jcur = -1;
        for (int base = 0; base < n6; base += 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = base + k;                                  // may run up to 3 past n6: tables are padded
                const PT px = qx[k], pe = qe[k];
                {
                    const int ox = xoff[i + 4], oe = eoff[i + 4];
                    qx[k] = ox >= 0 ? P::ld(xb + ox) : P::zero();
                    qe[k] = oe >= 0 ? P::ld(eb + oe) : P::zero();
                }
                const bool inr = xoff[i] >= 0;
                const float2 ev = P::up(pe);
                xw[k] = P::up(px);                                       // x[f]; x[f-m] sits in slot (k - m) & 3
                const int j = jtab[i];
                if (j != jcur) {
                    jcur = j;
                    duf = __fmul2_rn(P::up(P::ld(duf_b + (int64_t)j * D)), ps2);
                    dub = __fmul2_rn(P::up(P::ld(duf_b + uplane + (int64_t)j * D)), ps2);
                }
                const float2 cbv = __ffma2_rn(ev, Db, dub);
                cbw[k] = inr ? cbv : z2;                                 // dxc_b[f]
                const float2 x3 = xw[(k + 1) & 3], x2 = xw[(k + 2) & 3], x1 = xw[(k + 3) & 3], x0 = xw[k];  // x[f-3..f]
                const float2 cf = __ffma2_rn(wf[3], x0, __ffma2_rn(wf[2], x1, __ffma2_rn(wf[1], x2, __ffma2_rn(wf[0], x3, bf_))));
                const float2 cbk = __ffma2_rn(wb[3], x3, __ffma2_rn(wb[2], x2, __ffma2_rn(wb[1], x1, __ffma2_rn(wb[0], x0, bb_))));
                float2 Gf, Gb;
                if (WITH_DD) {
                    ew[k] = inr ? ev : z2;
                    float2 slf, slb;
                    Gf = __fmul2_rn(__ffma2_rn(ev, Df, duf), dsilu2_val<FAST>(cf, slf));  // G_f[f]
                    Gb = __fmul2_rn(cbw[(k + 1) & 3], dsilu2_val<FAST>(cbk, slb));        // G_b[f-3] (0 outside)
                    if (i >= 3 && i < n + 3 && inr) aDf = __ffma2_rn(ev, slf, aDf);       // xc_f[f] = silu(cf)
                    if (i >= 6 && i < n6) aDb = __ffma2_rn(ew[(k + 1) & 3], slb, aDb);    // xc_b[f-3] = silu(cbk)
                } else {
                    Gf = __fmul2_rn(__ffma2_rn(ev, Df, duf), dsilu2<FAST>(cf));           // G_f[f]
                    Gb = __fmul2_rn(cbw[(k + 1) & 3], dsilu2<FAST>(cbk));                 // G_b[f-3] (0 outside)
                }
                if (!inr) Gf = z2;
                gfw[k] = Gf;
                gbw[k] = Gb;
                // ownership: G_f[f] for f in [t0, t0+n) <=> 3 <= i < n+3;  token t = f-3 (dx, G_b) <=> 6 <= i < n+6
                const float2 Gfo = (i >= 3 && i < n + 3) ? Gf : z2;
                const bool own = i >= 6 && i < n6;
                const float2 Gbo = own ? Gb : z2;
                awf[0] = __ffma2_rn(Gfo, x3, awf[0]); awf[1] = __ffma2_rn(Gfo, x2, awf[1]);
                awf[2] = __ffma2_rn(Gfo, x1, awf[2]); awf[3] = __ffma2_rn(Gfo, x0, awf[3]);
                abf = __fadd2_rn(abf, Gfo);
                awb[0] = __ffma2_rn(Gbo, x0, awb[0]); awb[1] = __ffma2_rn(Gbo, x1, awb[1]);
                awb[2] = __ffma2_rn(Gbo, x2, awb[2]); awb[3] = __ffma2_rn(Gbo, x3, awb[3]);
                abb = __fadd2_rn(abb, Gbo);
                float2 acc = __fmul2_rn(wf[0], gfw[k]);
                acc = __ffma2_rn(wf[1], gfw[(k + 3) & 3], acc);
                acc = __ffma2_rn(wf[2], gfw[(k + 2) & 3], acc);
                acc = __ffma2_rn(wf[3], gfw[(k + 1) & 3], acc);
                acc = __ffma2_rn(wb[3], gbw[k], acc);
                acc = __ffma2_rn(wb[2], gbw[(k + 3) & 3], acc);
                acc = __ffma2_rn(wb[1], gbw[(k + 2) & 3], acc);
                acc = __ffma2_rn(wb[0], gbw[(k + 1) & 3], acc);
                if (own) P::st(dxb + xoff[i - 3], acc);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        atomicAdd(dcw + (int64_t)d0 * 4 + k, awf[k].x);
        atomicAdd(dcw + (int64_t)(d0 + 1) * 4 + k, awf[k].y);
        atomicAdd(dcw + ((int64_t)D + d0) * 4 + k, awb[k].x);
        atomicAdd(dcw + ((int64_t)D + d0 + 1) * 4 + k, awb[k].y);
    }
    if (dcb) {
        atomicAdd(dcb + d0, abf.x); atomicAdd(dcb + d0 + 1, abf.y);
        atomicAdd(dcb + D + d0, abb.x); atomicAdd(dcb + D + d0 + 1, abb.y);
    }
    if (WITH_DD) {
        atomicAdd(dDs + d0, aDf.x); atomicAdd(dDs + d0 + 1, aDf.y);
        atomicAdd(dDs + D + d0, aDb.x); atomicAdd(dDs + D + d0 + 1, aDb.y);
    }
}

template <typename T>
static int launch_conv_bwd_stream(const Geom& g, const T* x, int64_t ldx, int64_t xbs, const T* e, const T* du,
                                  const float* cw, const float* cb, const float* Dskip, float scale, T* dx, float* dcw,
                                  float* dcb, float* dDs, cudaStream_t st) {
    const int threads = stream_block(g.D), chunks = (g.D / 2) / threads;
    const int nseg = ceil_div(g.L, 56), seg_len = ceil_div(g.L, nseg);   // equal runs of <= 56 tokens
    const int64_t nitems = (int64_t)g.B * nseg;
    auto kern = dDs ? conv_pool_bwd_stream_kernel<T, true> : conv_pool_bwd_stream_kernel<T, false>;
    int occ = 0;
    cudaError_t er = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, 0);
    FV_REQUIRE(er == cudaSuccess && occ > 0, "fv_conv_pool_bwd: occupancy query failed (%s)", cudaGetErrorString(er));
    int64_t nslots = ((int64_t)sm_count() * occ) / chunks;
    if (nslots < 1) nslots = 1;
    if (nslots > nitems) nslots = nitems;
    // even number of runs per slot where possible (no ragged last round)
    const int64_t rounds = (nitems + nslots - 1) / nslots;
    nslots = (nitems + rounds - 1) / rounds;
    FV_LAUNCH_PDL((kern), (unsigned)(nslots * chunks), threads, 0, st, g, nseg, seg_len, nitems, (int)nslots, chunks, x, ldx, xbs, e, du, cw,
                                                          cb, Dskip, scale, dx, dcw, dcb, dDs);
    return finish_launch("conv_pool_bwd");
}

int check_geom(const fv_geom* g, const char* who);

template <typename T, int TT>
static size_t conv_bwd_smem(int D) {
    return (size_t)2 * (TT + 6) * D * 4 + (size_t)2 * (TT + 6) * D * sizeof(T) + sizeof(TileTab<TT>) + (TT + 6) * 4;
}

template <typename T, int TT>
static int launch_conv_bwd(const Geom& g, int tpg, int tile_len, const T* x, int64_t ldx, int64_t xbs, const T* e,
                           const T* du, const float* cw, const float* cb, const float* Dskip, float scale, T* dx,
                           float* dcw, float* dcb, cudaStream_t st) {
    const int threads = ((g.D / 4) + 31) / 32 * 32;
    const size_t smem = conv_bwd_smem<T, TT>(g.D);
    FV_REQUIRE(smem <= 227 * 1024, "fv_conv_pool_bwd: shared memory %zu too large (dim %d)", smem, g.D);
    const int tiles_per_img = g.Lp * tpg;
    const int64_t ntiles = (int64_t)tiles_per_img * g.B;
    void (*kern)(Geom, int64_t, int, int, int, int, const T*, int64_t, int64_t, const T*, const T*, const float*,
                 const float*, const float*, float, T*, float*, float*);
    if (threads <= 128) kern = conv_pool_bwd_kernel<T, TT, 128>;
    else if (threads <= 256) kern = conv_pool_bwd_kernel<T, TT, 256>;
    else if (threads <= 512) kern = conv_pool_bwd_kernel<T, TT, 512>;
    else kern = conv_pool_bwd_kernel<T, TT, 1024>;
    if (smem > 48 * 1024) {
        cudaError_t er = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        FV_REQUIRE(er == cudaSuccess, "fv_conv_pool_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(er));
    }
    int occ = 0;
    cudaError_t er = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
    FV_REQUIRE(er == cudaSuccess && occ > 0, "fv_conv_pool_bwd: occupancy query failed (%s)", cudaGetErrorString(er));
    const int64_t resident = (int64_t)sm_count() * occ;
    dim3 grid((unsigned)(ntiles < resident ? ntiles : resident)), block(threads);
    const int vec16 = (rows_vec16<T>(g.D, x, ldx, xbs) ? 1 : 0) | (rows_vec16<T>(g.D, e, g.D, (int64_t)g.L * g.D) ? 2 : 0);
    FV_LAUNCH_PDL((kern), grid, block, smem, st, g, ntiles, tiles_per_img, tpg, tile_len, vec16, x, ldx, xbs, e, du, cw, cb, Dskip,
                                    scale, dx, dcw, dcb);
    return finish_launch("conv_pool_bwd");
}

}  // namespace fv

extern "C" int fv_conv_pool_bwd(const fv_geom* g_, int dtype, const void* x, int64_t ldx, int64_t x_bstride,
                                const void* e, const void* du, const float* conv_w, const float* conv_b,
                                const float* Dskip, float scale, int pool_mode, void* dx, float* dconv_w,
                                float* dconv_b, float* dDskip, void* stream) {
    using namespace fv;
    if (int rc = check_geom(g_, "fv_conv_pool_bwd")) return rc;
    FV_REQUIRE(x && e && du && conv_w && Dskip && dx && dconv_w, "fv_conv_pool_bwd: null pointer");
    FV_REQUIRE(pool_mode == FV_POOL_MEAN, "fv_conv_pool_bwd: only mean pooling has a backward (as in the reference's fused path)");
    FV_REQUIRE(g_->inner == 1, "fv_conv_pool_bwd: channel layouts (inner > 1) are forward-only");
    FV_REQUIRE(ldx % 4 == 0 && x_bstride % 4 == 0, "fv_conv_pool_bwd: strides must be multiples of 4 elements");
    FV_REQUIRE(g_->dim <= 4096 && g_->batch <= 65535, "fv_conv_pool_bwd: dim > 4096 or batch > 65535");
    Geom g = make_geom(g_);
    cudaStream_t st = (cudaStream_t)stream;
    static const bool tiled_only = getenv("FASTVIM_CONV_BWD_TILED") != nullptr;   // A/B switch for tools/kbench.py
    if (!tiled_only && stream_block(g.D) > 0 && ldx % 2 == 0 && x_bstride % 2 == 0) {
        if (dtype == FV_F32)
            return launch_conv_bwd_stream<float>(g, (const float*)x, ldx, x_bstride, (const float*)e, (const float*)du, conv_w,
                                                 conv_b, Dskip, scale, (float*)dx, dconv_w, dconv_b, dDskip, st);
        if (dtype == FV_BF16)
            return launch_conv_bwd_stream<bf16>(g, (const bf16*)x, ldx, x_bstride, (const bf16*)e, (const bf16*)du, conv_w,
                                                conv_b, Dskip, scale, (bf16*)dx, dconv_w, dconv_b, dDskip, st);
        return fail("fv_conv_pool_bwd: unsupported dtype %d", dtype);
    }
    FV_REQUIRE(!dDskip, "fv_conv_pool_bwd: dDskip is produced by the streaming kernel only (dim %% 64 == 0, even strides)");
    const size_t budget = 200 * 1024;
    int maxlen = 8;
    if ((dtype == FV_F32 ? conv_bwd_smem<float, 8>(g.D) : conv_bwd_smem<bf16, 8>(g.D)) > budget) maxlen = 4;
    if (maxlen == 4 && (dtype == FV_F32 ? conv_bwd_smem<float, 4>(g.D) : conv_bwd_smem<bf16, 4>(g.D)) > budget) maxlen = 2;
    const int tpg = (g.pool + maxlen - 1) / maxlen;
    const int tile_len = (g.pool + tpg - 1) / tpg;
#define FV_CB(T_)                                                                                                      \
    do {                                                                                                               \
        if (tile_len <= 2)                                                                                             \
            return launch_conv_bwd<T_, 2>(g, tpg, tile_len, (const T_*)x, ldx, x_bstride, (const T_*)e, (const T_*)du,  \
                                          conv_w, conv_b, Dskip, scale, (T_*)dx, dconv_w, dconv_b, st);                \
        if (tile_len <= 4)                                                                                             \
            return launch_conv_bwd<T_, 4>(g, tpg, tile_len, (const T_*)x, ldx, x_bstride, (const T_*)e, (const T_*)du,  \
                                          conv_w, conv_b, Dskip, scale, (T_*)dx, dconv_w, dconv_b, st);                \
        if (tile_len <= 7)                                                                                             \
            return launch_conv_bwd<T_, 7>(g, tpg, tile_len, (const T_*)x, ldx, x_bstride, (const T_*)e, (const T_*)du,  \
                                          conv_w, conv_b, Dskip, scale, (T_*)dx, dconv_w, dconv_b, st);                \
        return launch_conv_bwd<T_, 8>(g, tpg, tile_len, (const T_*)x, ldx, x_bstride, (const T_*)e, (const T_*)du,      \
                                      conv_w, conv_b, Dskip, scale, (T_*)dx, dconv_w, dconv_b, st);                    \
    } while (0)
    if (dtype == FV_F32) FV_CB(float);
    if (dtype == FV_BF16) FV_CB(bf16);
#undef FV_CB
    return fail("fv_conv_pool_bwd: unsupported dtype %d", dtype);
}
