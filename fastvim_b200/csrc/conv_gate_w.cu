// Channel layouts (inner > 1: FastChannelVim Channel-First, mamba_simple_channel_faster.py:225-256, 325-340) on staged,
// streaming kernels.  The generic four-launch kernels give every thread its own 7-row window re-read through L1 (K1) and
// re-evaluate both convolutions in the gate (K2b): 137 + 182 us per block at the JUMP-CP shape (32 x 1568 x 768), 10-20 %
// of the HBM roofline.  Here
//   fv_conv_pool_w_fwd  CTA = one (image, outer) group = pool * inner consecutive sequence positions, staged through shared
//                       memory in double-buffered chunks (cp.async, 3-token halos, zero padding); thread = (inner slot, 4
//                       channels) walks its pooled group with the stride `inner`, accumulates the pool in registers and
//                       ALSO writes the D-skip term  w = (D_f xc_f + D_b xc_b) / 2  of every token (bf16) --
//                       x is read once, nothing is recomputed later;
//   fv_gate_w_fwd       one warp per token: v = w + (s_f[j] + s_b[j]) / 2, LayerNorm over d_inner by shuffles,
//                       * silu(z), y -- a pure streaming pass (w, z in; y out).
// Reference semantics: x.flip / causal_conv1d / reshape.mean (:258-289), repeat_interleave + D skip (:325-340),
// (out + out_b.flip) / 2, LayerNorm, * silu(z) (:400-420) -- same arithmetic as block_fwd.cu (w is rounded to bf16 there too).
#include <cstdlib>

#include "block_common.cuh"

namespace fv {

int sm_count();
int check_geom(const fv_geom* g, const char* who);
int conv_pool_plain_w(const Geom& g, const bf16* x, int64_t ldx, int64_t xbs, const float* cw, const float* cb, float scale,
                      int pool_mode, bf16* u, const float* Dskip, bf16* wout, cudaStream_t st);   // conv_pool.cu

constexpr int CG_MAX_NI = 4;  // inner slots per thread (inner / IH): pool accumulators stay in registers

template <bool MAXPOOL, int NIT, int IHT>
__global__ void __launch_bounds__(IHT > 0 ? 384 : 512)   // sliding form: window + taps + accumulators need > 128 registers
conv_pool_group_kernel(Geom g, int IH, int TP, const bf16* __restrict__ x, int64_t ldx, int64_t xbs,
                       const float* __restrict__ cw, const float* __restrict__ cb, float scale,
                       const float* __restrict__ Dskip, bf16* __restrict__ u, bf16* __restrict__ wout) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ __align__(16) unsigned char cgw_smem[];
    const int D = g.D, inner = g.inner, pool = g.pool;
    const int o = blockIdx.x, b = blockIdx.y;
    const int ncg = D >> 2;                 // 4-channel groups per token row
    const int tid = threadIdx.x;
    const int ih = tid / ncg, c4 = tid - ih * ncg;
    const bool live = ih < IH;
    const int d0 = c4 * 4;
    const int ntok = pool * inner, tbase = o * ntok;
    const int TPI = TP * inner;             // sequence positions per chunk
    const int bufrows = TPI + 6;
    bf16* xs = reinterpret_cast<bf16*>(cgw_smem);                                   // [2][bufrows][D]
    int* rowtab = reinterpret_cast<int*>(xs + (size_t)2 * bufrows * D);            // [ntok + 6]
    const bf16* xb = x + (int64_t)b * xbs;
    const int nchunk = (pool + TP - 1) / TP;
    const int NI = inner / IH;              // host guarantees inner % IH == 0 and NI <= NIT

    fill_rowtab(g, tbase - 3, ntok + 6, rowtab);
    __syncthreads();
    stage_rows(g, xb, ldx, rowtab, min(TP, pool) * inner + 6, xs, true);
    cp_async_commit();
    const Taps tf = load_taps(cw, cb, D, 0, live ? d0 : 0, 0.5f), tb = load_taps(cw, cb, D, 1, live ? d0 : 0, 0.5f);
    float4 Df = zero4(), Db = zero4();
    if (wout && live) {
        Df = scale4(ld4(Dskip + d0), 0.5f);
        Db = scale4(ld4(Dskip + D + d0), 0.5f);
    }
    float4 accf[NIT], accb[NIT];
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
        accf[k] = MAXPOOL ? make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY) : zero4();
        accb[k] = accf[k];
    }
    bf16* wb_ = wout ? wout + (int64_t)b * g.L * D : nullptr;
    for (int c = 0; c < nchunk; ++c) {
        const int p_lo = c * TP, np = min(TP, pool - p_lo);
        if (c + 1 < nchunk) {
            const int q_lo = p_lo + TP, nq = min(TP, pool - q_lo);
            stage_rows(g, xb, ldx, rowtab + q_lo * inner, nq * inner + 6, xs + (size_t)((c + 1) & 1) * bufrows * D, true);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (live && IHT > 0) {
            // IH known at compile time: the thread's inner slots of one pooled step are IHT rows apart, so their 7-row windows
            // overlap -- slide one window (7 - IHT rows carried in registers, IHT new rows unpacked per token) instead of
            // fetching and unpacking seven rows per token (ncu: the ALU pipe, bf16 unpacks, was this kernel's busiest)
            const bf16* cur = xs + (size_t)(c & 1) * bufrows * D + d0;
            for (int p = 0; p < np; ++p) {
                const int lt0 = p * inner + ih;                   // first slot's token; buffer rows lt0 .. lt0 + 6
                const bf16* xp = cur + (size_t)lt0 * D;
                float4 win[7];
#pragma unroll
                for (int r = 0; r < 7; ++r) win[r] = ld4(xp + (size_t)r * D);
#pragma unroll
                for (int k = 0; k < NIT; ++k) {
                    if (k < NI) {
                        if (k > 0) {
#pragma unroll
                            for (int r = 0; r < 7; ++r)
                                win[r] = r + IHT < 7 ? win[(r + IHT) % 7] : ld4(xp + (size_t)(k * IHT + r) * D);
                        }
                        float4 af = tf.b, ab = tb.b;
#pragma unroll
                        for (int r = 0; r < 7; ++r) {
                            if (r <= 3) af = fma4(tf.w[r], win[r], af);
                            if (r >= 3) ab = fma4(tb.w[6 - r], win[r], ab);
                        }
                        const float4 xf = silu4_pre<true>(af), xr_ = silu4_pre<true>(ab);
                        accf[k] = MAXPOOL ? max4(accf[k], xf) : accf[k] + xf;
                        accb[k] = MAXPOOL ? max4(accb[k], xr_) : accb[k] + xr_;
                        if (wb_) {
                            const float4 wv = make_float4(fmaf(Db.x, xr_.x, Df.x * xf.x), fmaf(Db.y, xr_.y, Df.y * xf.y),
                                                          fmaf(Db.z, xr_.z, Df.z * xf.z), fmaf(Db.w, xr_.w, Df.w * xf.w));
                            st4(wb_ + (int64_t)rowtab[p_lo * inner + lt0 + k * IHT + 3] * D + d0, wv);
                        }
                    }
                }
            }
        } else if (live) {
            const bf16* cur = xs + (size_t)(c & 1) * bufrows * D + d0;
#pragma unroll
            for (int k = 0; k < NIT; ++k) {
                if (k < NI) {
                    const int i = ih + k * IH;
                    for (int p = 0; p < np; ++p) {
                        const int lt = p * inner + i;               // buffer row lt + 3 is this token; window rows lt .. lt + 6
                        const bf16* xp = cur + (size_t)lt * D;
                        float4 af = tf.b, ab = tb.b;
#pragma unroll
                        for (int r = 0; r < 7; ++r) {
                            const float4 xr = ld4(xp + (size_t)r * D);
                            if (r <= 3) af = fma4(tf.w[r], xr, af);          // causal: x[t-3+r]
                            if (r >= 3) ab = fma4(tb.w[6 - r], xr, ab);      // anti-causal: x[t+3-k], k = 6 - r
                        }
                        const float4 xf = silu4_pre<true>(af), xr_ = silu4_pre<true>(ab);
                        accf[k] = MAXPOOL ? max4(accf[k], xf) : accf[k] + xf;
                        accb[k] = MAXPOOL ? max4(accb[k], xr_) : accb[k] + xr_;
                        if (wb_) {
                            const float4 wv = make_float4(fmaf(Db.x, xr_.x, Df.x * xf.x), fmaf(Db.y, xr_.y, Df.y * xf.y),
                                                          fmaf(Db.z, xr_.z, Df.z * xf.z), fmaf(Db.w, xr_.w, Df.w * xf.w));
                            st4(wb_ + (int64_t)rowtab[p_lo * inner + lt + 3] * D + d0, wv);
                        }
                    }
                }
            }
        }
        __syncthreads();  // all reads of this buffer done before it is refilled two chunks later
    }
    if (!live) return;
    const float m = MAXPOOL ? 1.f : scale / (float)pool;
    const int64_t plane = (int64_t)g.B * g.Lp * D;
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
        if (k < NI) {
            const int i = ih + k * IH;
            bf16* uo = u + ((int64_t)b * g.Lp + (int64_t)o * inner + i) * D + d0;
            st4(uo, MAXPOOL ? accf[k] : scale4(accf[k], m));
            st4(uo + plane, MAXPOOL ? accb[k] : scale4(accb[k], m));
        }
    }
}

// CTA = one (image, outer) group; the group's `inner` pooled rows (s_f + s_b) / 2 are staged in shared memory once (every
// row is shared by the `pool` tokens of its pooled position: read from L2 once per group instead of once per token, which
// would double the kernel's SM ingest), then one warp per token; NV 4-channel groups per lane (D <= 128 * NV)
template <int NV, bool NORM>
__global__ void __launch_bounds__(256)
gate_w_fwd_kernel(Geom g, const bf16* __restrict__ w, const bf16* __restrict__ z, int64_t ldz, int64_t zbs,
                  const float* __restrict__ s, const float* __restrict__ lnw, const float* __restrict__ lnb, float eps,
                  bf16* __restrict__ y, int64_t ldy, int64_t ybs) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ __align__(16) float gw_ssum[];   // [inner][D]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int o = blockIdx.x, b = blockIdx.y;
    const int D = g.D, nvec = D >> 2, inner = g.inner;
    {
        const float* sf = s + ((int64_t)b * g.Lp + (int64_t)o * inner) * D;
        const float* sb = sf + (int64_t)g.B * g.Lp * D;
        // four iterations' loads in flight before the first add / store (a load -> add -> store loop pays one L2 round trip
        // per iteration: 16 % of this kernel's stall samples sat on the add behind the load)
        constexpr int NB = 4;
        for (int i0 = threadIdx.x; i0 < inner * nvec; i0 += blockDim.x * NB) {
            float4 p[NB], q[NB];
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                const int i = i0 + k * blockDim.x;
                p[k] = q[k] = zero4();
                if (i < inner * nvec) {
                    p[k] = __ldg(reinterpret_cast<const float4*>(sf + i * 4));
                    q[k] = __ldg(reinterpret_cast<const float4*>(sb + i * 4));
                }
            }
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                const int i = i0 + k * blockDim.x;
                if (i < inner * nvec)
                    st4(gw_ssum + i * 4, make_float4(0.5f * (p[k].x + q[k].x), 0.5f * (p[k].y + q[k].y), 0.5f * (p[k].z + q[k].z),
                                                     0.5f * (p[k].w + q[k].w)));
            }
        }
    }
    __syncthreads();
    // a group may be split over gridDim.z CTAs (few images, long groups: 2048^2 has 128 groups of 128 tokens)
    const int ntok_all = g.pool * inner, per = (ntok_all + (int)gridDim.z - 1) / (int)gridDim.z;
    const int lt0 = (int)blockIdx.z * per, ntok = min(ntok_all, lt0 + per), tbase = o * ntok_all;
    const bf16* wimg = w + (int64_t)b * g.L * D;
    const bf16* zimg = z + (int64_t)b * zbs;
    bf16* yimg = y + (int64_t)b * ybs;
    // the next token's w and z rows are fetched (raw bf16) while the current one is reduced and stored: two tokens in flight
    // per warp, so 16 resident warps keep ~100 KB per SM outstanding
    uint2 wn[NV], zn[NV];
    int64_t row = 0;
    auto fetch = [&](int lt_) {
        row = seq_to_row(g, tbase + lt_);
        const bf16* wrow = wimg + row * D;
        const bf16* zrow = zimg + row * ldz;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c = lane + k * 32;
            wn[k] = make_uint2(0u, 0u);
            zn[k] = make_uint2(0u, 0u);
            if (c < nvec) {
                wn[k] = __ldg(reinterpret_cast<const uint2*>(wrow + c * 4));
                zn[k] = __ldg(reinterpret_cast<const uint2*>(zrow + c * 4));
            }
        }
    };
    if (lt0 + warp < ntok) fetch(lt0 + warp);
    for (int lt = lt0 + warp; lt < ntok; lt += nwarp) {
        const int i = lt % inner;
        const float* srow = gw_ssum + i * D;
        const int64_t row_cur = row;
        float4 v[NV];
        uint2 zc[NV];
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c = lane + k * 32;
            v[k] = zero4();
            zc[k] = zn[k];
            if (c < nvec) {
                const float2 a01 = unpack2(wn[k].x), a23 = unpack2(wn[k].y);
                const float4 p = ld4(srow + c * 4);
                v[k] = make_float4(a01.x + p.x, a01.y + p.y, a23.x + p.z, a23.y + p.w);
                sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
            }
        }
        if (lt + nwarp < ntok) fetch(lt + nwarp);
        float mean = 0.f, rstd = 1.f;
        if (NORM) {
            mean = warp_sum(sum) / (float)D;
            float sq = 0.f;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                if (lane + k * 32 < nvec) {
                    const float dx = v[k].x - mean, dy = v[k].y - mean, dz = v[k].z - mean, dw = v[k].w - mean;
                    sq += fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, dw * dw)));
                }
            }
            rstd = rsqrtf(warp_sum(sq) / (float)D + eps);
        }
        bf16* yrow = yimg + row_cur * ldy;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c = lane + k * 32;
            if (c < nvec) {
                float4 a = v[k];
                if (NORM) {
                    const float4 gm = ld4(lnw + c * 4), bb = lnb ? ld4(lnb + c * 4) : zero4();
                    a.x = fmaf((a.x - mean) * rstd, gm.x, bb.x); a.y = fmaf((a.y - mean) * rstd, gm.y, bb.y);
                    a.z = fmaf((a.z - mean) * rstd, gm.z, bb.z); a.w = fmaf((a.w - mean) * rstd, gm.w, bb.w);
                }
                const float2 z01 = unpack2(zc[k].x), z23 = unpack2(zc[k].y);
                float4 h = make_float4(0.5f * z01.x, 0.5f * z01.y, 0.5f * z23.x, 0.5f * z23.y);
                h = silu4_pre<true>(h);
                st4(yrow + c * 4, make_float4(a.x * h.x, a.y * h.y, a.z * h.z, a.w * h.w));
            }
        }
    }
}

struct GroupPlan {
    int ok, IH, TP, threads;
    size_t smem;
};
static GroupPlan plan_group(const fv_geom* g, int dtype) {
    GroupPlan p;
    p.ok = 0;
    if (dtype != FV_BF16 || g->inner < 2 || g->dim % 8 != 0) return p;
    const int ncg = g->dim / 4;
    if (ncg > 512) return p;
    // (measured at 32 x 1568 x 768: 384 threads / 126 registers 98.7 us; 768 threads capped at 80 registers, spilling: 116 us)
    // IH = largest divisor of inner with IH * ncg <= 512 threads (rounded up to a warp) and inner / IH <= CG_MAX_NI
    int IH = 0;
    for (int d = g->inner; d >= 1; --d)
        if (g->inner % d == 0 && (int64_t)d * ncg <= 512 && g->inner / d <= CG_MAX_NI) {
            IH = d;
            break;
        }
    if (!IH) return p;
    p.IH = IH;
    p.threads = (IH * ncg + 31) / 32 * 32;
    if (p.threads < ncg) return p;   // stage_rows needs one thread per 4 channels
    const int64_t ntok = (int64_t)g->pool * g->inner;
    if (ntok + 6 > 65536) return p;
    int TP = g->pool < 4 ? g->pool : 4;
    for (; TP >= 1; TP >>= 1) {
        const size_t smem = (size_t)2 * (TP * g->inner + 6) * g->dim * 2 + (size_t)(ntok + 6) * 4;
        if (smem <= 72 * 1024 || TP == 1) {
            p.TP = TP;
            p.smem = smem;
            break;
        }
    }
    if (p.smem > 200 * 1024) return p;
    p.ok = 1;
    return p;
}

}  // namespace fv

extern "C" int fv_conv_pool_w_supported(const fv_geom* g, int dtype) {
    if (!g || g->batch <= 0 || g->dim <= 0 || g->outer <= 0 || g->pool <= 0 || g->inner <= 0 || g->batch > 65535) return 0;
    if (dtype == FV_BF16 && g->inner == 1) return g->dim % 8 == 0 && g->dim <= 3072;   // staged / cluster kernels of conv_pool.cu
    return fv::plan_group(g, dtype).ok;
}

extern "C" int fv_conv_pool_w_fwd(const fv_geom* g_, int dtype, const void* x, int64_t ldx, int64_t x_bstride,
                                  const float* conv_w, const float* conv_b, float scale, int pool_mode, const float* Dskip,
                                  void* u_out, void* w_out, void* stream) {
    using namespace fv;
    if (int rc = check_geom(g_, "fv_conv_pool_w_fwd")) return rc;
    FV_REQUIRE(x && conv_w && u_out && (!w_out || Dskip), "fv_conv_pool_w_fwd: null pointer (w_out needs Dskip)");
    FV_REQUIRE(fv_conv_pool_w_supported(g_, dtype), "fv_conv_pool_w_fwd: unsupported configuration (bf16, dim %% 8 == 0)");
    if (g_->inner == 1) {
        FV_REQUIRE(ldx % 8 == 0 && x_bstride % 8 == 0 && ((uintptr_t)x % 16) == 0, "fv_conv_pool_w_fwd: x rows must be 16-byte aligned");
        return conv_pool_plain_w(make_geom(g_), (const bf16*)x, ldx, x_bstride, conv_w, conv_b, scale, pool_mode, (bf16*)u_out,
                                 Dskip, (bf16*)w_out, (cudaStream_t)stream);
    }
    const GroupPlan p = plan_group(g_, dtype);
    FV_REQUIRE(p.ok && g_->batch <= 65535, "fv_conv_pool_w_fwd: unsupported configuration (bf16, inner >= 2, dim %% 8 == 0, dim <= 2048)");
    FV_REQUIRE(ldx % 8 == 0 && x_bstride % 8 == 0 && ((uintptr_t)x % 16) == 0, "fv_conv_pool_w_fwd: x rows must be 16-byte aligned");
    FV_REQUIRE(((uintptr_t)u_out % 8) == 0 && (!w_out || ((uintptr_t)w_out % 8) == 0), "fv_conv_pool_w_fwd: misaligned output");
    Geom g = make_geom(g_);
    const int NI = g.inner / p.IH;
    const bool mx = pool_mode == FV_POOL_MAX;
    void (*kern)(Geom, int, int, const bf16*, int64_t, int64_t, const float*, const float*, float, const float*, bf16*, bf16*);
    const bool slide = p.threads <= 384 && (getenv("FASTVIM_CONV_GROUP_SLIDE") == nullptr || getenv("FASTVIM_CONV_GROUP_SLIDE")[0] != '0');
#define FV_CG_PICK(NI_, IH_) (mx ? conv_pool_group_kernel<true, NI_, IH_> : conv_pool_group_kernel<false, NI_, IH_>)
    if (NI <= 1) kern = FV_CG_PICK(1, 0);                                   // one slot per thread: nothing to slide over
    else if (NI <= 2) kern = slide && p.IH == 1 ? FV_CG_PICK(2, 1) : slide && p.IH == 2 ? FV_CG_PICK(2, 2)
                             : slide && p.IH == 4 ? FV_CG_PICK(2, 4) : FV_CG_PICK(2, 0);
    else kern = slide && p.IH == 1 ? FV_CG_PICK(4, 1) : slide && p.IH == 2 ? FV_CG_PICK(4, 2) : FV_CG_PICK(4, 0);
#undef FV_CG_PICK
    if (p.smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
        FV_REQUIRE(e == cudaSuccess, "fv_conv_pool_w_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    dim3 grid(g.outer, g.B), block(p.threads);
    FV_LAUNCH_PDL((kern), grid, block, p.smem, stream, g, p.IH, p.TP, (const bf16*)x, ldx, x_bstride, conv_w, conv_b, scale, Dskip,
                  (bf16*)u_out, (bf16*)w_out);
    return finish_launch("conv_pool_w_fwd");
}

extern "C" int fv_gate_w_fwd(const fv_geom* g_, int dtype, const void* w, const void* z, int64_t ldz, int64_t z_bstride,
                             const float* s, const float* ln_w, const float* ln_b, float eps, void* y, int64_t ldy,
                             int64_t y_bstride, void* stream) {
    using namespace fv;
    if (int rc = check_geom(g_, "fv_gate_w_fwd")) return rc;
    FV_REQUIRE(w && z && s && y, "fv_gate_w_fwd: null pointer");
    FV_REQUIRE(dtype == FV_BF16, "fv_gate_w_fwd: bf16 only");
    FV_REQUIRE(g_->dim % 4 == 0 && g_->dim <= 128 * 24, "fv_gate_w_fwd: dim (%d) must be a multiple of 4, <= 3072", g_->dim);
    FV_REQUIRE(ldz % 4 == 0 && z_bstride % 4 == 0 && ldy % 4 == 0 && y_bstride % 4 == 0 && ((uintptr_t)w % 8) == 0 &&
                   ((uintptr_t)z % 8) == 0 && ((uintptr_t)y % 8) == 0 && ((uintptr_t)s % 16) == 0,
               "fv_gate_w_fwd: rows must be 8-byte aligned");
    Geom g = make_geom(g_);
    FV_REQUIRE(g.B <= 65535, "fv_gate_w_fwd: batch > 65535");
    const size_t smem = (size_t)g.inner * g.D * 4;
    FV_REQUIRE(smem <= 200 * 1024, "fv_gate_w_fwd: inner * dim too large for the staged pooled rows");
    const int ntok = g.pool * g.inner;
    const int warps = ntok >= 16 ? 8 : 4;
    int nsplit = (int)((4ll * sm_count() + (int64_t)g.outer * g.B - 1) / ((int64_t)g.outer * g.B));
    if (nsplit > ntok / (2 * warps)) nsplit = ntok / (2 * warps);   // at least two tokens per warp
    if (nsplit < 1) nsplit = 1;
    if (nsplit > 64) nsplit = 64;
    dim3 grid(g.outer, g.B, nsplit);
    const int nv = (g.D / 4 + 31) / 32;
#define FV_GW2(NV_, NORM_)                                                                                             \
    {                                                                                                                  \
        auto kern = gate_w_fwd_kernel<NV_, NORM_>;                                                                     \
        if (smem > 48 * 1024) {                                                                                        \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
            FV_REQUIRE(e == cudaSuccess, "fv_gate_w_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));            \
        }                                                                                                              \
        FV_LAUNCH_PDL((kern), grid, warps * 32, smem, stream, g, (const bf16*)w, (const bf16*)z, ldz, z_bstride, s, ln_w, \
                      ln_b, eps, (bf16*)y, ldy, y_bstride);                                                            \
    }
#define FV_GW(NV_)                \
    {                             \
        if (ln_w) FV_GW2(NV_, true) \
        else FV_GW2(NV_, false)   \
    }
    if (nv <= 1) FV_GW(1)
    else if (nv <= 2) FV_GW(2)
    else if (nv <= 3) FV_GW(3)
    else if (nv <= 6) FV_GW(6)
    else if (nv <= 12) FV_GW(12)
    else FV_GW(24)
#undef FV_GW
#undef FV_GW2
    return finish_launch("gate_w_fwd");
}
