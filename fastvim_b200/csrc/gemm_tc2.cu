// General tcgen05 / TMEM / TMA GEMM: the backward GEMMs of the projections (dgrad, split-K wgrad) and ragged shapes.
//
//   C (Mo x No) = op(A) . op(B)^T over a reduction of length K, bf16 operands, fp32 accumulation in tensor memory.
//     a_mn = 0: A is stored (Mo x K) row-major -- the reduction index is contiguous ("K-major" UMMA operand)
//     a_mn = 1: A is stored (K x Mo) row-major -- the output-row index is contiguous ("MN-major" UMMA operand)
//     b_mn likewise for B: (No x K) or (K x No).
//
// What it replaces in the reference (paths relative to /root/reference): the cuBLAS calls of the fused backward,
// mamba_ssm/ops/selective_scan_interface.py:698-737 (and autograd through in_proj / out_proj of mamba_simple_faster.py):
//   dgrad   dX (M x K')  = dY (M x N') . W (N' x K')              a_mn = 0 (dY), b_mn = 1 (W as stored)   bf16 out
//   wgrad   dW (N' x K') = dY (M x N')^T . X (M x K')             a_mn = 1 (dY), b_mn = 1 (X)             fp32 out, split-K over M
//   x_proj  xdbl (M x 44) = u (M x D) . W_x (44 x D)^T            a_mn = 0, b_mn = 0, ragged N            bf16 out
// No operand is transposed in memory: an MN-major operand tile is fetched as (64 k-rows x 64 columns) TMA boxes with the
// 128-byte swizzle, which is exactly the canonical MN-major SWIZZLE_128B layout of the UMMA shared-memory descriptor
// (8 k-rows x 128 B atoms: stride between 8-row k-groups = 1024 B, between 64-column repeats = 8192 B); the instruction
// descriptor's a_major / b_major bits select the transposed read.  Out-of-range rows / columns / k are zero-filled by TMA on
// the way in and clipped by TMA on the way out, so Mo, No, K need no padding (only 16-byte row pitches).
//
// One persistent CTA per SM: warp 0 = TMA producer (ring of (A 16 KB + B BN x 128 B) stages), warp 1 = MMA issuer (one
// thread, tcgen05.mma cta_group::1 kind::f16, 128 x BN x 16), warps 2-5 = epilogue (tcgen05.ld -> swizzled smem tile -> TMA
// store), accumulators double-buffered in TMEM.  Split-K: tile = (split, row tile, column block); split s reduces its own
// range of k-blocks into plane s of the fp32 output (deterministic; planes are added by fv_reduce_planes).
#include "tc_common.cuh"

namespace fv {

int sm_count();

constexpr int G2_BM = 128, G2_BK = 64, G2_THREADS = 192;
constexpr uint32_t G2_A_BYTES = G2_BM * G2_BK * 2;  // 16 KB
constexpr uint32_t G2_C_BYTES = 128 * 128;          // staged store: 128 rows x 128 B (64 bf16 or 32 fp32 columns)

struct Gemm2Args {
    int Mo, No, KB, kb_per_split, ntm, ntn, ntiles, nstage, ncstage;
    int nbatch, splits;   // tile = (batch, split, row tile, column block); operands and C carry `nbatch` planes
    int red;   // fp32 output: every split adds its partial sum into the ONE output plane (TMA reduction store) instead of writing its own
};

// MN-major operand tile, 128-byte swizzle (cute::UMMA canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units):
// leading_byte_offset = stride between 64-element column repeats (8192 B), stride_byte_offset = stride between 8-row
// k-groups (1024 B), version = 1, layout_type = SWIZZLE_128B
__device__ __forceinline__ uint64_t tc_smem_desc_mn(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)512 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}

template <int BN, bool A_MN, bool B_MN, bool OUT_F32>
__global__ void __launch_bounds__(G2_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const Gemm2Args a) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    extern __shared__ unsigned char g2_smem_raw[];
    const uint32_t raw = smem_u32(g2_smem_raw);
    unsigned char* smem = g2_smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NS = a.nstage, NC = a.ncstage;
    constexpr uint32_t B_BYTES = BN * G2_BK * 2;
    constexpr uint32_t STAGE = G2_A_BYTES + B_BYTES;

    unsigned char* sAB = smem;
    unsigned char* sC = sAB + (size_t)NS * STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sC + (size_t)NC * G2_C_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = full + NS;
    uint64_t* acc_full = empty + NS;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);

    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 128);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int per_split = a.ntm * a.ntn;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int st = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
                const int bs = tile / per_split, batch = bs / a.splits, split = bs - batch * a.splits, rem = tile - bs * per_split;
                const int m0 = (rem / a.ntn) * G2_BM, n0 = (rem % a.ntn) * BN;
                const int kb0 = split * a.kb_per_split, kb1 = min(a.KB, kb0 + a.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    gt_mbar_wait(&empty[st], ph ^ 1u);
                    unsigned char* stg = sAB + (size_t)st * STAGE;
                    mbar_arrive_expect_tx(&full[st], STAGE);
                    const int k0 = kb * G2_BK;
                    if (A_MN) {
#pragma unroll
                        for (int i = 0; i < G2_BM / 64; ++i) gt_tma_3d(stg + i * 8192, &tmA, m0 + 64 * i, k0, batch, &full[st]);
                    } else {
                        gt_tma_3d(stg, &tmA, k0, m0, batch, &full[st]);
                    }
                    if (B_MN) {
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j)
                            gt_tma_3d(stg + G2_A_BYTES + j * 8192, &tmB, n0 + 64 * j, k0, batch, &full[st]);
                    } else {
                        gt_tma_3d(stg + G2_A_BYTES, &tmB, k0, n0, batch, &full[st]);
                    }
                    if (++st == NS) { st = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 (bits 4-5 = 1), A / B bf16 (bits 7-9, 10-12 = 1),
            // a_major bit 15, b_major bit 16 (1 = MN-major), N >> 3 at bit 17, M >> 4 at bit 24
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                       ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(G2_BM >> 4) << 24);
            int st = 0, as = 0;
            uint32_t ph = 0, aph = 0;
            const uint32_t s_u = smem_u32(sAB);
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
                const int split = (tile / per_split) % a.splits;
                const int kb0 = split * a.kb_per_split, kb1 = min(a.KB, kb0 + a.kb_per_split);
                gt_mbar_wait(&acc_empty[as], aph ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
                for (int kb = kb0; kb < kb1; ++kb) {
                    gt_mbar_wait(&full[st], ph);
                    tc_fence_after();
                    const uint32_t a_u = s_u + (uint32_t)st * STAGE, b_u = a_u + G2_A_BYTES;
#pragma unroll
                    for (int k = 0; k < G2_BK / 16; ++k) {
                        const uint64_t ad = A_MN ? tc_smem_desc_mn(a_u + k * 2048) : tc_smem_desc(a_u + k * 32);
                        const uint64_t bd = B_MN ? tc_smem_desc_mn(b_u + k * 2048) : tc_smem_desc(b_u + k * 32);
                        tc_mma(d_tmem, ad, bd, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
                    }
                    tc_commit(&empty[st]);
                    if (++st == NS) { st = 0; ph ^= 1u; }
                }
                tc_commit(&acc_full[as]);
                as ^= 1;
                if (as == 0) aph ^= 1u;
            }
        }
    } else {
        // ================= epilogue: TMEM lane = output row, one row per thread =================
        const int q = warp & 3;
        const int trow = q * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const bool issuer = warp == 2 && lane == 0;
        constexpr int CH = OUT_F32 ? 32 : 64;  // columns per staged 128-byte row
        int as = 0, cs = 0;
        uint32_t aph = 0;
        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
            const int bs = tile / per_split, rem = tile - bs * per_split;   // bs = batch * splits + split = output plane
            const int m0 = (rem / a.ntn) * G2_BM, n0 = (rem % a.ntn) * BN;
            gt_mbar_wait(&acc_full[as], aph);
            tc_fence_after();
            const uint32_t t0 = lane_addr + (uint32_t)(as * BN);
#pragma unroll 1
            for (int c = 0; c < BN / CH; ++c) {
                if (n0 + c * CH >= a.No) break;  // tile-uniform: the remaining chunks lie beyond the last column
                unsigned char* stage = sC + (size_t)cs * G2_C_BYTES;
                if (issuer) {
                    if (NC == 2) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(1) : "memory");
                    else asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(0) : "memory");
                }
                gt_epi_barrier();
                unsigned char* srow = stage + (size_t)trow * 128;
                if (OUT_F32) {
                    uint32_t r[32];
                    tc_ld32(t0 + c * 32, r);
                    tc_wait_ld();
#pragma unroll
                    for (int v = 0; v < 8; ++v) {
                        const int chunk16 = v ^ (trow & 7);
                        *reinterpret_cast<uint4*>(srow + chunk16 * 16) = make_uint4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
                    }
                } else {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t r[32];
                        tc_ld32(t0 + c * 64 + h * 32, r);
                        tc_wait_ld();
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            uint4 o;
                            o.x = gt_pack(r[8 * v], r[8 * v + 1]); o.y = gt_pack(r[8 * v + 2], r[8 * v + 3]);
                            o.z = gt_pack(r[8 * v + 4], r[8 * v + 5]); o.w = gt_pack(r[8 * v + 6], r[8 * v + 7]);
                            const int chunk16 = (h * 4 + v) ^ (trow & 7);
                            *reinterpret_cast<uint4*>(srow + chunk16 * 16) = o;
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                gt_epi_barrier();
                if (issuer) {
                    if (OUT_F32 && a.red) gt_tma_red_add_3d(&tmC, stage, n0 + c * CH, m0, bs / a.splits);
                    else gt_tma_store_3d(&tmC, stage, n0 + c * CH, m0, bs);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (++cs == NC) cs = 0;
            }
            tc_fence_before();
            gt_mbar_arrive(&acc_empty[as]);
            as ^= 1;
            if (as == 0) aph ^= 1u;
        }
        if (issuer) asm volatile("cp.async.bulk.wait_group %0;" ::"n"(0) : "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- host ------------------------------------------------------------------------------------------------------
struct Tm2Key {
    const void* base;
    int64_t d0, d1, d2, ld, ps;
    int b0, b1, es;
    bool operator==(const Tm2Key& o) const {
        return base == o.base && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && ld == o.ld && ps == o.ps && b0 == o.b0 && b1 == o.b1 &&
               es == o.es;
    }
};
// Tensor map over a row-major matrix (d1 rows x d0 columns, row pitch ld elements of es bytes) with d2 planes of d1 * ld
// elements; box = (b0 columns x b1 rows x 1 plane), 128-byte swizzle, out-of-range elements read as zero / are not written.
int get_tmap(CUtensorMap* out, const void* base, int es, int64_t d0, int64_t d1, int64_t d2, int64_t ld, int b0, int b1,
             int64_t plane_stride) {
    if (plane_stride <= 0) plane_stride = d1 * ld;
    constexpr int NCACHE = 192;
    static thread_local Tm2Key keys[NCACHE];
    static thread_local CUtensorMap maps[NCACHE];
    static thread_local int used = 0, next = 0;
    const Tm2Key k{base, d0, d1, d2, ld, plane_stride, b0, b1, es};
    for (int i = 0; i < used; ++i)
        if (keys[i] == k) {
            *out = maps[i];
            return 0;
        }
    TmEncodeFn encode = tm_encode_fn();
    if (!encode) return 1;
    const bool three = d2 > 0;
    cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)(three ? d2 : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)(ld * es), (cuuint64_t)(plane_stride * es)};
    cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    CUtensorMap m;
    const CUresult r = encode(&m, es == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, three ? 3 : 2,
                              const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("fv_gemm_bf16: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    const int slot = used < NCACHE ? used++ : (next++ % NCACHE);
    keys[slot] = k;
    maps[slot] = m;
    *out = m;
    return 0;
}

static int pick_bn2(int No) {
    const int cands[4] = {256, 192, 128, 64};
    int best = 64;
    int64_t best_cost = -1;
    for (int c : cands) {
        const int64_t cost = (int64_t)((No + c - 1) / c) * c;
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best = c;
        }
    }
    return best;
}

static bool plan_smem2(int BN, int* ns, int* nc, size_t* bytes) {
    const size_t stage = G2_A_BYTES + (size_t)BN * G2_BK * 2, fixed = 32 * 8 + 16 + 1024, cap = 227 * 1024;
    const int opts[6][2] = {{6, 2}, {5, 2}, {4, 2}, {4, 1}, {3, 2}, {3, 1}};
    for (auto& o : opts) {
        const size_t tot = (size_t)o[0] * stage + (size_t)o[1] * G2_C_BYTES + fixed;
        if (tot <= cap) {
            *ns = o[0]; *nc = o[1]; *bytes = tot;
            return true;
        }
    }
    return false;
}

template <int BN, bool A_MN, bool B_MN, bool OUT_F32>
static int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const Gemm2Args& a, size_t smem,
                        cudaStream_t st) {
    auto kern = gemm_tc2_kernel<BN, A_MN, B_MN, OUT_F32>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    FV_REQUIRE(e == cudaSuccess, "fv_gemm_bf16: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    const int grid = a.ntiles < sm_count() ? a.ntiles : sm_count();
    FV_LAUNCH_PDL((kern), grid, G2_THREADS, smem, st, tmA, tmB, tmC, a);
    return finish_launch("gemm_tc2");
}

template <int BN>
static int dispatch_gemm2(int a_mn, int b_mn, int out_f32, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                          const Gemm2Args& a, size_t smem, cudaStream_t st) {
#define FV_G2(A_, B_, O_) \
    if (a_mn == A_ && b_mn == B_ && out_f32 == O_) return launch_gemm2<BN, A_ != 0, B_ != 0, O_ != 0>(tmA, tmB, tmC, a, smem, st);
    FV_G2(0, 0, 0) FV_G2(0, 0, 1) FV_G2(0, 1, 0) FV_G2(0, 1, 1) FV_G2(1, 0, 0) FV_G2(1, 0, 1) FV_G2(1, 1, 0) FV_G2(1, 1, 1)
#undef FV_G2
    return fail("fv_gemm_bf16: internal error");
}

}  // namespace fv

extern "C" int fv_gemm_bf16_splits(int64_t Mo, int No, int64_t K) {
    using namespace fv;
    if (Mo <= 0 || No <= 0 || K <= 0) return 1;
    const int BN = pick_bn2(No);
    const int64_t tiles = ((Mo + G2_BM - 1) / G2_BM) * ((No + BN - 1) / BN);
    const int64_t KB = (K + G2_BK - 1) / G2_BK;
    int64_t s = sm_count() / tiles;
    if (s > KB / 8) s = KB / 8;  // at least 8 k-blocks per split
    if (s < 1) s = 1;
    if (s > 64) s = 64;
    // no empty split: ceil(KB / s) * (s - 1) < KB
    while (s > 1 && ((KB + s - 1) / s) * (s - 1) >= KB) --s;
    return (int)s;
}

namespace fv {
static int gemm2_impl(int nbatch, int64_t Mo, int No, int64_t K, int a_mn, const void* A, int64_t lda, int64_t a_bs, int b_mn,
                      const void* B, int64_t ldb, int64_t b_bs, int out_dtype, void* C, int64_t ldc, int64_t c_bs, int splits,
                      void* stream);
}
extern "C" int fv_gemm_bf16(int64_t Mo, int No, int64_t K, int a_mn, const void* A, int64_t lda, int b_mn, const void* B,
                            int64_t ldb, int out_dtype, void* C, int64_t ldc, int splits, void* stream) {
    return fv::gemm2_impl(1, Mo, No, K, a_mn, A, lda, 0, b_mn, B, ldb, 0, out_dtype, C, ldc, 0, splits, stream);
}
// `nbatch` independent products in one launch: operand / result b starts a_bs / b_bs / c_bs ELEMENTS after b - 1 (multiples of
// 8); no split-K.  Used for the two directions of x_proj (mamba_simple_faster.py:321-323, 377-379).
extern "C" int fv_gemm_bf16_batched(int nbatch, int64_t Mo, int No, int64_t K, int a_mn, const void* A, int64_t lda, int64_t a_bs,
                                    int b_mn, const void* B, int64_t ldb, int64_t b_bs, int out_dtype, void* C, int64_t ldc,
                                    int64_t c_bs, int splits, void* stream) {
    FV_REQUIRE(nbatch >= 1 && nbatch <= 65535 && a_bs % 8 == 0 && b_bs % 8 == 0 && a_bs > 0 && b_bs > 0 && c_bs > 0 &&
                   (c_bs * (out_dtype == FV_BF16 ? 2 : 4)) % 16 == 0,
               "fv_gemm_bf16_batched: batch count / strides (16-byte multiples) invalid");
    FV_REQUIRE(out_dtype == FV_BF16 || out_dtype == FV_F32 || out_dtype == FV_F32_ACC,
               "fv_gemm_bf16_batched: out_dtype must be FV_BF16, FV_F32 or FV_F32_ACC");
    FV_REQUIRE(splits == 1 || out_dtype == FV_F32_ACC, "fv_gemm_bf16_batched: split-K only with FV_F32_ACC (one plane per batch)");
    return fv::gemm2_impl(nbatch, Mo, No, K, a_mn, A, lda, a_bs, b_mn, B, ldb, b_bs, out_dtype, C, ldc, c_bs, splits, stream);
}
namespace fv {
static int gemm2_impl(int nbatch, int64_t Mo, int No, int64_t K, int a_mn, const void* A, int64_t lda, int64_t a_bs, int b_mn,
                      const void* B, int64_t ldb, int64_t b_bs, int out_dtype, void* C, int64_t ldc, int64_t c_bs, int splits,
                      void* stream) {
    FV_REQUIRE(A && B && C, "fv_gemm_bf16: null pointer");
    FV_REQUIRE(Mo > 0 && No > 0 && K > 0 && Mo < (1ll << 31) && K < (1ll << 31), "fv_gemm_bf16: bad sizes (%lld x %d x %lld)",
               (long long)Mo, No, (long long)K);
    FV_REQUIRE(out_dtype == FV_BF16 || out_dtype == FV_F32 || out_dtype == FV_F32_ACC,
               "fv_gemm_bf16: out_dtype must be FV_BF16, FV_F32 or FV_F32_ACC");
    const bool red = out_dtype == FV_F32_ACC;
    if (red) out_dtype = FV_F32;
    const int es_c = out_dtype == FV_F32 ? 4 : 2;
    FV_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && (ldc * es_c) % 16 == 0 && ((uintptr_t)A % 16) == 0 && ((uintptr_t)B % 16) == 0 &&
                   ((uintptr_t)C % 16) == 0,
               "fv_gemm_bf16: operands must be 16-byte aligned with 16-byte row pitches (lda %lld, ldb %lld, ldc %lld)",
               (long long)lda, (long long)ldb, (long long)ldc);
    FV_REQUIRE(lda >= (a_mn ? Mo : K) && ldb >= (b_mn ? No : K) && ldc >= No, "fv_gemm_bf16: row pitch smaller than the row");
    FV_REQUIRE(splits >= 1 && (splits == 1 || out_dtype == FV_F32), "fv_gemm_bf16: split-K needs fp32 output planes");
    const int BN = pick_bn2(No);
    Gemm2Args a;
    a.Mo = (int)Mo; a.No = No; a.red = red ? 1 : 0; a.nbatch = nbatch; a.splits = splits;
    a.KB = (int)((K + G2_BK - 1) / G2_BK);
    FV_REQUIRE(splits <= a.KB, "fv_gemm_bf16: more splits (%d) than k-blocks (%d)", splits, a.KB);
    a.kb_per_split = (a.KB + splits - 1) / splits;
    FV_REQUIRE((int64_t)a.kb_per_split * (splits - 1) < a.KB, "fv_gemm_bf16: split count %d leaves an empty split (use fv_gemm_bf16_splits)", splits);
    a.ntm = (int)((Mo + G2_BM - 1) / G2_BM);
    a.ntn = (No + BN - 1) / BN;
    const int64_t nt = (int64_t)a.ntm * a.ntn * splits * nbatch;
    FV_REQUIRE(nt < (1ll << 31), "fv_gemm_bf16: too many tiles");
    a.ntiles = (int)nt;
    size_t smem = 0;
    FV_REQUIRE(plan_smem2(BN, &a.nstage, &a.ncstage, &smem), "fv_gemm_bf16: no shared-memory plan for BN = %d", BN);
    CUtensorMap tmA, tmB, tmC;
    if (a_mn) {
        if (int rc = get_tmap(&tmA, A, 2, Mo, K, nbatch, lda, 64, 64, a_bs)) return rc;
    } else {
        if (int rc = get_tmap(&tmA, A, 2, K, Mo, nbatch, lda, 64, G2_BM, a_bs)) return rc;
    }
    if (b_mn) {
        if (int rc = get_tmap(&tmB, B, 2, No, K, nbatch, ldb, 64, 64, b_bs)) return rc;
    } else {
        if (int rc = get_tmap(&tmB, B, 2, K, No, nbatch, ldb, 64, BN, b_bs)) return rc;
    }
    if (int rc = get_tmap(&tmC, C, es_c, No, Mo, nbatch * (red ? 1 : splits), ldc, out_dtype == FV_F32 ? 32 : 64, G2_BM,
                          nbatch > 1 ? c_bs : 0))
        return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int of = out_dtype == FV_F32 ? 1 : 0;
    if (BN == 256) return dispatch_gemm2<256>(a_mn ? 1 : 0, b_mn ? 1 : 0, of, tmA, tmB, tmC, a, smem, st);
    if (BN == 192) return dispatch_gemm2<192>(a_mn ? 1 : 0, b_mn ? 1 : 0, of, tmA, tmB, tmC, a, smem, st);
    if (BN == 128) return dispatch_gemm2<128>(a_mn ? 1 : 0, b_mn ? 1 : 0, of, tmA, tmB, tmC, a, smem, st);
    return dispatch_gemm2<64>(a_mn ? 1 : 0, b_mn ? 1 : 0, of, tmA, tmB, tmC, a, smem, st);
}
}  // namespace fv
