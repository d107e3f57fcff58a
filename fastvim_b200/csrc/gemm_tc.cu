// tcgen05 / TMEM / TMA GEMM for the block's projections:  C (M x N) = A (M x K) . W (N x K)^T, bf16 in, fp32 accumulate.
//
// What it replaces in the reference (paths relative to /root/reference): the cuBLAS calls behind
//   in_proj   rearrange(self.in_proj.weight @ rearrange(hidden_states, "b l d -> d (b l)"), ...)   mamba_simple_faster.py:189-195
//   out_proj  F.linear(y, out_proj.weight, out_proj.bias)                                          :442-444
//
// B200 mapping (one CTA per SM, persistent over 128-row tiles of A):
//   * FastVim-T/S: W is small (192 x 384 / 256 x 192 bf16): the CTA's column block of W is loaded ONCE by TMA into shared memory
//     (128-byte swizzle, K-major; k-blocks interleaved with the first tile's A boxes) and stays resident; only A tiles
//     stream from HBM (3-4 stage TMA ring of 128 x 64 boxes).  Large K (FastVim-B, patch embedding: K = 768 / 1536): the
//     weight block does not fit, so W k-blocks stream through the same ring as A (L2-resident after the first tile).
//   * warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread issues tcgen05.mma.cta_group::1.kind::f16,
//     M = 128, N = BN, K = 16 per instruction; accumulators live in TMEM, double-buffered 2 x BN columns so the
//     epilogue of tile i overlaps the MMAs of tile i+1), warps 2-5 = epilogue (tcgen05.ld 32x32b: one TMEM lane =
//     one output row per thread -> bf16 -> swizzled smem tile -> TMA store, so global writes are whole 128-byte lines).
// Synchronisation: mbarriers (TMA complete_tx, tcgen05.commit arrivals) plus a 128-thread named barrier among the
// epilogue warps; no __syncthreads in the main loop.
#include "tc_common.cuh"

namespace fv {

int sm_count();

constexpr int GT_BM = 128, GT_BK = 64, GT_THREADS = 192;
constexpr uint32_t GT_A_STAGE_BYTES = GT_BM * GT_BK * 2;  // 16 KB

struct GemmArgs {
    int M, N, K, KB, nstage, ncstage, ntiles;
    int stream_w;   // 0: the CTA's W block stays resident (small K*BN); 1: W k-blocks stream through the ring with A
    int nblocks;    // column blocks of BN (stream mode: a tile is (row tile, column block), column block fastest)
    const float* bias;  // (N) fp32 added before the bf16 rounding, or null
};

// Epilogue staging: the four epilogue warps convert their TMEM rows to bf16 and write them into a (128 x 64) bf16
// tile in shared memory with the TMA 128-byte swizzle (16-byte chunk index XOR row % 8: conflict-free 16-byte stores
// although every thread owns a different row), then ONE thread hands the tile to the TMA engine
// (cp.async.bulk.tensor store): global writes are full 128-byte lines regardless of the row pitch.
constexpr int GT_CCHUNK = 64;                                  // columns per staged store
constexpr uint32_t GT_C_STAGE_BYTES = GT_BM * GT_CCHUNK * 2;   // 16 KB

template <int BN>
__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmC, const GemmArgs a) {
    extern __shared__ unsigned char gt_smem_raw[];
    const uint32_t raw = smem_u32(gt_smem_raw);
    unsigned char* smem = gt_smem_raw + (((raw + 1023u) & ~1023u) - raw);  // 128B-swizzle tiles need 1024-byte alignment
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KB = a.KB, NS = a.nstage, NC = a.ncstage;
    constexpr uint32_t W_BLK_BYTES = BN * GT_BK * 2;

    // resident mode: [W: KB blocks][A ring: NS x 16 KB][C staging]; stream mode: [ring: NS x (A 16 KB + W block)][C staging]
    const bool stream = a.stream_w != 0;
    const uint32_t stage_bytes = stream ? GT_A_STAGE_BYTES + W_BLK_BYTES : GT_A_STAGE_BYTES;
    unsigned char* sW = smem;
    unsigned char* sA = stream ? smem : sW + (size_t)KB * W_BLK_BYTES;
    unsigned char* sC = sA + (size_t)NS * stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sC + (size_t)NC * GT_C_STAGE_BYTES);
    uint64_t* w_full = bars;            // [KB]: W k-block kb has landed (once per kernel)
    uint64_t* a_full = w_full + KB;
    uint64_t* a_empty = a_full + NS;
    uint64_t* acc_full = a_empty + NS;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);

    if (threadIdx.x == 0) {
        for (int i = 0; i < KB; ++i) mbar_init(&w_full[i], 1);
        for (int i = 0; i < NS; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 128);
        }
        fence_mbar_init();
    }
    if (warp == 1) {  // TMEM: 2 accumulator stages of BN fp32 columns (allocation is a power of two >= 32)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();     // everything above overlapped the previous kernel's tail; A (and C) belong to the stream order from here
    pdl_trigger();
    const int nblk = a.nblocks;
    // resident mode: this CTA's column block is blockIdx.y and a tile is a row tile; stream mode: tile = (row tile, block)
#define GT_TILE_M(tile_) (stream ? (tile_) / nblk : (tile_))
#define GT_TILE_N0(tile_) ((stream ? (tile_) % nblk : (int)blockIdx.y) * BN)

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int st = 0;
            uint32_t ph = 0;
            bool first = true;
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
                const int m0 = GT_TILE_M(tile) * GT_BM, n0 = GT_TILE_N0(tile);
                for (int kb = 0; kb < KB; ++kb) {
                    if (!stream && first) {  // W k-blocks are interleaved with the first tile's A boxes: the first MMAs start early
                        mbar_arrive_expect_tx(&w_full[kb], W_BLK_BYTES);
                        gt_tma_2d(sW + (size_t)kb * W_BLK_BYTES, &tmW, kb * GT_BK, n0, &w_full[kb]);
                    }
                    gt_mbar_wait(&a_empty[st], ph ^ 1u);
                    unsigned char* stg = sA + (size_t)st * stage_bytes;
                    mbar_arrive_expect_tx(&a_full[st], stage_bytes);
                    gt_tma_2d(stg, &tmA, kb * GT_BK, m0, &a_full[st]);
                    if (stream) gt_tma_2d(stg + GT_A_STAGE_BYTES, &tmW, kb * GT_BK, n0, &a_full[st]);
                    if (++st == NS) { st = 0; ph ^= 1u; }
                }
                first = false;
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            // instruction descriptor: D fp32 (bits 4-5 = 1), A/B bf16 (bits 7-9, 10-12 = 1), both K-major, N >> 3 at bit 17,
            // M >> 4 at bit 24 (cute::UMMA::InstrDescriptor)
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(GT_BM >> 4) << 24);
            int st = 0, as = 0;
            uint32_t ph = 0, aph = 0;
            bool first = true;
            const uint32_t sA_u = smem_u32(sA), sW_u = smem_u32(sW);
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
                gt_mbar_wait(&acc_empty[as], aph ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
                for (int kb = 0; kb < KB; ++kb) {
                    if (!stream && first) gt_mbar_wait(&w_full[kb], 0);
                    gt_mbar_wait(&a_full[st], ph);
                    tc_fence_after();
                    const uint32_t a_u = sA_u + (uint32_t)st * stage_bytes;
                    const uint32_t b_u = stream ? a_u + GT_A_STAGE_BYTES : sW_u + (uint32_t)kb * W_BLK_BYTES;
#pragma unroll
                    for (int k = 0; k < GT_BK / 16; ++k) {
                        const uint64_t ad = tc_smem_desc(a_u + k * 32);
                        const uint64_t bd = tc_smem_desc(b_u + k * 32);
                        tc_mma(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    tc_commit(&a_empty[st]);  // frees the A stage once these MMAs have read it
                    if (++st == NS) { st = 0; ph ^= 1u; }
                }
                first = false;
                tc_commit(&acc_full[as]);     // accumulator of this tile complete
                as ^= 1;
                if (as == 0) aph ^= 1u;
            }
        }
    } else {
        // ================= epilogue: TMEM lane = output row, one row per thread =================
        const int q = warp & 3;  // a warp may only touch TMEM lanes [32 * (warp % 4), +32)
        const int trow = q * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const bool issuer = warp == 2 && lane == 0;
        int as = 0, cs = 0;
        uint32_t aph = 0;
        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
            gt_mbar_wait(&acc_full[as], aph);
            tc_fence_after();
            const uint32_t t0 = lane_addr + (uint32_t)(as * BN);
            const int m0 = GT_TILE_M(tile) * GT_BM, n0 = GT_TILE_N0(tile);
#pragma unroll 1
            for (int c = 0; c < BN / GT_CCHUNK; ++c) {
                unsigned char* stage = sC + (size_t)cs * GT_C_STAGE_BYTES;
                // the staging buffer may still be read by the TMA store issued NC chunks ago
                if (issuer) {
                    if (NC == 2) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(1) : "memory");
                    else asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(0) : "memory");
                }
                gt_epi_barrier();
                unsigned char* srow = stage + (size_t)trow * (GT_CCHUNK * 2);
#pragma unroll
                for (int h = 0; h < GT_CCHUNK / 32; ++h) {
                    uint32_t r[32];
                    tc_ld32(t0 + c * GT_CCHUNK + h * 32, r);
                    tc_wait_ld();
                    if (a.bias) {
                        const float4* bp = reinterpret_cast<const float4*>(a.bias + n0 + c * GT_CCHUNK + h * 32);
#pragma unroll
                        for (int v = 0; v < 8; ++v) {
                            const float4 bq = __ldg(bp + v);
                            r[4 * v] = __float_as_uint(__uint_as_float(r[4 * v]) + bq.x);
                            r[4 * v + 1] = __float_as_uint(__uint_as_float(r[4 * v + 1]) + bq.y);
                            r[4 * v + 2] = __float_as_uint(__uint_as_float(r[4 * v + 2]) + bq.z);
                            r[4 * v + 3] = __float_as_uint(__uint_as_float(r[4 * v + 3]) + bq.w);
                        }
                    }
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        uint4 o;
                        o.x = gt_pack(r[8 * v], r[8 * v + 1]); o.y = gt_pack(r[8 * v + 2], r[8 * v + 3]);
                        o.z = gt_pack(r[8 * v + 4], r[8 * v + 5]); o.w = gt_pack(r[8 * v + 6], r[8 * v + 7]);
                        const int chunk16 = (h * 4 + v) ^ (trow & 7);  // TMA SWIZZLE_128B: 16-byte chunk ^ (row % 8)
                        *reinterpret_cast<uint4*>(srow + chunk16 * 16) = o;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> TMA (async proxy) reads
                gt_epi_barrier();
                if (issuer) {
                    gt_tma_store_2d(&tmC, stage, n0 + c * GT_CCHUNK, m0);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (++cs == NC) cs = 0;
            }
            tc_fence_before();
            gt_mbar_arrive(&acc_empty[as]);
            as ^= 1;
            if (as == 0) aph ^= 1u;
        }
        if (issuer) asm volatile("cp.async.bulk.wait_group %0;" ::"n"(0) : "memory");  // all stores done before exit
    }
    tc_fence_before();
    __syncthreads();
#undef GT_TILE_M
#undef GT_TILE_N0
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- host ------------------------------------------------------------------------------------------------------
TmEncodeFn tm_encode_fn() {
    static TmEncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        const cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            fail("fv_gemm: cuTensorMapEncodeTiled is not available from the CUDA driver (%s)", cudaGetErrorString(e));
            return nullptr;
        }
        encode = (TmEncodeFn)fn;
    }
    return encode;
}

struct TmKey {
    const void* base;
    int64_t rows, cols, ld;
    int box_rows;
    bool operator==(const TmKey& o) const {
        return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
    }
};
// 2-D tensor map over a row-major (rows x cols) bf16 matrix, box = (64 columns x box_rows rows), 128-byte swizzle
static int get_tmap2d(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    constexpr int NCACHE = 128;
    static thread_local TmKey keys[NCACHE];
    static thread_local CUtensorMap maps[NCACHE];
    static thread_local int used = 0, next = 0;
    const TmKey k{base, rows, cols, ld, box_rows};
    for (int i = 0; i < used; ++i)
        if (keys[i] == k) {
            *out = maps[i];
            return 0;
        }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)(ld * 2)};
    cuuint32_t box[2] = {(cuuint32_t)GT_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUtensorMap m;
    TmEncodeFn encode = tm_encode_fn();
    if (!encode) return 1;
    const CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("fv_gemm: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    const int slot = used < NCACHE ? used++ : (next++ % NCACHE);
    keys[slot] = k;
    maps[slot] = m;
    *out = m;
    return 0;
}

static int pick_bn(int N) {
    const int cands[4] = {256, 192, 128, 64};
    for (int c : cands)
        if (N % c == 0) return c;
    return 0;
}

// Shared-memory plan within 227 KB.  Resident mode (small K x BN weight block): W + A ring (2..4) + C staging (1..2).
// Stream mode (large K, e.g. FastVim-B and the patch embedding, K = 768 / 1536): ring of (A + W k-block) stages.
static bool plan_smem(int KB, int BN, int* ns, int* nc, int* stream, size_t* bytes) {
    const size_t wblk = (size_t)BN * GT_BK * 2, fixed = (size_t)(KB + 16) * 8 + 16 + 1024, cap = 227 * 1024;
    const int opts[7][2] = {{6, 2}, {5, 2}, {4, 2}, {3, 2}, {4, 1}, {3, 1}, {2, 1}};   // deepest A ring that fits: more bytes in flight
    for (auto& o : opts) {
        const size_t tot = (size_t)KB * wblk + (size_t)o[0] * GT_A_STAGE_BYTES + (size_t)o[1] * GT_C_STAGE_BYTES + fixed;
        if (tot <= cap && o[0] >= 3) {
            *ns = o[0]; *nc = o[1]; *stream = 0; *bytes = tot;
            return true;
        }
    }
    const int sopts[4][2] = {{4, 2}, {4, 1}, {3, 2}, {3, 1}};
    for (auto& o : sopts) {
        const size_t tot = (size_t)o[0] * (GT_A_STAGE_BYTES + wblk) + (size_t)o[1] * GT_C_STAGE_BYTES + fixed;
        if (tot <= cap) {
            *ns = o[0]; *nc = o[1]; *stream = 1; *bytes = tot;
            return true;
        }
    }
    return false;
}

template <int BN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmC, GemmArgs& a, int nblocks,
                       size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    FV_REQUIRE(e == cudaSuccess, "fv_gemm: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    dim3 grid;
    if (a.stream_w) {
        grid = dim3((unsigned)(a.ntiles < sm_count() ? a.ntiles : sm_count()), 1u);
    } else {
        int gx = sm_count() / nblocks;
        if (gx < 1) gx = 1;
        if (gx > a.ntiles) gx = a.ntiles;
        grid = dim3((unsigned)gx, (unsigned)nblocks);
    }
    e = launch_pdl(gemm_tc_kernel<BN>, grid, dim3(GT_THREADS), smem, st, tmA, tmW, tmC, (const GemmArgs)a);
    FV_REQUIRE(e == cudaSuccess, "fv_gemm: launch: %s", cudaGetErrorString(e));
    return finish_launch("gemm_tc");
}

}  // namespace fv

extern "C" int fv_gemm_supported(int64_t M, int N, int K) {
    using namespace fv;
    if (M <= 0 || N <= 0 || K <= 0 || K % GT_BK != 0 || M >= (1ll << 31)) return 0;
    const int BN = pick_bn(N);
    if (!BN) return 0;
    int ns, nc, stream;
    size_t bytes;
    return plan_smem(K / GT_BK, BN, &ns, &nc, &stream, &bytes) ? 1 : 0;
}

extern "C" int fv_gemm_bf16_tn(int64_t M, int N, int K, const void* A, int64_t lda, const void* W, int64_t ldw,
                               const float* bias, void* C, int64_t ldc, void* stream) {
    using namespace fv;
    FV_REQUIRE(A && W && C, "fv_gemm: null pointer");
    FV_REQUIRE(M > 0 && N > 0 && K > 0 && K % GT_BK == 0, "fv_gemm: K (%d) must be a positive multiple of 64", K);
    FV_REQUIRE(M < (1ll << 31), "fv_gemm: M too large");
    FV_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && ldc % 8 == 0 && ((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0 &&
                   ((uintptr_t)C % 16) == 0, "fv_gemm: operands must be 16-byte aligned with row strides multiple of 8");
    const int BN = pick_bn(N);
    FV_REQUIRE(BN > 0, "fv_gemm: N (%d) must be a multiple of 64", N);
    GemmArgs a;
    a.M = (int)M; a.N = N; a.K = K; a.KB = K / GT_BK; a.bias = bias;
    FV_REQUIRE(!bias || ((uintptr_t)bias % 16) == 0, "fv_gemm: bias must be 16-byte aligned");
    a.ntiles = (int)((M + GT_BM - 1) / GT_BM);
    size_t smem = 0;
    FV_REQUIRE(plan_smem(a.KB, BN, &a.nstage, &a.ncstage, &a.stream_w, &smem), "fv_gemm: no shared-memory plan for BN = %d, K = %d", BN, K);
    a.nblocks = N / BN;
    if (a.stream_w) a.ntiles *= a.nblocks;
    CUtensorMap tmA, tmW, tmC;
    if (int rc = get_tmap2d(&tmA, A, M, K, lda, GT_BM)) return rc;
    if (int rc = get_tmap2d(&tmW, W, N, K, ldw, BN)) return rc;
    if (int rc = get_tmap2d(&tmC, C, M, N, ldc, GT_BM)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int nblocks = N / BN;
#define FV_GT(BN_) \
    if (BN == BN_) return launch_gemm<BN_>(tmA, tmW, tmC, a, nblocks, smem, st);
    FV_GT(256) FV_GT(192) FV_GT(128) FV_GT(64)
#undef FV_GT
    return fail("fv_gemm: internal error");
}
