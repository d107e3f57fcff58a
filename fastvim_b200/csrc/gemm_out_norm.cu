// out_proj with the NEXT block's residual add + RMSNorm folded into its epilogue: one launch instead of
// fv_gemm_bf16_tn + fv_add_norm_fwd, and the GEMM result never goes to HBM.
//
//   C       = A (M x K) . W (N x K)^T                     bf16 operands, fp32 accumulation in tensor memory
//   res_new = res_in + C                                  fp32, written to res_out (optional)
//   Y       = res_new * rsqrt(mean(res_new^2) + eps) * w  bf16
//
// What it replaces in the reference (paths relative to /root/reference): F.linear(y, out_proj.weight)
// (mamba_ssm/modules/mamba_simple_faster.py:442-444) followed, in the next Block, by layer_norm_fn(hidden, norm.weight, None,
// residual=residual, prenorm=True, residual_in_fp32=True, is_rms_norm=True) (models/fastvim.py:175-190), i.e. the Triton
// _layer_norm_fwd_1pass_kernel (mamba_ssm/ops/triton/layernorm.py:66-121).
//
// B200 mapping.  N = d_model <= 256, so ONE accumulator (128 lanes x N fp32 columns of TMEM) holds whole output rows and one
// epilogue thread owns a whole row: the row's sum of squares is thread-local (no shuffles, no second kernel).  Seven warps:
//   warp 0  TMA producer of the (A 16 KB + W N x 128 B) k-block ring (W streams from L2, 147 KB at FastVim-T);
//   warp 1  MMA issuer (tcgen05.mma cta_group::1 kind::f16, 128 x N x 16), accumulators double-buffered in TMEM;
//   warp 6  TMA producer of the fp32 residual tile, 32 columns (128 rows x 128 B, 128-byte swizzle) at a time, ring of 3;
//   warps 2-5  epilogue, thread = row.  Pass 1, per 32-column chunk: acc (tcgen05.ld) + residual (conflict-free 16-byte
//           reads of the swizzled tile) -> sum of squares; the new residual goes (a) back into the accumulator's TMEM
//           columns (tcgen05.st) for pass 2 and (b) into a swizzled staging tile that one thread hands to TMA (whole
//           128-byte lines to HBM).  Pass 2, per 64 columns: tcgen05.ld, * rstd * w, bf16, staged, TMA store.
// HBM traffic per row: A read, residual read + written, Y written = K*2 + N*(4 + 4 + 2) bytes -- the C write + read of the
// two-launch form (4 N bytes per row) is gone.
#include "tc_common.cuh"

namespace fv {

int sm_count();

constexpr int ON_BM = 128, ON_BK = 64, ON_THREADS = 224;
constexpr uint32_t ON_A_BYTES = ON_BM * ON_BK * 2;  // 16 KB
constexpr uint32_t ON_T_BYTES = 128 * 128;          // one staged / residual tile: 128 rows x 128 B
constexpr int ON_RS = 4;                            // residual chunk ring (3 -> 4: the epilogue's largest stall was the wait for a chunk)
constexpr int ON_CS = 2;                            // output staging tiles
constexpr int ON_TQ = 4;                            // tile queue depth (flow mode)

struct OutNormArgs {
    int M, KB, ntiles, nstage;
    int has_res_out;
    const float* norm_w;
    float eps;
    // flow mode (gemm_out_norm_kernel<BN, true>): A is being produced by fv_block_fwd_signal WHILE this kernel runs
    const int* flags;   // flags[i] >= epoch  <=>  rows [i * rows_per_flag, (i + 1) * rows_per_flag) of A are complete
    int epoch, rows_per_flag;
    int* tile_ctr;      // device counter, zeroed once per forward; this launch owns values [ctr_base, ctr_base + ntiles + grid)
    int ctr_base;
};

// Tile order.  Static: tile = blockIdx.x + i * gridDim.x.  Flow: the TMA producer thread draws tiles from a device-wide
// counter (CTAs that become resident early -- on SMs the producing kernel has already left -- keep drawing) and hands them
// to the other roles through a small shared-memory queue (mbarrier full / empty pairs).
struct TileQ {
    int* slots;
    uint64_t* full;
    uint64_t* empty;
};
template <bool FLOW>
struct TileIter {
    int it = 0, slot = 0;
    uint32_t ph = 0;
    __device__ __forceinline__ int next(const OutNormArgs& a, const TileQ& q) {
        if (!FLOW) {
            const int t = blockIdx.x + it * gridDim.x;
            ++it;
            return t < a.ntiles ? t : -1;
        }
        gt_mbar_wait(&q.full[slot], ph);
        const int t = *reinterpret_cast<volatile int*>(&q.slots[slot]);
        gt_mbar_arrive(&q.empty[slot]);
        if (++slot == ON_TQ) { slot = 0; ph ^= 1u; }
        return t;
    }
};

template <int BN, bool FLOW>
__global__ void __launch_bounds__(ON_THREADS, 1)
gemm_out_norm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                     const __grid_constant__ CUtensorMap tmRin, const __grid_constant__ CUtensorMap tmRout,
                     const __grid_constant__ CUtensorMap tmY, const OutNormArgs a) {
    extern __shared__ unsigned char on_smem_raw[];
    const uint32_t raw = smem_u32(on_smem_raw);
    unsigned char* smem = on_smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KB = a.KB, NS = a.nstage;
    constexpr uint32_t W_BYTES = BN * ON_BK * 2;
    constexpr uint32_t STAGE = ON_A_BYTES + W_BYTES;
    constexpr int NCH = BN / 32;  // residual chunks per tile

    unsigned char* sAB = smem;
    unsigned char* sR = sAB + (size_t)NS * STAGE;
    unsigned char* sC = sR + (size_t)ON_RS * ON_T_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sC + (size_t)ON_CS * ON_T_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = full + NS;
    uint64_t* acc_full = empty + NS;
    uint64_t* acc_empty = acc_full + 2;
    uint64_t* r_full = acc_empty + 2;
    uint64_t* r_empty = r_full + ON_RS;
    TileQ tq;
    tq.full = r_empty + ON_RS;
    tq.empty = tq.full + ON_TQ;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tq.empty + ON_TQ);
    tq.slots = reinterpret_cast<int*>(tmem_ptr + 2);
    float* sNw = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tq.slots + ON_TQ) + 15) & ~(uintptr_t)15);   // [BN] norm weights
    for (int i = threadIdx.x; i < BN; i += ON_THREADS) sNw[i] = a.norm_w[i];

    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 128);
        }
        for (int i = 0; i < ON_RS; ++i) {
            mbar_init(&r_full[i], 1);
            mbar_init(&r_empty[i], 128);
        }
        for (int i = 0; i < ON_TQ; ++i) {
            mbar_init(&tq.full[i], 1);
            mbar_init(&tq.empty[i], 130);  // MMA issuer + residual producer + 128 epilogue threads
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
    // Static mode: the previous grid (the producer of A) must have completed.  Flow mode: no grid-level wait -- every
    // input other than A was complete before the producer of A started (it waited for ITS predecessor), and A is
    // consumed image by image behind the per-image flags.
    if (!FLOW) pdl_wait();
    pdl_trigger();

    if (warp == 0) {
        // ================= TMA producer: A and W k-blocks =================
        if (lane == 0) {
            int st = 0, it = 0, ps = 0;
            uint32_t ph = 0, pph = 0;
            for (;;) {
                int tile;
                if (FLOW) {
                    gt_mbar_wait(&tq.empty[ps], pph ^ 1u);
                    tile = atomicAdd(a.tile_ctr, 1) - a.ctr_base;
                    if (tile >= a.ntiles) tile = -1;
                    tq.slots[ps] = tile;
                    gt_mbar_arrive(&tq.full[ps]);
                    if (++ps == ON_TQ) { ps = 0; pph ^= 1u; }
                } else {
                    tile = blockIdx.x + it * gridDim.x;
                    ++it;
                    if (tile >= a.ntiles) tile = -1;
                }
                if (tile < 0) break;
                const int m0 = tile * ON_BM;
                if (FLOW) {  // the images this tile's rows belong to must be published
                    const int last_row = min(a.M - 1, m0 + ON_BM - 1);
                    for (int f = m0 / a.rows_per_flag; f <= last_row / a.rows_per_flag; ++f) {
                        int v, spins = 0;
                        for (;;) {
                            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(a.flags + f) : "memory");
                            if (v >= a.epoch || ++spins > (1 << 22)) break;  // bounded: a lost flag must not hang the GPU
                            __nanosleep(100);
                        }
                    }
                    asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy writes of A -> TMA (async proxy) reads
                }
                for (int kb = 0; kb < KB; ++kb) {
                    gt_mbar_wait(&empty[st], ph ^ 1u);
                    unsigned char* stg = sAB + (size_t)st * STAGE;
                    mbar_arrive_expect_tx(&full[st], STAGE);
                    gt_tma_2d(stg, &tmA, kb * ON_BK, m0, &full[st]);
                    gt_tma_2d(stg + ON_A_BYTES, &tmW, kb * ON_BK, 0, &full[st]);
                    if (++st == NS) { st = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(ON_BM >> 4) << 24);
            int st = 0, as = 0;
            uint32_t ph = 0, aph = 0;
            const uint32_t s_u = smem_u32(sAB);
            TileIter<FLOW> ti;
            while (ti.next(a, tq) >= 0) {
                gt_mbar_wait(&acc_empty[as], aph ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
                for (int kb = 0; kb < KB; ++kb) {
                    gt_mbar_wait(&full[st], ph);
                    tc_fence_after();
                    const uint32_t a_u = s_u + (uint32_t)st * STAGE, b_u = a_u + ON_A_BYTES;
#pragma unroll
                    for (int k = 0; k < ON_BK / 16; ++k)
                        tc_mma(d_tmem, tc_smem_desc(a_u + k * 32), tc_smem_desc(b_u + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
                    tc_commit(&empty[st]);
                    if (++st == NS) { st = 0; ph ^= 1u; }
                }
                tc_commit(&acc_full[as]);
                as ^= 1;
                if (as == 0) aph ^= 1u;
            }
        }
    } else if (warp == 6) {
        // ================= TMA producer: fp32 residual chunks (32 columns x 128 rows) =================
        if (lane == 0) {
            int rs = 0;
            uint32_t rph = 0;
            TileIter<FLOW> ti;
            for (int tile; (tile = ti.next(a, tq)) >= 0;) {
                const int m0 = tile * ON_BM;
                for (int c = 0; c < NCH; ++c) {
                    gt_mbar_wait(&r_empty[rs], rph ^ 1u);
                    mbar_arrive_expect_tx(&r_full[rs], ON_T_BYTES);
                    gt_tma_2d(sR + (size_t)rs * ON_T_BYTES, &tmRin, c * 32, m0, &r_full[rs]);
                    if (++rs == ON_RS) { rs = 0; rph ^= 1u; }
                }
            }
        }
    } else {
        // ================= epilogue (warps 2-5): TMEM lane = output row, one row per thread =================
        const int q = warp & 3;  // a warp may only touch TMEM lanes [32 * (warp % 4), +32)
        const int trow = q * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const bool issuer = warp == 2 && lane == 0;
        const uint32_t sw = (uint32_t)(trow & 7);
        int as = 0, cs = 0, rs = 0;
        uint32_t aph = 0, rph = 0;
        TileIter<FLOW> ti;
        for (int tile; (tile = ti.next(a, tq)) >= 0;) {
            const int m0 = tile * ON_BM;
            gt_mbar_wait(&acc_full[as], aph);
            tc_fence_after();
            const uint32_t t0 = lane_addr + (uint32_t)(as * BN);
            // ---- pass 1: res_new = res_in + acc, sum of squares; res_new -> TMEM (pass 2) and -> res_out
            float ss = 0.f;
#pragma unroll 1
            for (int c = 0; c < NCH; ++c) {
                uint32_t r[32];
                tc_ld32(t0 + c * 32, r);
                gt_mbar_wait(&r_full[rs], rph);
                const unsigned char* rrow = sR + (size_t)rs * ON_T_BYTES + (size_t)trow * 128;
                float4 rv[8];
#pragma unroll
                for (int v = 0; v < 8; ++v) rv[v] = *reinterpret_cast<const float4*>(rrow + ((uint32_t)v ^ sw) * 16);
                tc_wait_ld();
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    float4 o;
                    o.x = __uint_as_float(r[4 * v]) + rv[v].x; o.y = __uint_as_float(r[4 * v + 1]) + rv[v].y;
                    o.z = __uint_as_float(r[4 * v + 2]) + rv[v].z; o.w = __uint_as_float(r[4 * v + 3]) + rv[v].w;
                    ss = fmaf(o.x, o.x, ss); ss = fmaf(o.y, o.y, ss); ss = fmaf(o.z, o.z, ss); ss = fmaf(o.w, o.w, ss);
                    r[4 * v] = __float_as_uint(o.x); r[4 * v + 1] = __float_as_uint(o.y);
                    r[4 * v + 2] = __float_as_uint(o.z); r[4 * v + 3] = __float_as_uint(o.w);
                }
                gt_mbar_arrive(&r_empty[rs]);  // this thread's part of the residual chunk is in registers
                if (++rs == ON_RS) { rs = 0; rph ^= 1u; }
                tc_st32(t0 + c * 32, r);
                if (a.has_res_out) {
                    unsigned char* stage = sC + (size_t)cs * ON_T_BYTES;
                    if (issuer) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(ON_CS - 1) : "memory");
                    gt_epi_barrier();  // the staging tile's previous TMA store has read it
                    unsigned char* srow = stage + (size_t)trow * 128;
#pragma unroll
                    for (int v = 0; v < 8; ++v)
                        *reinterpret_cast<uint4*>(srow + ((uint32_t)v ^ sw) * 16) = make_uint4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    gt_epi_barrier();
                    if (issuer) {
                        gt_tma_store_2d(&tmRout, stage, c * 32, m0);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    if (++cs == ON_CS) cs = 0;
                }
            }
            tc_wait_st();
            const float rstd = rsqrtf(ss * (1.f / (float)BN) + a.eps);
            // ---- pass 2: Y = res_new * rstd * w -> bf16, 64 columns per staged tile
#pragma unroll 1
            for (int c = 0; c < BN / 64; ++c) {
                unsigned char* stage = sC + (size_t)cs * ON_T_BYTES;
                if (issuer) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(ON_CS - 1) : "memory");
                gt_epi_barrier();
                unsigned char* srow = stage + (size_t)trow * 128;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t r[32];
                    tc_ld32(t0 + c * 64 + h * 32, r);
                    tc_wait_ld();
                    const float4* wp = reinterpret_cast<const float4*>(sNw + c * 64 + h * 32);
#pragma unroll
                    for (int v = 0; v < 8; ++v) {
                        const float4 wq = wp[v];
                        r[4 * v] = __float_as_uint(__uint_as_float(r[4 * v]) * rstd * wq.x);
                        r[4 * v + 1] = __float_as_uint(__uint_as_float(r[4 * v + 1]) * rstd * wq.y);
                        r[4 * v + 2] = __float_as_uint(__uint_as_float(r[4 * v + 2]) * rstd * wq.z);
                        r[4 * v + 3] = __float_as_uint(__uint_as_float(r[4 * v + 3]) * rstd * wq.w);
                    }
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        uint4 o;
                        o.x = gt_pack(r[8 * v], r[8 * v + 1]); o.y = gt_pack(r[8 * v + 2], r[8 * v + 3]);
                        o.z = gt_pack(r[8 * v + 4], r[8 * v + 5]); o.w = gt_pack(r[8 * v + 6], r[8 * v + 7]);
                        *reinterpret_cast<uint4*>(srow + ((uint32_t)(h * 4 + v) ^ sw) * 16) = o;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                gt_epi_barrier();
                if (issuer) {
                    gt_tma_store_2d(&tmY, stage, c * 64, m0);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (++cs == ON_CS) cs = 0;
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

static bool plan_out_norm(int BN, int* ns, size_t* bytes) {
    const size_t stage = ON_A_BYTES + (size_t)BN * ON_BK * 2, fixed = 48 * 8 + 64 + 1024 + 1024, cap = 227 * 1024;
    for (int n = 4; n >= 2; --n) {
        const size_t tot = (size_t)n * stage + (size_t)(ON_RS + ON_CS) * ON_T_BYTES + fixed;
        if (tot <= cap) {
            *ns = n; *bytes = tot;
            return true;
        }
    }
    return false;
}

template <int BN, bool FLOW>
static int launch_out_norm(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmRin, const CUtensorMap& tmRout,
                           const CUtensorMap& tmY, const OutNormArgs& a, int grid, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(gemm_out_norm_kernel<BN, FLOW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    FV_REQUIRE(e == cudaSuccess, "fv_gemm_out_norm: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    e = launch_pdl(gemm_out_norm_kernel<BN, FLOW>, dim3(grid), dim3(ON_THREADS), smem, st, tmA, tmW, tmRin, tmRout, tmY, a);
    FV_REQUIRE(e == cudaSuccess, "fv_gemm_out_norm: launch: %s", cudaGetErrorString(e));
    return finish_launch("gemm_out_norm");
}

static int out_norm_impl(int64_t M, int N, int K, const void* A, int64_t lda, const void* W, int64_t ldw, const float* res_in,
                         int64_t ldr, float* res_out, const float* norm_w, float eps, void* Y, int64_t ldy, const int* flags,
                         int epoch, int rows_per_flag, int* tile_ctr, int launch_index, cudaStream_t st);

}  // namespace fv

extern "C" int fv_gemm_out_norm_supported(int64_t M, int N, int K) {
    using namespace fv;
    if (M <= 0 || K <= 0 || K % ON_BK != 0 || M >= (1ll << 31)) return 0;
    if (N != 64 && N != 128 && N != 192 && N != 256) return 0;
    int ns;
    size_t bytes;
    return plan_out_norm(N, &ns, &bytes) ? 1 : 0;
}

extern "C" int fv_gemm_out_norm(int64_t M, int N, int K, const void* A, int64_t lda, const void* W, int64_t ldw,
                                const float* res_in, int64_t ldr, float* res_out, const float* norm_w, float eps, void* Y,
                                int64_t ldy, void* stream) {
    return fv::out_norm_impl(M, N, K, A, lda, W, ldw, res_in, ldr, res_out, norm_w, eps, Y, ldy, nullptr, 0, 0, nullptr, 0,
                             (cudaStream_t)stream);
}

// Flow form: A (the gated y of fv_block_fwd_signal) is consumed image by image behind `flags` while that kernel is still
// running -- launch this right after it on the same stream.  sync[0] is the tile counter, sync[1 + i] the flag of image i
// (rows_per_flag = tokens per image); the caller zeroes `sync` once per forward and numbers the launches that share it
// launch_index = 0, 1, ... with epoch = launch_index + 1 on both kernels.
extern "C" int fv_gemm_out_norm_flow(int64_t M, int N, int K, const void* A, int64_t lda, const void* W, int64_t ldw,
                                     const float* res_in, int64_t ldr, float* res_out, const float* norm_w, float eps, void* Y,
                                     int64_t ldy, int* sync, int rows_per_flag, int launch_index, void* stream) {
    using namespace fv;
    FV_REQUIRE(sync && rows_per_flag > 0 && launch_index >= 0 && M % rows_per_flag == 0,
               "fv_gemm_out_norm_flow: sync buffer / rows_per_flag (%d) / launch_index (%d) invalid", rows_per_flag, launch_index);
    return out_norm_impl(M, N, K, A, lda, W, ldw, res_in, ldr, res_out, norm_w, eps, Y, ldy, sync + 1, launch_index + 1,
                         rows_per_flag, sync, launch_index, (cudaStream_t)stream);
}

namespace fv {
static int out_norm_impl(int64_t M, int N, int K, const void* A, int64_t lda, const void* W, int64_t ldw, const float* res_in,
                         int64_t ldr, float* res_out, const float* norm_w, float eps, void* Y, int64_t ldy, const int* flags,
                         int epoch, int rows_per_flag, int* tile_ctr, int launch_index, cudaStream_t st) {
    FV_REQUIRE(A && W && Y && res_in && norm_w, "fv_gemm_out_norm: null pointer");
    FV_REQUIRE(fv_gemm_out_norm_supported(M, N, K),
               "fv_gemm_out_norm: unsupported shape (%lld x %d x %d): N must be 64 / 128 / 192 / 256, K %% 64 == 0", (long long)M, N, K);
    FV_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && ldy % 8 == 0 && ((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0 &&
                   ((uintptr_t)Y % 16) == 0,
               "fv_gemm_out_norm: operands must be 16-byte aligned with row strides multiple of 8");
    FV_REQUIRE(ldr % 4 == 0 && ldr >= N && ((uintptr_t)res_in % 16) == 0 && (!res_out || ((uintptr_t)res_out % 16) == 0) &&
                   ((uintptr_t)norm_w % 16) == 0,
               "fv_gemm_out_norm: residual rows / norm weight must be 16-byte aligned");
    OutNormArgs a;
    a.M = (int)M; a.KB = K / ON_BK; a.norm_w = norm_w; a.eps = eps; a.has_res_out = res_out ? 1 : 0;
    a.ntiles = (int)((M + ON_BM - 1) / ON_BM);
    const int grid = a.ntiles < sm_count() ? a.ntiles : sm_count();
    a.flags = flags; a.epoch = epoch; a.rows_per_flag = rows_per_flag; a.tile_ctr = tile_ctr;
    a.ctr_base = launch_index * (a.ntiles + grid);   // every CTA draws until it gets a value >= ntiles: ntiles + grid draws
    size_t smem = 0;
    FV_REQUIRE(plan_out_norm(N, &a.nstage, &smem), "fv_gemm_out_norm: no shared-memory plan");
    CUtensorMap tmA, tmW, tmRin, tmRout, tmY;
    if (int rc = get_tmap(&tmA, A, 2, K, M, 0, lda, 64, ON_BM)) return rc;
    if (int rc = get_tmap(&tmW, W, 2, K, N, 0, ldw, 64, N)) return rc;
    if (int rc = get_tmap(&tmRin, res_in, 4, N, M, 0, ldr, 32, ON_BM)) return rc;
    if (int rc = get_tmap(&tmRout, res_out ? res_out : res_in, 4, N, M, 0, ldr, 32, ON_BM)) return rc;
    if (int rc = get_tmap(&tmY, Y, 2, N, M, 0, ldy, 64, ON_BM)) return rc;
#define FV_ON(BN_)                                                                                             \
    if (N == BN_)                                                                                              \
        return flags ? launch_out_norm<BN_, true>(tmA, tmW, tmRin, tmRout, tmY, a, grid, smem, st)              \
                     : launch_out_norm<BN_, false>(tmA, tmW, tmRin, tmRout, tmY, a, grid, smem, st);
    FV_ON(256) FV_ON(192) FV_ON(128) FV_ON(64)
#undef FV_ON
    return fail("fv_gemm_out_norm: internal error");
}
}  // namespace fv
