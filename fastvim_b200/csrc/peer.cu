// Peer-memory exchanges for the single-image (2048 x 2048) multi-GPU mode: the collectives of the sharded block are done
// by OUR kernels over NVLink peer pointers (symmetric memory mapped by the host), not by NCCL.
//
// The reference has no counterpart (it is data-parallel only).  Round 1 sharded d_inner over the
// ranks and called three NCCL collectives per block (72 per step): at 11 KB .. 12.6 MB per message they are latency-bound
// (~20 us each inside a CUDA graph) and N = 2 ran at 0.55x of one GPU.  Here every exchange is ONE kernel:
//   barrier   CTA 0: thread q tells peer q "I have entered my k-th peer kernel" with a one-way system-scope RED on the
//             peer's arrival counter and spins on its own counter for q; k is a device-resident sequence number, so the
//             protocol is replay-safe (no epochs baked into a captured CUDA graph) and costs one NVLink one-way latency.
//             The other CTAs of the launch are released through local tokens.
//   data      every CTA then reads the peers' buffers directly (ld.global.cg over NVLink) -- a rank-ordered sum of fp32
//             partials (x_proj partial products, LayerNorm sums: deterministic, identical on all ranks) or a strided 2-D
//             copy (the token <-> channel all-to-all) -- and writes local memory.
// A spin that lasts longer than ~2 s sets an error word and falls through instead of hanging the GPU.
// Symmetric buffer layout (same on every rank): [0, 384) flag words (arrivals[8] | seq | go, one 128-byte line each), [2048] error word, data from
// byte 4096 on (the host carves it).

#include "common.cuh"

namespace fv {

constexpr int PEER_MAX = 8, PEER_MAX_CTAS = 128;
constexpr int PEER_OFF_CTA = 1024, PEER_OFF_ERR = 2048;

struct PeerBase {
    unsigned char* buf[PEER_MAX];  // base of every rank's symmetric buffer, as mapped in THIS process
    int world, rank;
};

// flag words of a rank's buffer: bytes [0, 32) arrivals[q] = number of peer kernels rank q has entered (written by q with a
// one-way system-scope RED), byte 128 seq = number of peer kernels THIS rank has entered, byte 256 go = release tokens for
// the other CTAs of the current launch (each on its own 128-byte line).  All state lives in device memory, so a captured CUDA graph replays correctly.
__device__ __forceinline__ unsigned int ld_volatile_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_add_sys(unsigned int* p, unsigned int v) {
    asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// all ranks have reached this kernel (so everything they launched before it is complete and visible)
__device__ __forceinline__ void peer_sync(const PeerBase& pb) {
    unsigned char* mine = pb.buf[pb.rank];
    unsigned int* err = reinterpret_cast<unsigned int*>(mine + PEER_OFF_ERR);
    unsigned int* flags = reinterpret_cast<unsigned int*>(mine);   // arrivals[8]: their own 128-byte line (remote REDs land here)
    unsigned int* seq = flags + 32;                                   // byte 128
    unsigned int* go = flags + 64;                                    // byte 256: polled by the other CTAs, away from the arrivals
    __shared__ unsigned int s_expected;
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) s_expected = ++seq[0];          // single writer: CTA 0 of the (stream-serialised) peer kernels
        __syncthreads();
        const unsigned int expected = s_expected;
        const int q = threadIdx.x;
        if (q < pb.world && q != pb.rank) {
            __threadfence_system();
            red_add_sys(reinterpret_cast<unsigned int*>(pb.buf[q]) + pb.rank, 1u);   // one-way: "rank entered kernel #expected"
            const long long t0 = clock64();
            while ((int)(ld_volatile_sys(flags + q) - expected) < 0) {
                if (clock64() - t0 > (1ll << 32)) {  // ~2 s: a peer never arrived
                    atomicExch(err, 1u);
                    break;
                }
                __nanosleep(20);
            }
            __threadfence_system();
        }
        __syncthreads();
        if (threadIdx.x == 0 && gridDim.x > 1) {
            __threadfence();
            atomicAdd(go, gridDim.x - 1);
        }
    } else {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            while (true) {   // take one release token
                // exactly gridDim.x - 1 tokens are published for the gridDim.x - 1 waiting CTAs, so once tokens are visible
                // ONE atomic decrement per CTA always succeeds (a CAS retry loop made the release O(CTAs^2): 100 us at 128 CTAs)
                if (*reinterpret_cast<volatile unsigned int*>(go) > 0) {
                    atomicSub(go, 1u);
                    break;
                }
                if (clock64() - t0 > (1ll << 33)) {
                    atomicExch(err, 1u);
                    break;
                }
                __nanosleep(100);
            }
            __threadfence();
        }
        __syncthreads();
    }
}

struct PeerSumArgs {
    PeerBase pb;
    int64_t off;   // byte offset of the fp32 partial inside every rank's buffer
    int64_t n;     // floats (multiple of 4)
    float* out32;
    bf16* out16;
};

__global__ void __launch_bounds__(256) peer_sum_kernel(const PeerSumArgs a) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    peer_sync(a.pb);
    const int64_t n4 = a.n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < PEER_MAX; ++r) {   // rank order: the same sum on every rank
            if (r < a.pb.world) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(a.pb.buf[r] + a.off) + i);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        if (a.out32) reinterpret_cast<float4*>(a.out32)[i] = acc;
        if (a.out16) st4(a.out16 + 4 * i, acc);
    }
}

struct PeerCopyArgs {
    PeerBase pb;
    int nparts;                    // column segments per peer (1 or 2)
    int rows;                      // rows copied from every peer
    int seg16;                     // 16-byte chunks per row segment
    int64_t src_off[PEER_MAX][2];  // byte offset of (first row, segment) inside peer q's buffer
    int64_t src_ld;                // bytes between rows at the source
    int64_t dst_off[PEER_MAX][2];  // byte offset inside the local destination
    int64_t dst_ld;
    unsigned char* dst;
};

__global__ void __launch_bounds__(256) peer_copy_kernel(const PeerCopyArgs a) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    peer_sync(a.pb);
    const int64_t per_peer = (int64_t)a.nparts * a.rows * a.seg16;
    const int64_t total = per_peer * a.pb.world;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    constexpr int UN = 8;   // 16-byte NVLink loads in flight per thread: the copy is latency- not bandwidth-bound otherwise
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += stride * UN) {
        uint4 v[UN];
        unsigned char* d[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int64_t i = i0 + u * stride;
            d[u] = nullptr;
            if (i < total) {
                // peer varies slowest so that consecutive threads stream one peer's rows; every rank starts on a different
                // peer (rank + k) to spread the load over the NVSwitch ports
                const int k = (int)(i / per_peer);
                int64_t r = i - (int64_t)k * per_peer;
                const int q = (a.pb.rank + k) % a.pb.world;
                const int part = (int)(r / ((int64_t)a.rows * a.seg16));
                r -= (int64_t)part * a.rows * a.seg16;
                const int row = (int)(r / a.seg16), c = (int)(r - (int64_t)row * a.seg16);
                v[u] = __ldcg(reinterpret_cast<const uint4*>(a.pb.buf[q] + a.src_off[q][part] + (int64_t)row * a.src_ld) + c);
                d[u] = a.dst + a.dst_off[q][part] + (int64_t)row * a.dst_ld + (int64_t)c * 16;
            }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u)
            if (d[u]) *reinterpret_cast<uint4*>(d[u]) = v[u];
    }
}

static int fill_base(PeerBase& pb, int world, int rank, const void* const* bufs, const char* who) {
    FV_REQUIRE(world >= 1 && world <= PEER_MAX && rank >= 0 && rank < world && bufs, "%s: world %d / rank %d out of range (max %d ranks)",
               who, world, rank, PEER_MAX);
    for (int q = 0; q < PEER_MAX; ++q) pb.buf[q] = q < world ? (unsigned char*)bufs[q] : nullptr;
    for (int q = 0; q < world; ++q) FV_REQUIRE(pb.buf[q] && ((uintptr_t)pb.buf[q] % 16) == 0, "%s: peer buffer %d is null or misaligned", who, q);
    pb.world = world;
    pb.rank = rank;
    return 0;
}

int sm_count();

}  // namespace fv

extern "C" int64_t fv_peer_header_bytes(void) { return 4096; }

/* out = sum over ranks of the fp32 array at byte offset `off` of every rank's symmetric buffer (n floats, n % 4 == 0),
 * after a barrier over all ranks.  out32 (fp32) and / or out16 (bf16) receive the sum. */
extern "C" int fv_peer_sum_f32(int world, int rank, const void* const* bufs, int64_t off, int64_t n, float* out32, void* out16,
                               void* stream) {
    using namespace fv;
    PeerSumArgs a;
    if (int rc = fill_base(a.pb, world, rank, bufs, "fv_peer_sum_f32")) return rc;
    FV_REQUIRE(n > 0 && n % 4 == 0 && off >= 4096 && off % 16 == 0, "fv_peer_sum_f32: n (%lld) must be a positive multiple of 4 and off >= 4096, 16-byte aligned",
               (long long)n);
    FV_REQUIRE(out32 || out16, "fv_peer_sum_f32: no output");
    a.off = off; a.n = n; a.out32 = out32; a.out16 = (bf16*)out16;
    int64_t want = (n / 4 + 255) / 256;
    const int grid = (int)(want < 1 ? 1 : (want > 64 ? 64 : want));
    FV_LAUNCH_PDL((peer_sum_kernel), grid, 256, 0, (cudaStream_t)stream, a);
    return finish_launch("peer_sum_f32");
}

/* Strided 2-D gather from every rank (the token <-> channel all-to-all), after a barrier over all ranks: for peer q and
 * segment k < nparts copy `rows` rows of `row_bytes` bytes from (bufs[q] + src_off[q*2+k] + row*src_ld) to
 * (dst + dst_off[q*2+k] + row*dst_ld).  row_bytes, offsets and strides are multiples of 16. */
extern "C" int fv_peer_copy2d(int world, int rank, const void* const* bufs, int nparts, int rows, int64_t row_bytes,
                              const int64_t* src_off, int64_t src_ld, const int64_t* dst_off, int64_t dst_ld, void* dst,
                              void* stream) {
    using namespace fv;
    PeerCopyArgs a;
    if (int rc = fill_base(a.pb, world, rank, bufs, "fv_peer_copy2d")) return rc;
    FV_REQUIRE(dst && src_off && dst_off && (nparts == 1 || nparts == 2) && rows > 0 && row_bytes > 0, "fv_peer_copy2d: bad arguments");
    FV_REQUIRE(row_bytes % 16 == 0 && src_ld % 16 == 0 && dst_ld % 16 == 0 && ((uintptr_t)dst % 16) == 0,
               "fv_peer_copy2d: row bytes / strides / destination must be 16-byte aligned");
    for (int q = 0; q < world; ++q)
        for (int k = 0; k < nparts; ++k) {
            FV_REQUIRE(src_off[q * 2 + k] >= 4096 && src_off[q * 2 + k] % 16 == 0 && dst_off[q * 2 + k] % 16 == 0,
                       "fv_peer_copy2d: offsets must be 16-byte aligned (source past the 4096-byte header)");
            a.src_off[q][k] = src_off[q * 2 + k];
            a.dst_off[q][k] = dst_off[q * 2 + k];
        }
    a.nparts = nparts; a.rows = rows; a.seg16 = (int)(row_bytes / 16);
    a.src_ld = src_ld; a.dst_ld = dst_ld; a.dst = (unsigned char*)dst;
    const int64_t total = (int64_t)world * nparts * rows * a.seg16;
    int64_t want = (total + 256 * 8 - 1) / (256 * 8);
    const int grid = (int)(want < 1 ? 1 : (want > PEER_MAX_CTAS ? PEER_MAX_CTAS : want));
    FV_LAUNCH_PDL((peer_copy_kernel), grid, 256, 0, (cudaStream_t)stream, a);
    return finish_launch("peer_copy2d");
}

/* 1 when a peer spin timed out since the buffer was initialised (reads the error word of the LOCAL buffer; synchronises). */
extern "C" int fv_peer_error(const void* local_buf) {
    unsigned int v = 0;
    if (!local_buf) return -1;
    if (cudaMemcpy(&v, (const unsigned char*)local_buf + fv::PEER_OFF_ERR, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int)v;
}
