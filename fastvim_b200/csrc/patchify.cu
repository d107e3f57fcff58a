// Patch unfolding (im2col of a non-overlapping p x p convolution) fused with the cast to bf16.
//
// What it replaces: the reference embeds patches with nn.Conv2d(k = stride = patch) (models/fastvim.py:67-69, 95) or,
// for FastChannelVim, one shared nn.Conv3d(1, E, (1, p, p)) (models_channel_mamba_faster.py:113-121, 180-184).  A
// non-overlapping convolution is a GEMM over unfolded patches; round 1 built the unfolded matrix with three eager torch
// passes (cast, reshape/permute copy, contiguous).  This kernel reads the NCHW image ONCE in its host dtype (fp32, bf16
// or uint8) and writes the bf16 A operand of the patch-embedding GEMM once:
//   joint mode        out[(b, gy, gx), (c, py, px)]   (B*gh*gw, C*p*p)   -- Conv2d over all channels
//   per-channel mode  out[(b, c, gy, gx), (py, px)]   (B*C*gh*gw, p*p)   -- shared Conv3d filter, one token per channel
// uint8 pixels are exact in bf16 (0..255); the (x/255 - mean)/std normalisation of a uint8 pipeline is folded into the
// GEMM's weights and bias on the host (fastvim_b200/vision.py), so HBM and PCIe carry one byte per pixel.
// HBM-bound streaming: a thread moves 8 consecutive pixels of one image row (one 32-byte sector of fp32 in, 16 bytes
// out); consecutive threads walk an image row, so reads are fully coalesced and writes are whole 32-byte sectors.

#include "common.cuh"

namespace fv {

template <typename TI>
__device__ __forceinline__ void load8(const TI* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<bf16>(const bf16* p, float (&v)[8]) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
template <>
__device__ __forceinline__ void load8<uint8_t>(const uint8_t* p, float (&v)[8]) {
    const uint2 q = __ldg(reinterpret_cast<const uint2*>(p));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[i] = (float)((q.x >> (8 * i)) & 0xffu);
        v[4 + i] = (float)((q.y >> (8 * i)) & 0xffu);
    }
}

template <typename TI, bool BF16_PASSTHROUGH>
__global__ void __launch_bounds__(256)
patchify_kernel(const TI* __restrict__ img, int C, int H, int W, int p, int per_channel, int64_t total8,
                bf16* __restrict__ out) {
    pdl_wait();  // PDL: launched while the previous kernel drains; its output is visible from here
    pdl_trigger();
    const int W8 = W >> 3, p8 = p >> 3;
    const int gh = H / p, gw = W / p;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (int64_t)gridDim.x * blockDim.x) {
        // i enumerates 8-pixel groups of the image in memory order: ((b*C + c)*H + y)*W8 + x8
        const int x8 = (int)(i % W8);
        int64_t r = i / W8;
        const int y = (int)(r % H);
        r /= H;
        const int c = (int)(r % C);
        const int64_t b = r / C;
        const int gy = y / p, py = y - gy * p;
        const int gx = x8 / p8, px8 = x8 - gx * p8;
        int64_t row, col;
        int64_t K;
        if (per_channel) {
            row = ((b * C + c) * gh + gy) * gw + gx;
            col = (int64_t)py * p + px8 * 8;
            K = (int64_t)p * p;
        } else {
            row = (b * gh + gy) * gw + gx;
            col = ((int64_t)c * p + py) * p + px8 * 8;
            K = (int64_t)C * p * p;
        }
        uint4 o;
        if (BF16_PASSTHROUGH) {
            o = __ldg(reinterpret_cast<const uint4*>(img + i * 8));
        } else {
            float v[8];
            load8<TI>(img + i * 8, v);
            __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
            o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
            o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
        }
        *reinterpret_cast<uint4*>(out + row * K + col) = o;
    }
}

int sm_count();

}  // namespace fv

extern "C" int fv_patchify_supported(int in_dtype, int C, int H, int W, int patch) {
    if (in_dtype < 0 || in_dtype > 2 || C <= 0 || H <= 0 || W <= 0 || patch <= 0) return 0;
    return patch % 8 == 0 && H % patch == 0 && W % patch == 0;
}

extern "C" int fv_patchify(int in_dtype, int batch, int C, int H, int W, int patch, int per_channel, const void* img,
                           void* out, void* stream) {
    using namespace fv;
    FV_REQUIRE(img && out, "fv_patchify: null pointer");
    FV_REQUIRE(fv_patchify_supported(in_dtype, C, H, W, patch),
               "fv_patchify: needs in_dtype in {0 f32, 1 bf16, 2 u8}, patch %% 8 == 0 and H, W multiples of the patch "
               "(got dtype %d, %d x %d, patch %d)", in_dtype, H, W, patch);
    FV_REQUIRE(batch > 0, "fv_patchify: batch must be positive");
    FV_REQUIRE(((uintptr_t)img % 16) == 0 && ((uintptr_t)out % 16) == 0, "fv_patchify: pointers must be 16-byte aligned");
    const int64_t total8 = (int64_t)batch * C * H * (W / 8);
    const int64_t want = (total8 + 255) / 256;
    const int grid = (int)(want < (int64_t)sm_count() * 16 ? want : (int64_t)sm_count() * 16);
    cudaStream_t st = (cudaStream_t)stream;
    if (in_dtype == 0)
        FV_LAUNCH_PDL((patchify_kernel<float, false>), grid, 256, 0, st, (const float*)img, C, H, W, patch, per_channel, total8, (bf16*)out);
    else if (in_dtype == 1)
        FV_LAUNCH_PDL((patchify_kernel<bf16, true>), grid, 256, 0, st, (const bf16*)img, C, H, W, patch, per_channel, total8, (bf16*)out);
    else
        FV_LAUNCH_PDL((patchify_kernel<uint8_t, false>), grid, 256, 0, st, (const uint8_t*)img, C, H, W, patch, per_channel, total8,
                                                              (bf16*)out);
    return finish_launch("patchify");
}
