/* fastvim_b200 -- C ABI of the B200-native FastVim SSM-block hot path.
 *
 * The reference binds its native code through pybind11 tensor-level entry points:
 *   selective_scan_cuda.fwd / .bwd        mamba-1p1p1/csrc/selective_scan/selective_scan.cpp:226-336, 338-492, 494-497
 *   causal_conv1d_cuda.causal_conv1d_fwd  call sites mamba_ssm/ops/selective_scan_interface.py:496-498, 751-753
 *   Triton add+norm                       mamba_ssm/ops/triton/layernorm.py:124-191, 307-399
 * This header is the drop-in boundary that replaces them: plain pointers and sizes, no
 * torch types.  Contract for every entry point:
 *   - all data pointers are DEVICE pointers owned by the caller; the library never
 *     allocates, frees or synchronises, and launches only on `stream` (a cudaStream_t);
 *   - returns 0 on success, non-zero on error; fv_last_error() gives the thread-local text;
 *   - activations are `dtype` (FV_F32 | FV_BF16); parameters and scan carries are fp32;
 *   - no mutable global state: safe to call from several host threads / streams.
 *
 * Layout.  Between in_proj and out_proj the reference keeps activations as (B, D, L) with
 * L contiguous (mamba_simple_faster.py:189-193).  The B200 path keeps them TOKEN-MAJOR,
 * (B, L, D) with D contiguous, so that (i) in_proj/out_proj are plain row-major GEMMs with
 * no transposes, (ii) every kernel is coalesced across channels, and (iii) the odd-layer
 * token rotation of models/fastvim.py:192-210 becomes an index map (fv_geom strides)
 * instead of two copies.  The (B, D, L) operator API (selective_scan_fn) is served by
 * fv_selective_scan_*.
 */
#ifndef FASTVIM_B200_H_
#define FASTVIM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { FV_F32 = 0, FV_BF16 = 1 } fv_dtype;
/* fv_gemm_bf16 only: fp32 output that is ADDED to (C += A.B, every K split adds its partial sum with a TMA reduction store;
 * the caller provides the initial value, e.g. zeros; summation order over the splits is not fixed) */
#define FV_F32_ACC 2
typedef enum { FV_POOL_MEAN = 0, FV_POOL_MAX = 1 } fv_pool_mode;

/* Sequence geometry of one mixer call.  The L tokens of an image are viewed as
 * (outer, pool, inner): sequence position t = (o*pool + p)*inner + i, pooled position
 * j = o*inner + i (Lp = outer*inner), and the token lives at memory row
 * o*tok_stride_outer + p*tok_stride_pool + i*tok_stride_inner of its image.
 *   FastVim even layer (Hr x Wc grid): (Hr, Wc, 1), strides (Wc, 1, 0)   mamba_simple_faster.py:287-297
 *   FastVim odd layer (rotated):       (Wc, Hr, 1), strides (1, Wc, 0)   models/fastvim.py:192-210, 244-260
 *   ChannelVim channel-first:          (Hr, Wc, tpp)                      mamba_simple_channel_faster.py:225-256
 */
typedef struct {
    int32_t batch;  /* images */
    int32_t dim;    /* d_inner channels handled by this call (a shard when channel-sharded) */
    int32_t outer, pool, inner;
    int64_t tok_stride_outer, tok_stride_pool, tok_stride_inner;
} fv_geom;

const char* fv_last_error(void);
int fv_version(void);
/* Number of kernels launched by this library on the calling thread since the last reset
 * (bench.py's `gpu_launches`). */
/* Programmatic dependent launch of the inference-chain kernels (in_proj GEMM, fv_block_fwd, out_proj GEMM): on by default
 * (FASTVIM_PDL=0 disables it); fv_set_pdl() switches it at run time (process-wide) and returns the previous setting -- bench.py
 * captures a second, serialised graph with it off to read per-kernel durations that do not include the wait for the
 * predecessor. */
int fv_set_pdl(int on);
/* The same attribute for every other kernel of the library (default off: FASTVIM_PDL_ALL=1 or this call; returns the previous
 * setting).  Measured: -1..2 % on FastVim-S/B inference, +0.7 % at 2048^2, within the run-to-run spread on the training step
 * (FASTVIM_TRAIN_PDL=1 switches it on around the Python training path's forward and backward). */
int fv_set_pdl_all(int on);
int64_t fv_launch_count(void);
void fv_reset_launch_count(void);

/* ---- K1: depthwise causal conv (both directions) + SiLU + pooling ------------------
 * Replaces x.flip + 2x causal_conv1d_fn + 2x reshape.mean  (mamba_simple_faster.py:272-305;
 * fused-path selective_scan_interface.py:496-508).  In original token coordinates the
 * b-direction is the anti-causal conv.
 *   x        (B, L, dim) token-major, row stride ldx elements, image stride x_bstride
 *   conv_w   (2, dim, 4) fp32   [0] = conv1d.weight, [1] = conv1d_b.weight
 *   conv_b   (2, dim)    fp32 or NULL
 *   u_out    (2, B, Lp, dim) dtype -- pooled conv output, direction-major
 */
int fv_conv_pool_fwd(const fv_geom* g, int dtype, const void* x, int64_t ldx, int64_t x_bstride,
                     const float* conv_w, const float* conv_b, float scale, int pool_mode,
                     void* u_out, void* stream);

/* K1 / K2b for channel layouts (inner > 1; FastChannelVim Channel-First, mamba_simple_channel_faster.py:225-289, 325-340,
 * 400-420), bf16: fv_conv_pool_w_fwd = fv_conv_pool_fwd that stages each (image, outer) group of pool*inner consecutive
 * sequence positions through shared memory and ALSO writes the D-skip term w = (D_f xc_f + D_b xc_b) / 2 of every token
 * (w_out (B, L, dim) bf16 in memory-row order, or NULL); fv_gate_w_fwd then is a streaming pass
 * y = LayerNorm(w + (s_f[j] + s_b[j]) / 2) * silu(z) with one warp per token (no convolution is re-evaluated).
 * Dskip (2, dim) fp32.  fv_conv_pool_w_supported(): bf16, inner >= 2, dim % 8 == 0, dim <= 2048, inner slots fit. */
int fv_conv_pool_w_supported(const fv_geom* g, int dtype);
int fv_conv_pool_w_fwd(const fv_geom* g, int dtype, const void* x, int64_t ldx, int64_t x_bstride, const float* conv_w,
                       const float* conv_b, float scale, int pool_mode, const float* Dskip, void* u_out, void* w_out,
                       void* stream);
int fv_gate_w_fwd(const fv_geom* g, int dtype, const void* w, const void* z, int64_t ldz, int64_t z_bstride,
                  const float* s, const float* ln_w, const float* ln_b, float eps, void* y, int64_t ldy,
                  int64_t y_bstride, void* stream);

/* ---- K2a: bidirectional selective scan over the pooled sequence, dt_proj fused -----
 * Replaces dt_proj matmul + 2x selective_scan_fn(D=None, z=None, delta_softplus=True)
 * (mamba_simple_faster.py:328-354, 384-410; kernel csrc/selective_scan/selective_scan_fwd_kernel.cuh:67-303).
 *   u        (2, B, Lp, dim) dtype
 *   xdbl     (2, B*Lp, ld_xdbl) dtype: columns [0,R) dt low-rank, [R,R+N) B, [R+N,R+2N) C  (x_proj output)
 *   dt_w     (2, dim, R) fp32; dt_bias (2, dim) fp32; A (2, dim, N) fp32 (A_log if a_is_log)
 *   s_out    (2, B, Lp, dim) fp32: [0] = scan_f[j], [1] = scan_b[j], both in original row order
 */
int fv_scan_fwd(const fv_geom* g, int dtype, const void* u, const void* xdbl, int64_t ld_xdbl,
                int dt_rank, int dstate, const float* dt_w, const float* dt_bias, const float* A,
                int a_is_log, float* s_out, void* stream);

/* ---- K2b: broadcast-back + D skip + direction average + LayerNorm + SiLU(z) gate ---
 * Replaces repeat_interleave, += D*x, flip/add//2, LayerNorm(d_inner), *silu(z)
 * (mamba_simple_faster.py:356-358, 412-416, 434-441).  The conv outputs are recomputed
 * from x instead of being stored.
 *   s         (2, B, Lp, dim) fp32 from fv_scan_fwd (the two planes are added here)
 *   ln_w/ln_b NULL => use_norm_after_ssm=False (:445-453).
 *   y        (B, L, dim) dtype token-major, row stride ldy: the out_proj GEMM input.
 * Channel-sharded mode (dim is a shard of d_inner): pass stats (B, L, 2) fp32; the kernel
 * then writes the PRE-norm value to y and the shard's per-token (sum, sum of squares) to
 * stats; after the all-reduce call fv_norm_gate_apply.
 */
int fv_gate_fwd(const fv_geom* g, int dtype, const void* x, const void* z, int64_t ldxz,
                int64_t xz_bstride, const float* s, const float* conv_w, const float* conv_b,
                const float* Dskip, const float* ln_w, const float* ln_b, float eps, void* y,
                int64_t ldy, int64_t y_bstride, float* stats, void* stream);
int fv_norm_gate_apply(const fv_geom* g, int dtype, int full_dim, void* y, int64_t ldy,
                       int64_t y_bstride, const void* z, int64_t ldz, int64_t z_bstride,
                       const float* stats, const float* ln_w, const float* ln_b, float eps,
                       void* stream);

/* ---- K-fused: the whole block interior (K1 + x_proj + dt_proj + K2a + K2b) in one launch ------
 * Replaces mamba_simple_faster.py:272-453 between the in_proj output and the out_proj input for the
 * common case: bf16, plain (outer, pool, 1) geometry, mean pooling, d_state 16, dt_rank <= 16, dim <= 384, and one
 * image's (L + 6) x dim bf16 slab + pooled buffers fitting the 227 KB of shared memory of one SM
 * (224^2 FastVim-T: 196 x 384).  A persistent CTA per SM keeps the image's x resident in shared memory,
 * so HBM sees x and z once and y once.  fv_block_fwd_supported() returns 1 when the configuration
 * qualifies; callers use the four-launch path (fv_conv_pool_fwd .. fv_gate_fwd) otherwise.
 *   x, z       (B, L, dim) bf16 token-major halves of the in_proj output, row stride ldxz
 *   xproj_w    (2, R+2N, dim) bf16   [x_proj.weight, x_proj_b.weight]
 *   xproj_w_packed  the same weights in MMA-fragment order, produced once per weight version by
 *              fv_block_pack_xproj into fv_block_pack_xproj_bytes(dim, R+2N) bytes (dim % 64 == 0); NULL =>
 *              the kernel gathers fragments from xproj_w (slower: 24 scattered 4-byte loads per lane)
 *   dt_w       (2, dim, R) fp32      [dt_proj.weight, dt_proj_b.weight]; R in {4, 8, 12, 16}
 *   conv_w, conv_b, dt_bias, A, Dskip, ln_w, ln_b: fp32 as in the calls above (ln_w NULL => no norm)
 *   scale      scaling_factor (the mean's 1/pool is applied inside)
 *   y          (B, L, dim) bf16, row stride ldy
 *   u_out (2, B, Lp, dim) bf16, xdbl_out (2, B*Lp, R+2N) bf16, s_out (2, B, Lp, dim) fp32: optional
 *              (NULL) copies of the pooled intermediates, saved for the backward kernels.
 */
int fv_block_fwd_supported(const fv_geom* g, int dtype, int dt_rank, int dstate);
int64_t fv_block_pack_xproj_bytes(int dim, int ncols);
int fv_block_pack_xproj(int dim, int ncols, const void* xproj_w, void* packed, void* stream);
int fv_block_fwd(const fv_geom* g, int dtype, const void* x, const void* z, int64_t ldxz,
                 int64_t xz_bstride, const float* conv_w, const float* conv_b, const void* xproj_w,
                 const void* xproj_w_packed, const float* dt_w, const float* dt_bias, const float* A, int a_is_log, int dt_rank,
                 int dstate, const float* Dskip, const float* ln_w, const float* ln_b, float eps,
                 float scale, void* y, int64_t ldy, int64_t y_bstride, void* u_out,
                 void* xdbl_out, float* s_out, void* v_out, float* pre_out, void* stream);
/* pre_out (optional, with v_out): (2, B, Lp, dim) fp32 dt_proj pre-activation dt_bias + W_dt . dt, the input
 * fv_scan_bwd_short otherwise needs from a separate fp32 GEMM.
 * v_out (optional): (B, L, dim) bf16 in memory token order, the pre-norm merged value
 *   v = (s_f[j] + s_b[j] + D_f xc_f + D_b xc_b) / 2   (the reference's (out + out_b.flip) / 2, mamba_simple_faster.py:438)
 * saved for fv_gate_bwd_v.  Only the cluster kernel (dim a multiple of 192, <= 16 pooled rows) writes it:
 * fv_block_fwd_saves_v() == 1; fv_block_fwd fails when v_out is given and the configuration does not qualify. */
int fv_block_fwd_saves_v(const fv_geom* g, int dtype, int dt_rank, int dstate);

/* ---- fused residual add + RMSNorm / LayerNorm (prenorm form) ------------------------
 * Replaces mamba_ssm/ops/triton/layernorm.py:66-121 as used by Block.forward
 * (models/fastvim.py:167-190): residual_out = x + residual (fp32), y = norm(residual_out)*w (+b).
 *   residual_in may be NULL (first block); rstd_out/mean_out (rows) optional, for backward.
 */
int fv_add_norm_fwd(int dtype, int64_t rows, int cols, const void* x, int64_t ldx,
                    const float* residual_in, const float* weight, const float* bias, float eps,
                    int is_rms, void* y, int64_t ldy, float* residual_out, float* mean_out,
                    float* rstd_out, void* stream);

/* ---- operator API: selective_scan_fn on (B, D, L), L contiguous ---------------------
 * Replaces selective_scan_cuda.fwd (selective_scan.cpp:226-336).  Real A only.
 *   u, delta, z, out: (batch, dim, L) dtype, row stride = L (contiguous)
 *   A (dim, N) fp32; B, C: (batch, groups, N, L) dtype ("variable") ; D, delta_bias (dim) fp32 or NULL
 *   last_state (batch, dim, N) fp32 or NULL
 */
int fv_selective_scan_fwd(int dtype, int batch, int dim, int64_t L, int dstate, int groups,
                          const void* u, const void* delta, const float* A, const void* B,
                          const void* C, const float* D, const void* z, const float* delta_bias,
                          int delta_softplus, void* out, float* last_state, void* stream);


/* ---- patch unfolding + cast: the A operand of the patch-embedding GEMM -----------------------------
 * Replaces the Conv2d(k = stride = patch) of PatchEmbed.proj (models/fastvim.py:67-69, 95) / the shared
 * Conv3d(1, E, (1, p, p)) of PatchEmbedPerChannel (models_channel_mamba_faster.py:113-121, 180-184) by
 * "unfold once, then GEMM".  img (batch, C, H, W) contiguous in in_dtype (0 = fp32, 1 = bf16, 2 = uint8);
 * out bf16: joint mode (batch*gh*gw, C*p*p) with column (c, py, px); per-channel mode (batch*C*gh*gw, p*p).
 * patch % 8 == 0, H and W multiples of patch (pad first otherwise).  uint8 pixels are converted exactly.
 */
int fv_patchify_supported(int in_dtype, int C, int H, int W, int patch);
int fv_patchify(int in_dtype, int batch, int C, int H, int W, int patch, int per_channel, const void* img,
                void* out, void* stream);

/* y = LayerNorm_cols(v; ln_w, ln_b) * silu(z) per row (ln_w NULL: y = v * silu(z)): the token-side half of
 * mamba_simple_faster.py:437-453 for rows holding ALL d_inner channels of a token (hybrid-sharded multi-GPU mode). */
int fv_ln_gate_fwd(int dtype, int64_t rows, int cols, const void* v, int64_t ldv, const void* z, int64_t ldz,
                   const float* ln_w, const float* ln_b, float eps, void* y, int64_t ldy, void* stream);

/* ---- peer-memory exchanges of the single-image multi-GPU mode (csrc/peer.cu) ------------------------------------
 * The reference is data-parallel only; BASELINE.json configs[4] (one 2048 x 2048 image over 8 GPUs) shards the block.
 * Every rank owns one SYMMETRIC buffer (same size and layout on all ranks, mapped into every process by the host, e.g.
 * torch.distributed._symmetric_memory); bufs[q] is rank q's buffer as addressed from the calling process.  The first
 * fv_peer_header_bytes() bytes hold the flag words (zero them once, then barrier on the host before the first call).
 * Each call = a barrier over all ranks followed by direct NVLink reads of the peers' buffers, in ONE kernel; all ranks
 * must issue the same sequence of fv_peer_* calls.  No NCCL on the data path. */
int64_t fv_peer_header_bytes(void);
int fv_peer_sum_f32(int world, int rank, const void* const* bufs, int64_t off, int64_t n, float* out32, void* out16,
                    void* stream);
int fv_peer_copy2d(int world, int rank, const void* const* bufs, int nparts, int rows, int64_t row_bytes,
                   const int64_t* src_off, int64_t src_ld, const int64_t* dst_off, int64_t dst_ld, void* dst,
                   void* stream);
int fv_peer_error(const void* local_buf);

/* ---- tcgen05 / TMEM / TMA GEMMs for the projections -------------------------------------------
 * C (M x N) = A (M x K) . W (N x K)^T: bf16 operands (row-major, row strides lda / ldw / ldc elements,
 * 16-byte aligned, multiples of 8), fp32 accumulation in tensor memory, bf16 result.  Replaces the cuBLAS
 * calls of in_proj / out_proj (mamba_simple_faster.py:189-195, 442-444) and of the patch embedding
 * (models/fastvim.py:67-103).  K % 64 == 0, N % 64 == 0.  When the CTA's (N-block x K) slice of W fits shared
 * memory it stays resident (FastVim-T/S); otherwise W k-blocks stream with A (FastVim-B, patch embedding).
 * fv_gemm_supported() returns 1 when (M, N, K) qualifies; callers use cuBLAS otherwise.
 */
int fv_gemm_supported(int64_t M, int N, int K);
int fv_gemm_bf16_tn(int64_t M, int N, int K, const void* A, int64_t lda, const void* W, int64_t ldw,
                    const float* bias /* (N) fp32 or NULL */, void* C, int64_t ldc, void* stream);
/* out_proj with the NEXT block's residual add + RMSNorm folded into its epilogue (one launch instead of fv_gemm_bf16_tn +
 * fv_add_norm_fwd; the GEMM result never goes to HBM): C = A . W^T stays in tensor memory, res_new = res_in + C is written
 * to res_out (fp32, may alias res_in, may be NULL for the model's final norm) and parked back in tensor memory, then
 * Y = res_new * rsqrt(mean(res_new^2) + eps) * norm_w (bf16).  Reference: F.linear (mamba_simple_faster.py:442-444) followed
 * by layer_norm_fn(..., residual, prenorm=True, is_rms_norm=True) (models/fastvim.py:175-190; ops/triton/layernorm.py:66-121).
 * N in {64, 128, 192, 256} (the whole row sits in one accumulator; FastVim-T: d_model = 192), K % 64 == 0. */
int fv_gemm_out_norm_supported(int64_t M, int N, int K);
int fv_gemm_out_norm(int64_t M, int N, int K, const void* A, int64_t lda, const void* W, int64_t ldw,
                     const float* res_in, int64_t ldr, float* res_out, const float* norm_w, float eps, void* Y,
                     int64_t ldy, void* stream);
/* Dataflow between fv_block_fwd and the out_proj GEMM (inference, one-CTA-per-image kernel).  At batch 256 on 148 SMs the block
 * kernel runs two rounds and leaves 40 SMs idle in the second; fv_block_fwd_signal publishes done_flags[img] = done_epoch
 * (release) as each image's y rows complete, and fv_gemm_out_norm_flow -- launched right after it on the same stream, with
 * programmatic dependent launch, so its CTAs become resident on the SMs the block kernel has already left -- draws
 * 128-row tiles from a device counter and waits only for the flags of the images a tile covers.  `sync` is an int32 buffer
 * of 1 + batch entries ([0] tile counter, [1 + i] flag of image i) that the caller zeroes once per forward; the launch pairs
 * sharing it are numbered launch_index = 0, 1, ... (done_epoch = launch_index + 1).  Same arithmetic as fv_block_fwd /
 * fv_gemm_out_norm; with FASTVIM_PDL=0 the pair simply runs back to back. */
int fv_block_fwd_signal_supported(const fv_geom* g, int dtype, int dt_rank, int dstate);
int fv_block_fwd_signal(const fv_geom* g, int dtype, const void* x, const void* z, int64_t ldxz, int64_t xz_bstride,
                        const float* conv_w, const float* conv_b, const void* xproj_w, const void* xproj_w_packed,
                        const float* dt_w, const float* dt_bias, const float* A, int a_is_log, int dt_rank, int dstate,
                        const float* Dskip, const float* ln_w, const float* ln_b, float eps, float scale, void* y,
                        int64_t ldy, int64_t y_bstride, int* done_flags, int done_epoch, void* stream);
int fv_gemm_out_norm_flow(int64_t M, int N, int K, const void* A, int64_t lda, const void* W, int64_t ldw,
                          const float* res_in, int64_t ldr, float* res_out, const float* norm_w, float eps, void* Y,
                          int64_t ldy, int* sync, int rows_per_flag, int launch_index, void* stream);
/* General form for the backward GEMMs and ragged shapes (csrc/gemm_tc2.cu): C (Mo x No) = op(A) . op(B)^T over a
 * reduction of length K, bf16 operands, fp32 accumulation in tensor memory.
 *   a_mn = 0: A stored (Mo x K) row-major;  a_mn = 1: A stored (K x Mo) row-major (used transposed, not copied)
 *   b_mn = 0: B stored (No x K) row-major;  b_mn = 1: B stored (K x No) row-major
 *   out_dtype FV_BF16: C (Mo x No) bf16, splits = 1;  FV_F32: C = `splits` fp32 planes of (Mo x ldc), plane s holding
 *   the partial sum over its own K range (add them with fv_reduce_planes; splits = 1 writes the result itself);
 *   FV_F32_ACC: one fp32 plane that every split adds into (no workspace, no second pass; not bit-reproducible for > 2 splits).
 * Replaces the cuBLAS calls of the reference's backward (selective_scan_interface.py:698-737 and autograd through
 * in_proj / out_proj): dgrad dX = dY . W (a_mn 0, b_mn 1), wgrad dW = dY^T . X (a_mn 1, b_mn 1, fp32 planes), and the
 * x_proj GEMM with N = dt_rank + 2 d_state (mamba_simple_faster.py:321-323).  Sizes need no padding: TMA zero-fills
 * out-of-range reads and clips writes; row pitches must be multiples of 16 bytes, bases 16-byte aligned.
 * fv_gemm_bf16_splits() returns the split count that fills the SMs for a shape. */
int fv_gemm_bf16_splits(int64_t Mo, int No, int64_t K);
int fv_gemm_bf16(int64_t Mo, int No, int64_t K, int a_mn, const void* A, int64_t lda, int b_mn, const void* B,
                 int64_t ldb, int out_dtype, void* C, int64_t ldc, int splits, void* stream);
/* nbatch independent products in one launch (operand / result planes a_bs / b_bs / c_bs elements apart, 16-byte multiples;
 * split-K only with FV_F32_ACC): the two directions of x_proj and of the x_proj / dt_proj backward. */
int fv_gemm_bf16_batched(int nbatch, int64_t Mo, int No, int64_t K, int a_mn, const void* A, int64_t lda, int64_t a_bs,
                         int b_mn, const void* B, int64_t ldb, int64_t b_bs, int out_dtype, void* C, int64_t ldc,
                         int64_t c_bs, int splits, void* stream);
/* ---- operator API helpers on (batch, dim, L), L contiguous ---------------------------------
 * The reference's fused autograd functions (selective_scan_interface.py:208-330, 452-605) call, on
 * (B, D, L) tensors: causal_conv1d_cuda.causal_conv1d_fwd(x, w, bias, None, True) (:496-498; third-party
 * causal-conv1d 1.1.3), conv1d_out.reshape(pre_x_shape).mean(3) (:503-508), selective_scan_cuda.fwd, and
 * out.repeat_interleave(num_of_col, 2) + D * conv1d_out (:570-571).  These serve that API in the same layout.
 *   x (batch, dim, L) with strides (x_bstride, x_dstride, 1); w (dim, 4) fp32; bias (dim) fp32 or NULL
 *   out, xc (batch, dim, L) contiguous; pooled tensors (batch, dim, outer*inner) contiguous
 */
int fv_causal_conv1d_fwd(int dtype, int batch, int dim, int64_t L, const void* x, int64_t x_bstride,
                         int64_t x_dstride, const float* w, const float* bias, int silu, void* out,
                         void* stream);
int fv_pool_bdl_fwd(int dtype, int batch, int dim, int outer, int pool, int inner, const void* x,
                    int pool_mode, float scale, void* out, void* stream);
int fv_bcast_skip_bdl_fwd(int dtype, int batch, int dim, int outer, int pool, int inner, const void* s,
                          const void* xc, const float* Dskip, void* out, void* stream);

/* ---- operator API, backward ------------------------------------------------------------------
 * fv_selective_scan_bwd replaces selective_scan_cuda.bwd (selective_scan.cpp:338-492; kernel
 * selective_scan_bwd_kernel.cuh:75-489) behind SelectiveScanFn.backward (selective_scan_interface.py:59-102).
 * Same tensors as fv_selective_scan_fwd plus dout (batch, dim, L).  Outputs: du, ddelta, dz (batch, dim, L) dtype
 * (dz NULL iff z NULL); dA (dim, N), dD (dim), d_delta_bias (dim) fp32 and dB, dC (batch, groups, N, L) FP32 --
 * all five ACCUMULATED (caller zero-fills; the caller casts dB / dC to the input dtype as the reference does,
 * selective_scan.cpp:470-472).  ddelta is the gradient of the delta INPUT (before bias and softplus).
 * No forward checkpoint is needed: the kernel re-runs the recurrence and keeps the states entering each
 * 128-step chunk in `workspace` (fv_selective_scan_bwd_workspace_bytes; 0 when L <= 128).
 * fv_causal_conv1d_bwd replaces causal_conv1d_cuda.causal_conv1d_bwd (call site :751-753): dx strided like x,
 * dw (dim, 4) / dbias (dim) fp32 accumulated.  fv_rowdot_bdl: out[d] += sum_{b,l} a[b,d,l] c[b,d,l] (the D-skip
 * gradient of the fused functions, :640-642). */
int64_t fv_selective_scan_bwd_workspace_bytes(int batch, int dim, int64_t L, int dstate);
int fv_selective_scan_bwd(int dtype, int batch, int dim, int64_t L, int dstate, int groups,
                          const void* u, const void* delta, const float* A, const void* B,
                          const void* C, const float* D, const void* z, const float* delta_bias,
                          int delta_softplus, const void* dout, void* du, void* ddelta, float* dA,
                          float* dB, float* dC, float* dD, void* dz, float* d_delta_bias,
                          void* workspace, int64_t workspace_bytes, void* stream);
int fv_causal_conv1d_bwd(int dtype, int batch, int dim, int64_t L, const void* x, int64_t x_bstride,
                         int64_t x_dstride, const float* w, const float* bias, int silu, const void* dout,
                         void* dx, int64_t dx_bstride, int64_t dx_dstride, float* dw, float* dbias,
                         void* stream);
int fv_rowdot_bdl(int dtype, int batch, int dim, int64_t L, const void* a, const void* c, float* out,
                  void* stream);

/* ======================= backward (training) entry points ==============================
 * Replace SelectiveScanFn.backward / selective_scan_cuda.bwd (selective_scan_interface.py:59-102,
 * csrc/selective_scan/selective_scan.cpp:338-492), FastVim_MambaInnerFnNoOutProj_withoutZ.backward
 * (selective_scan_interface.py:605-776), causal_conv1d_bwd (:751-753) and the Triton
 * _layer_norm_bwd_kernel (ops/triton/layernorm.py:210-305).  Plain (outer, pool, 1) geometries,
 * mean pooling.  Parameter-gradient outputs are ACCUMULATED (caller zero-fills).
 */

/* number of tiles the backward kernels cut one pooled group into (= planes of ds) */
int fv_bwd_tiles_per_group(const fv_geom* g, int dtype);

/* K2b-bwd.  dy (B, L, dim) dtype, row stride lddy.  Outputs: dz (same layout/strides as z: the z half
 * of the d(xz) buffer), e (B, L, dim) dtype contiguous = dL/dv / 2, ds_planes (tiles_per_group, B, Lp, dim)
 * fp32, dDskip (2, dim), dln_w, dln_b (dim) fp32 accumulated. */
int fv_gate_bwd(const fv_geom* g, int dtype, const void* x, const void* z, int64_t ldxz, int64_t xz_bstride,
                const void* dy, int64_t lddy, int64_t dy_bstride, const float* s, const float* conv_w,
                const float* conv_b, const float* Dskip, const float* ln_w, const float* ln_b, float eps,
                void* dz, void* e_out, float* ds_planes, float* dDskip, float* dln_w, float* dln_b,
                void* stream);

/* K2b-bwd, streaming form (csrc/gate_bwd_stream.cu): same inputs and outputs as fv_gate_bwd, but LayerNorm's coupling of
 * the channels is carried by four per-token sums in `stats` (B, L, 4) fp32 (caller zero-fills; NULL without LayerNorm), so
 * both passes are channel-local and stream (two channels per thread, register windows).  `ds` is ONE plane (B, Lp, dim)
 * fp32, ACCUMULATED (caller zero-fills).  fv_gate_bwd_stream_supported: plain geometry, dim % 64 == 0, even strides. */
int fv_gate_bwd_stream_supported(const fv_geom* g, int64_t ldxz, int64_t lddy);
int fv_gate_bwd_stream(const fv_geom* g, int dtype, const void* x, const void* z, int64_t ldxz, int64_t xz_bstride,
                       const void* dy, int64_t lddy, int64_t dy_bstride, const float* s, const float* conv_w,
                       const float* conv_b, const float* Dskip, const float* ln_w, const float* ln_b, float eps,
                       float* stats, void* dz, void* e_out, float* ds, float* dDskip, float* dln_w, float* dln_b,
                       void* stream);

/* K2a-bwd.  ds (nplanes_ds, B, Lp, dim) fp32 (summed on load; the same gradient feeds both directions).
 * Outputs: du, ddelta (2, B, Lp, dim) dtype (ddelta = gradient of the dt_proj pre-activation);
 * dbc_planes (fv_scan_bwd_planes(g), 2, B*Lp, 2N) fp32 partial sums over channel columns of [dB | dC]
 * (32-channel columns for Lp <= 16, 128-channel columns otherwise; add them with fv_reduce_planes); dA (2, dim, N) (gradient of A_log if a_is_log), d_dt_bias (2, dim)
 * fp32 accumulated. */
int fv_scan_bwd_planes(const fv_geom* g);
/* Same outputs for short pooled sequences (Lp <= 16: every 224^2 model), with the dt_proj GEMM done by the caller:
 * delta_pre (2, B, Lp, dim) fp32 = dt_bias + W_dt . dt  (the reference runs this projection as a GEMM too,
 * mamba_simple_faster.py:328-334).  Two states per thread, packed f32x2 arithmetic; dbc_planes has
 * fv_scan_bwd_planes(g) planes. */
int fv_scan_bwd_short_supported(const fv_geom* g, int dstate);
int fv_scan_bwd_short(const fv_geom* g, int dtype, int nplanes_ds, const void* u, const void* xdbl,
                      int64_t ld_xdbl, int dt_rank, int dstate, const float* delta_pre, const float* A,
                      int a_is_log, const float* ds, void* du, void* ddelta, float* dbc_planes, float* dA,
                      float* d_dt_bias, void* stream);
int fv_scan_bwd(const fv_geom* g, int dtype, int nplanes_ds, const void* u, const void* xdbl,
                int64_t ld_xdbl, int dt_rank, int dstate, const float* dt_w, const float* dt_bias,
                const float* A, int a_is_log, const float* ds, void* du, void* ddelta,
                float* dbc_planes, float* dA, float* d_dt_bias, void* stream);

/* out[i] = sum_p in[p*n + i], out dtype FV_F32 | FV_BF16 */
int fv_reduce_planes(int out_dtype, const float* in, int nplanes, int64_t n, void* out, void* stream);

/* K1-bwd.  e from fv_gate_bwd, du (2, B, Lp, dim) dtype = total gradient of the pooled conv outputs.
 * Outputs: dx (same layout/strides as x: the x half of the d(xz) buffer), dconv_w (2, dim, 4),
 * dconv_b (2, dim) fp32 accumulated; optionally dDskip (2, dim) fp32 accumulated = sum e * xc_{f,b} (pass it when
 * e comes from fv_gate_bwd_v, which does not form the D-skip gradients; NULL otherwise). */
int fv_conv_pool_bwd(const fv_geom* g, int dtype, const void* x, int64_t ldx, int64_t x_bstride,
                     const void* e, const void* du, const float* conv_w, const float* conv_b,
                     const float* Dskip, float scale, int pool_mode, void* dx, float* dconv_w,
                     float* dconv_b, float* dDskip, void* stream);

/* K2b-bwd, streaming form: backward of (LayerNorm + SiLU gate) from the pre-norm value v saved by fv_block_fwd (v_out).
 * Replaces autograd through mamba_simple_faster.py:434-453.  v, e, dy-independent layouts as fv_block_fwd: v and e are
 * (B, L, dim) bf16 in memory token order, z / dz rows have stride ldz (the z half of the (d)xz buffer), dy has lddy.
 * Outputs: dz, e = dv / 2 (consumed by fv_conv_pool_bwd), ds (B, Lp, dim) fp32 = sum of e over each pooled row (ONE
 * plane, valid for both scan directions), dln_w / dln_b (dim) fp32 accumulated.  bf16, plain geometry,
 * dim in {384, 768, 1536, 3072}: fv_gate_bwd_v_supported(). */
int fv_gate_bwd_v_supported(const fv_geom* g, int dtype);
int fv_gate_bwd_v(const fv_geom* g, int dtype, const void* v, const void* z, int64_t ldz, int64_t z_bstride,
                  const void* dy, int64_t lddy, int64_t dy_bstride, const float* ln_w, const float* ln_b, float eps,
                  void* dz, void* e, float* ds, float* dln_w, float* dln_b, void* stream);

/* add + norm backward.  residual_out is the fp32 sum saved by the forward; dresidual_out (grad of that
 * output) may be NULL.  dx (dtype) and/or dresidual_in (fp32) receive the same gradient. */
int fv_add_norm_bwd(int dtype, int64_t rows, int cols, const void* dy, int64_t lddy,
                    const float* dresidual_out, const float* residual_out, const float* weight,
                    float eps, int is_rms, void* dx, int64_t lddx, float* dresidual_in, float* dweight,
                    float* dbias, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FASTVIM_B200_H_ */
