/*
 * tamtr_b200 -- C ABI of the B200-native TAM-TR detection-head hot path (sm_100a only, no CPU fallback).
 *
 * This is the drop-in boundary: every entry point takes plain device pointers, sizes and a CUDA stream
 * (no torch types), so any host language can bind it (ctypes binding: tamtr_b200/_lib.py; the reference-side
 * binding a maintainer would add is shown in INTEGRATION.md).  Each function cites the reference interface
 * (file:line under the TAM-TR repo) whose arithmetic it replaces.
 *
 * Conventions
 *   - all tensors are contiguous, row-major, on the CURRENT device of the calling thread;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls only enqueue work;
 *   - return value: 0 on success, a positive cudaError_t, or a negative TAMTR_E_* for argument errors;
 *     tamtr_last_error() returns a thread-local message for the last non-zero return;
 *   - inputs are borrowed for the duration of the enqueued work, outputs are caller-allocated.
 */
#ifndef TAMTR_B200_H
#define TAMTR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TAMTR_B200_ABI_VERSION 2

enum tamtr_dtype { TAMTR_F32 = 0, TAMTR_BF16 = 1 };

enum tamtr_error {
    TAMTR_E_BADARG = -1,       /* null pointer / non-positive size */
    TAMTR_E_UNSUPPORTED = -2,  /* shape outside the compiled template set (see tamtr_last_error) */
    TAMTR_E_NODEVICE = -3      /* no sm_100 device visible */
};

int tamtr_abi_version(void);
const char *tamtr_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
unsigned long long tamtr_launch_count(void);

/* Per-kernel device timing for bench.py's roofline: when enabled, every kernel launch of this library is bracketed by
 * CUDA events on its launching stream (skipped while that stream is being captured into a CUDA graph).
 * tamtr_profile_enable(on) also clears what was recorded; tamtr_profile_read synchronises and sums. */
enum tamtr_kernel {
    TAMTR_K_MSDA_FWD = 0, TAMTR_K_MSDA_BWD, TAMTR_K_LOCW_FWD, TAMTR_K_LOCW_BWD, TAMTR_K_CONTRASTIVE_FWD,
    TAMTR_K_CONTRASTIVE_BWD, TAMTR_K_MAX_SIGMOID_FWD, TAMTR_K_MAX_SIGMOID_BWD, TAMTR_K_MAX_SIGMOID_TC_FWD,
    TAMTR_K_COUNT
};
/* cudaMemsetAsync(ptr, 0, bytes) on `stream` (a memset node is ~1.5x faster than a fill kernel for GB-sized buffers) */
int tamtr_memset_zero(void *ptr, unsigned long long bytes, void *stream);
/* Zero `bytes` at `ptr` (16-byte aligned) with a kernel of `n_ctas` small CTAs (0: one per SM; ~62 GB/s each) issuing bulk shared->global
 * stores (n_ctas < 0: -n_ctas CTAs storing from registers, no shared memory at all): meant to run on a side stream BESIDE latency-bound work (the gradient arena of the samplers' backward,
 * zeroed during the decoder forward; replaces torch.zeros_like in the reference's autograd of transformer.py:273),
 * where a full-grid memset node would take every SM and the whole memory system. */
int tamtr_zero_fill_background(void *ptr, unsigned long long bytes, int n_ctas, void *stream);
int tamtr_profile_enable(int on);
int tamtr_profile_read(int kernel_id, double *total_ms, unsigned long long *launches);
const char *tamtr_kernel_name(int kernel_id);

/* ---------------------------------------------------------------------------------------------------------
 * Multi-scale deformable attention core op.
 * Replaces ultralytics/nn/modules/utils.py:42-89 multi_scale_deformable_attn_pytorch (3x F.grid_sample +
 * stack * weights .sum), i.e. ATen grid_sampler_2d bilinear / zeros padding / align_corners=False
 * (torch/include/ATen/native/GridSampler.h:27-36, 205-207).
 *
 *   value  [B, Lv, H, Dh]  f32 or bf16 (dtype)    level l = tokens [start_l, start_l + H_l*W_l), row-major
 *   loc    [B, Lq, H, L, P, 2] f32  (x, y) normalised to [0,1]; may leave [0,1] (zeros padding)
 *   attn   [B, Lq, H, L, P]    f32
 *   out    [B, Lq, H*Dh]       same dtype as value
 *   level_shapes_host [L][2] = (H_l, W_l), HOST memory, read during the call (utils.py:56 value_spatial_shapes)
 *   value_token_stride: elements between consecutive tokens of `value` (0 = H*Dh, contiguous).  A larger stride
 *     lets one GEMM project the values of all decoder layers at once ([B, Lv, n_layers*H*Dh]; transformer.py:273 is
 *     called per layer on the same `feats`, transformer.py:870) and each layer sample its own column slice.
 * Supported: Dh*sizeof(elt) in {32,64,128,256} bytes, L*P <= 32, sum(H_l*W_l) == Lv.
 * Index math is bit-exact with the reference: ix = fmaf((2*loc-1)+1, W_l, -1) * 0.5, floor, 4 corners, each
 * corner contributes iff 0<=x<W_l && 0<=y<H_l.  Accumulation in fp32.
 */
int tamtr_msda_forward(const void *value, const float *loc, const float *attn, void *out, int dtype,
                       int B, int Lv, int H, int Dh, int Lq, int L, int P,
                       const int32_t *level_shapes_host, int value_token_stride, void *stream);

/* Backward of the above (what autograd derives for utils.py:42-89: grid_sampler_2d_backward + mul/sum).
 *   grad_out   [B, Lq, H*Dh]   same dtype as value
 *   grad_value [B, Lv, H, Dh]  grad_value_dtype: the value dtype, or TAMTR_F32 for bf16 values (fp32 accumulation of
 *                              the scattered gradient: each atomic add then rounds at 2^-24 instead of 2^-9 of the
 *                              running sum); same token stride (in elements) as value; zeroed by this call when
 *                              zero_grad_value != 0
 *                              (otherwise the caller zeroed the whole strided buffer once for all layers), then
 *                              accumulated with vector atomics (REDG f32x4 / bf16x8) -> run-to-run bit differences,
 *                              like grid_sampler backward
 *   grad_loc   [B, Lq, H, L, P, 2] f32, grad_attn [B, Lq, H, L, P] f32 (fully overwritten)
 *   tap_weight_sum [B, Lq, H] f32 or NULL: sum of the in-bounds A*w_k of each (query, head); the column sums of
 *                  grad_value (= value_proj's bias gradient) are sum_q tap_weight_sum[q,h] * grad_out[q,h,:]
 */
int tamtr_msda_backward(const void *grad_out, const void *value, const float *loc, const float *attn,
                        void *grad_value, float *grad_loc, float *grad_attn, int dtype,
                        int B, int Lv, int H, int Dh, int Lq, int L, int P,
                        const int32_t *level_shapes_host, int value_token_stride, int zero_grad_value,
                        float *tap_weight_sum, int grad_value_dtype, void *stream);

/* The same sampler with a different number of points on each level: the reference's "decoupled" cross-attention pair
 * multi_scale_deformable_attn_pytorch_cls (ultralytics/nn/modules/utils.py:92-140, points 2/4/6 on the three levels) and
 * multi_scale_deformable_attn_pytorch_box (utils.py:143-191, points 6/4/2), called from MSDeformAttncls / MSDeformAttnbox
 * (nn/modules/transformer.py:396,494).
 *   points_host [L] int32, HOST memory: points of level l; samples are ordered level by level, exactly like the
 *                 reference's torch.split(sampling_grids, points, dim=-2) (utils.py:108,159)
 *   loc  [B, Lq, H, S, 2] f32, attn [B, Lq, H, S] f32 (the reference's [B,Lq,H,L,P] view with L*P == S), S = sum(points) <= 32
 * Everything else (value layout, token stride, index-math contract, gradients) as in tamtr_msda_forward / _backward. */
int tamtr_msda_forward_ragged(const void *value, const float *loc, const float *attn, void *out, int dtype,
                              int B, int Lv, int H, int Dh, int Lq, int L, const int32_t *points_host,
                              const int32_t *level_shapes_host, int value_token_stride, void *stream);
int tamtr_msda_backward_ragged(const void *grad_out, const void *value, const float *loc, const float *attn,
                               void *grad_value, float *grad_loc, float *grad_attn, int dtype,
                               int B, int Lv, int H, int Dh, int Lq, int L, const int32_t *points_host,
                               const int32_t *level_shapes_host, int value_token_stride, int zero_grad_value,
                               float *tap_weight_sum, int grad_value_dtype, void *stream);
int tamtr_msda_corners_ragged(const float *loc, int32_t *x0, int32_t *y0, uint8_t *inb, int B, int Lq, int H, int L,
                              const int32_t *points_host, const int32_t *level_shapes_host, void *stream);

/* Parity export of the index math alone (the "bit-exact sampling-location indexing" object):
 *   x0, y0 [B,Lq,H,L,P] int32 = floor(ix), floor(iy);  inb [B,Lq,H,L,P,4] uint8 = in-bounds flags (nw,ne,sw,se).
 * Runs the same __device__ function the two kernels above use. */
int tamtr_msda_corners(const float *loc, int32_t *x0, int32_t *y0, uint8_t *inb,
                       int B, int Lq, int H, int L, int P, const int32_t *level_shapes_host, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Sampling locations + attention weights from the fused offset/weight projection.
 * Replaces ultralytics/nn/modules/transformer.py:278-293 (bias add of the two nn.Linear, softmax over L*P,
 * loc = ref_xy + off / P * ref_wh * 0.5  [RD == 4]   or   loc = ref_xy + off / (W_l, H_l)  [RD == 2]).
 *   raw  [M, 3*H*L*P] f32   GEMM output of query x concat(sampling_offsets.weight, attention_weights.weight)^T
 *                           columns [0, 2*H*L*P) = offsets (h,l,p,xy), then H*L*P logits (h, l*P+p)
 *   bias [3*H*L*P]    f32   concat(sampling_offsets.bias, attention_weights.bias)
 *   ref  [M, RL, RD]  f32   RL in {1, L}
 *   loc  [M, H, L, P, 2], attn [M, H, L, P]  f32
 * The location arithmetic keeps the reference's op order and per-op fp32 rounding (no FMA contraction).
 * level_shapes_host may be NULL when RD == 4.
 */
int tamtr_locw_forward(const float *raw, const float *bias, const float *ref, float *loc, float *attn,
                       int M, int H, int L, int P, int RL, int RD, const int32_t *level_shapes_host, void *stream);

/* The same two Linears (ultralytics/nn/modules/transformer.py:278-279) AND the epilogue above as ONE tensor-core kernel
 * for bf16 activations: TMA-staged query / weight tiles, tcgen05.mma into TMEM (fp32 accumulate), the softmax and the
 * location arithmetic applied to each query's TMEM row; the [M, 3*H*L*P] GEMM result never goes through HBM.
 *   query_bf16 [M, C] bf16 row-major; w_off_bf16 [2*H*L*P, C], w_attn_bf16 [H*L*P, C] bf16 = sampling_offsets.weight,
 *   attention_weights.weight as they are (no concatenation); b_off [2*H*L*P], b_attn [H*L*P] f32 or bf16 (bias_dtype)
 *   raw [M, 3*H*L*P] f32 or NULL: the pre-bias GEMM result, only needed by tamtr_locw_backward for grad_ref
 * The H heads of a 128-query tile are split over up to 4 CTAs so that the launch covers the SMs.
 * tamtr_locw_tc_supported() tells whether the problem fits (L*P in {12, 16}, C % 64 == 0, a head split whose 3*Hc*L*P
 * columns are <= 512 and divisible into equal chunks <= 256 that are multiples of 16, RL == 1, RD == 4); otherwise
 * callers use a library GEMM + tamtr_locw_forward. */
int tamtr_locw_tc_supported(int M, int C, int H, int L, int P, int RL, int RD);
int tamtr_locw_tc_forward(const void *query_bf16, const void *w_off_bf16, const void *w_attn_bf16, const void *b_off,
                          const void *b_attn, int bias_dtype, const float *ref, float *loc, float *attn, float *raw,
                          int M, int C, int H, int L, int P, int RL, int RD, void *stream);

/* Backward: grad_raw [M, 3*H*L*P] (gradient w.r.t. the GEMM output, = per-row bias gradient) is fully written;
 * grad_ref [M, RL, RD] may be NULL (reference boxes are detached in training, transformer.py:889); `raw` and `bias`
 * may be NULL when grad_ref is NULL. */
int tamtr_locw_backward(const float *grad_loc, const float *grad_attn, const float *attn, const float *raw,
                        const float *bias, const float *ref, float *grad_raw, float *grad_ref,
                        int M, int H, int L, int P, int RL, int RD, const int32_t *level_shapes_host, void *stream);

/* Iterative box refinement of the decoders (transformer.py:875,882 / 699,706; inverse_sigmoid: utils.py:34-39):
 *   out = sigmoid(bbox + log(clamp(ref,0,1).clamp(min=eps) / (1 - clamp(ref,0,1)).clamp(min=eps)))     all f32 [n]
 * backward: grad_bbox always, grad_ref optional (NULL when the reference boxes are detached). */
int tamtr_box_refine_forward(const float *bbox, const float *ref, float *out, int n, float eps, void *stream);
int tamtr_box_refine_backward(const float *grad_out, const float *out, const float *ref, float *grad_bbox,
                              float *grad_ref, int n, float eps, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Text-guided classification branch (region-text contrastive head).
 * Replaces ultralytics/nn/modules/block.py:534-541 ContrastiveHeadMLP.forward:
 *   out[b,q,k] = <x[b,q,:]/max(|x|,1e-12), w[b,k,:]/max(|w|,1e-12)> * exp(*logit_scale) + *bias
 *   x [B, Lq, C] f32|bf16 (dtype), w [B, K, C] f32, logit_scale / bias: DEVICE scalars, out [B, Lq, K] f32.
 * Supported: C % 128 == 0, C <= 1024, K <= 256.
 */
int tamtr_contrastive_forward(const void *x, const float *w, const float *logit_scale, const float *bias,
                              float *out, int dtype, int B, int Lq, int K, int C, void *stream);

/* grad_x [B, Lq, C] (dtype of x); grad_scalars[2] = {d/d logit_scale, d/d bias} (zeroed by the call). */
int tamtr_contrastive_backward(const float *grad_out, const void *x, const float *w, const float *logit_scale,
                               void *grad_x, float *grad_scalars, int dtype, int B, int Lq, int K, int C,
                               void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * BTA-PAN text-image attention gate (max-sigmoid attention), exact-fp32 CUDA-core path.
 * Replaces ultralytics/nn/extra_modules/block.py:216-220:
 *   aw[b,m,pix] = sigmoid( max_n <embed[b, m*hc:(m+1)*hc, pix], guide[b,n,m,:]> / sqrt(hc) + bias[m] )
 *   embed [B, nh*hc, HW] f32|bf16 (NCHW), guide [B, N, nh, hc] f32 (= gl(guide) of block.py:212-213), bias [nh]
 *   aw [B, nh, HW] f32, amax [B, nh, HW] uint8 (arg max over n, input of the backward)
 * Supported: hc in {16, 32, 64}, N <= 255.
 */
int tamtr_max_sigmoid_forward(const void *embed, const float *guide, const float *bias, float *aw, uint8_t *amax,
                              int dtype, int B, int nh, int hc, int HW, int N, void *stream);

/* Same gate on the tensor cores for bf16 activations: tcgen05.mma (M=128 pixels, N=pad16(N), K=hc=32) with TMEM
 * accumulators, TMA-staged 128B-swizzled tiles of `embed` (MN-major A operand straight from NCHW), warp-specialised
 * producer / MMA / epilogue.  embed bf16 [B, nh*32, HW] (16-byte aligned, HW % 8 == 0), guide f32 [B,N,nh,32]
 * (rounded to bf16 inside), N <= 128.  Same outputs as tamtr_max_sigmoid_forward; the backward is shared. */
int tamtr_max_sigmoid_tc_forward(const void *embed_bf16, const float *guide, const float *bias, float *aw,
                                 uint8_t *amax, int B, int nh, int hc, int HW, int N, void *stream);

/* The gated 3x3 projection of the same block (ultralytics/nn/extra_modules/block.py:222-225: proj_conv = Conv2d 3x3 pad 1
 * + BatchNorm2d, then `x * aw.unsqueeze(2)`) as ONE tensor-core kernel: implicit GEMM (M = pixels, N = Cout, K = 9*Cin),
 * tcgen05.mma with TMEM accumulators, TMA boxes whose out-of-bounds zero fill is the conv padding, epilogue
 *   y[b,h,w,co] = (conv[b,h,w,co] * bn_scale[co] + bn_shift[co]) * gate[b, co / (Cout/nh), h, w]     (gate may be NULL).
 * Channels-last operands: x [B,H,W,Cin] bf16, w [Cout,3,3,Cin] bf16 (= weight.permute(0,2,3,1)), y [B,H,W,Cout] bf16,
 * gate f32 [B,nh,H,W] (the `aw` of tamtr_max_sigmoid_*_forward), bn_scale/bn_shift f32 [Cout] (the BatchNorm affine:
 * gamma*rstd, beta - mean*gamma*rstd).  Cin % 64 == 0, Cout % 32 == 0, Cout <= 256, (Cout/nh) % 32 == 0. */
int tamtr_gate_conv3x3_tc_forward(const void *x_nhwc, const void *w_ohwi, const float *bn_scale, const float *bn_shift,
                                  const float *gate, void *y_nhwc, int B, int H, int W, int Cin, int Cout, int nh,
                                  void *stream);

/* [B, C, HW] -> [B, HW, C] (f32 | bf16): the layout change in front of the channels-last tensor-core kernel when the
 * caller holds NCHW maps (the reference's layout). */
int tamtr_nchw_to_nhwc(const void *x, void *y, int dtype, int B, int C, int HW, void *stream);

/* grad_embed [B, nh*hc, HW] (dtype of embed, fully written: the gate's contribution only);
 * grad_guide [B, N, nh, hc] f32 and grad_bias [nh] f32 are zeroed by the call, then accumulated. */
int tamtr_max_sigmoid_backward(const float *grad_aw, const float *aw, const uint8_t *amax, const void *embed,
                               const float *guide, void *grad_embed, float *grad_guide, float *grad_bias,
                               int dtype, int B, int nh, int hc, int HW, int N, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Encoder-side token kernels (everything on the [B, Lv, d] token tensor that is not a GEMM).
 *
 * Token-major BatchNorm of `input_proj` (ultralytics/nn/modules/head.py:1202-1218: conv1x1 + BatchNorm2d in NCHW, then
 * flatten(2).permute(0,2,1) + cat).  The 1x1 conv is a GEMM that writes [B, HW_l, d] into the level's slice of the
 * token tensor; the statistics and both normalisation passes are:
 *   tamtr_col_reduce2: partial[cta][0][c] = sum_rows a[r][c], partial[cta][1][c] = sum_rows a[r][c]*b[r][c] over the
 *                      tokens [tok0, tok0+ntok) of every image; n_cta = tamtr_col_reduce2_ctas(B, ntok); fp32 sums
 *   tamtr_affine_rows: out[b,t,c] = A[l][c]*a[b,t,c] + Bc[l][c]*b[b,t,c] + Cc[l][c], l = level of token t
 *                      (b / Bc may be NULL); level_starts_host = first token of each of the L levels
 * a, b, out: [B, Lv, d] f32|bf16 (dtype); A, Bc, Cc: [L, d] f32.  d % (16/sizeof) == 0, d <= 1024.
 */
int tamtr_col_reduce2_ctas(int B, int ntok);
int tamtr_col_reduce2(const void *a, const void *b, float *partial, int dtype, int B, int Lv, int d, int tok0, int ntok,
                      void *stream);
int tamtr_affine_rows(void *out, const void *a, const void *b, const float *A, const float *Bc, const float *Cc,
                      int dtype, int B, int Lv, int d, int L, const int32_t *level_starts_host, void *stream);

/* BatchNorm bookkeeping for one level between the two passes above (one thread per channel, fp64 inside):
 *   forward : from the partials of tamtr_col_reduce2(pre, pre) over M = B*ntok rows (use_batch_stats) or from the running
 *             statistics: scale = gamma*rstd, shift = beta - mu*scale (the A / Cc rows of tamtr_affine_rows), mu and
 *             rstd (f64 [d]) for the backward; with batch statistics and running_mean != NULL the running statistics
 *             are updated like nn.BatchNorm2d (momentum, unbiased variance).
 *   backward: from the partials of tamtr_col_reduce2(G, pre): A, Bc, Cc rows of the backward affine pass, d_gamma,
 *             d_beta (f32 [d]).  batch_stats = 0 (eval-mode statistics) gives Bc = Cc = 0. */
int tamtr_bn_forward_coeffs(const float *partial, int n_cta, double M, const float *gamma, const float *beta, double eps,
                            int use_batch_stats, double momentum, float *running_mean, float *running_var, float *scale,
                            float *shift, double *mu, double *rstd, int d, void *stream);
int tamtr_bn_backward_coeffs(const float *partial, int n_cta, double M, const float *gamma, const double *mu,
                             const double *rstd, int batch_stats, float *A, float *Bc, float *Cc, float *d_gamma,
                             float *d_beta, int d, void *stream);

/* out[c] = sum_r g[r][c]: bias gradient of the decoder's Linear layers (g: [rows, n] f32|bf16, n % (16/sizeof) == 0,
 * out f32 [n], zeroed by the call and accumulated with one fp32 reduction per column and CTA). */
int tamtr_col_sum(const void *g, float *out, int dtype, int rows, int n, void *stream);

/* Query-selection ranking (head.py:1229-1237: enc_output = Linear + LayerNorm over all tokens, enc_score_head, max over
 * classes), fused after the two GEMMs:
 *   E   [B*Lv, d] f32|bf16 = feats @ enc_output.0.weight^T (no bias, no validity mask)
 *   raw [B*Lv, raw_stride >= nc] f32 = E @ (enc_score_head.weight * ln.weight)^T   (columns padded so that the
 *                            skinny GEMM keeps 16-byte aligned operands)
 *   v = valid[t] ? E + enc_bias : enc_bias; (mean, rstd) = LayerNorm statistics of v (eps)
 *   out[row] = max_k rstd * (raw[k] + bw[k] - mean*sw[k]) + ck[k]        (raw treated as 0 for invalid tokens)
 *   bw[k] = enc_bias . W'[k], sw[k] = sum_c W'[k][c], ck[k] = ln.bias . score_w[k] + score_b[k]
 */
int tamtr_rank_tokens(const void *E, const float *raw, const float *enc_bias, const uint8_t *valid, const float *bw,
                      const float *sw, const float *ck, float *out, int dtype, int B, int Lv, int d, int nc,
                      int raw_stride, float eps, void *stream);

/* Query selection proper (head.py:1240 and :437: `torch.topk(scores, num_queries, dim=1).indices` over the ranking scores
 * of tamtr_rank_tokens / tamtr_tok_project_rank): scores [rows, n] f32 -> out_idx [rows, k] int64 (and, unless NULL,
 * out_val [rows, k] f32), best first; equal scores in order of their index (torch leaves that order unspecified).
 * One CTA per row: order-preserving 32-bit keys in shared memory (re-read from global memory when the row does not
 * fit), 12 + 12 + 8 bit radix select over the few keys that can still be winners, one ordered collection pass, ranking by
 * counting.  1 <= k <= min(n, 4096).  _supported: 0 = no, 2 = yes with the row in shared memory (n up to ~47 000: 24 us for
 * 16 x 33 600, k = 300, the library's ~10 launches take 70-140 us), 1 = yes but re-reading the row from L2 on every pass
 * (slower than the library on long rows; callers use it only for its tie order). */
int tamtr_topk_rows_supported(int n, int k);
int tamtr_topk_rows(const float *scores, long long *out_idx, float *out_val, int rows, int n, int k, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Residual add + LayerNorm of the decoder layers (ultralytics/nn/modules/transformer.py:548,553,537:
 * `embed = self.normN(embed + self.dropoutN(tgt))`, dropout p = 0) -- one kernel forward, one backward.
 *   forward : z = x + res (res may be NULL), y = (z - mean) * rstd * w + b over the last dimension d.
 *             x, res, y: [rows, d] f32|bf16 (each with its own dtype code); w, b f32 [d]; z f32 [rows, d] and
 *             mean, rstd f32 [rows] are written for the backward (any of them may be NULL at inference).
 *   backward: dx / dres (either may be NULL) = d(loss)/dz in the dtype of x / res; dwb f32 [2, d] = (dw, db), zeroed
 *             by the call (memset node) and accumulated with one fp32 reduction per column and CTA.
 * d % 128 == 0, 128 <= d <= 512. */
int tamtr_add_layernorm_forward(const void *x, int x_dtype, const void *res, int res_dtype, const float *w, const float *b,
                                void *y, int y_dtype, float *z, float *mean, float *rstd, int rows, int d, float eps,
                                void *stream);
int tamtr_add_layernorm_backward(const void *dy, int dy_dtype, const float *z, const float *mean, const float *rstd,
                                 const float *w, void *dx, int dx_dtype, void *dres, int dres_dtype, float *dwb, int rows,
                                 int d, void *stream);
/* the same with `extra` f32 [rows, d] added to d(loss)/dz: the gradient that reaches z directly through a residual
 * connection around the normalised branch (VSSBlock: x + mlp(norm2(x)), vmamba.py:1249) */
int tamtr_add_layernorm_backward_res(const void *dy, int dy_dtype, const float *z, const float *mean, const float *rstd,
                                     const float *w, const float *extra, void *dx, int dx_dtype, void *dres, int dres_dtype,
                                     float *dwb, int rows, int d, void *stream);

/* Decoder layer under bf16 autocast (transformer.py:544-557): the normalised stream stays fp32, but every consumer of it
 * (the attention projections, the FFN) computes in bf16, on `embed` or on `embed + query_pos`.  The casts and the
 * `with_pos_embed` adds autocast puts between them are side outputs / side gradients of the kernels above:
 *   forward_sides : also writes y_bf16 = bf16(y) and q_bf16 = bf16(y + pos) (pos [rows, d] f32|bf16; either output may be
 *                   NULL) -- the values autocast's casts would produce.
 *   backward_sides: dy (may be NULL) + the gradients of the two bf16 outputs (bf16 [rows, d], either may be NULL) are
 *                   summed in fp32 while they are loaded.
 *   pos_cast      : x_bf16 = bf16(x), q_bf16 = bf16(x + pos) for the layer's input (n elements, n % 4 == 0).
 *   grad_sum3     : out = a + b + c (any two may be NULL; fp32 sum): the gradient of that input, which also feeds the
 *                   residual connection. */
int tamtr_add_layernorm_forward_sides(const void *x, int x_dtype, const void *res, int res_dtype, const float *w,
                                      const float *b, void *y, int y_dtype, float *z, float *mean, float *rstd,
                                      const void *pos, int pos_dtype, void *y_bf16, void *q_bf16, int rows, int d, float eps,
                                      void *stream);
int tamtr_add_layernorm_backward_sides(const void *dy, int dy_dtype, const void *g_y_bf16, const void *g_q_bf16,
                                       const float *z, const float *mean, const float *rstd, const float *w, void *dx,
                                       int dx_dtype, void *dres, int dres_dtype, float *dwb, int rows, int d, void *stream);
int tamtr_pos_cast(const void *x, int x_dtype, const void *pos, int pos_dtype, void *x_bf16, void *q_bf16, long n,
                   void *stream);
int tamtr_grad_sum3(const void *a, int a_dtype, const void *b, int b_dtype, const void *c, int c_dtype, void *out,
                    int out_dtype, long n, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Query self-attention of the decoder layers (ultralytics/nn/modules/transformer.py:544-548: nn.MultiheadAttention with
 * q = k = embed + pos, v = embed and the denoising attention mask of models/utils/ops.py:273-284), after the packed input
 * projection -- softmax(q k^T / sqrt(Dh) + mask) v per (image, head), one fused kernel forward, two backward (csrc/selfattn.cu).
 *   qk, d_qk: bf16 [Bn, L, 2 * H * Dh] (q in the first half of every row, k in the second); v, o, d_o, d_v: bf16 [Bn, L, H * Dh];
 *   mask_bits: u64 [L, ceil(L / 64)] or NULL: the bool mask of nn.MultiheadAttention (non-zero = query row may NOT attend to
 *             key column) packed by tamtr_self_attention_pack_mask from u8 [L, L] -- bit c of word [q][t] = key 64 t + c;
 *   lse2    : f32 [Bn, H, L], log2-sum-exp of the scaled scores, written by the forward for the backward;
 *   scratch : bf16 [2, Bn, H, Lp, Lp] with Lp = tamtr_self_attention_padded_len(L) (probabilities and score gradients).
 * Dh = 32 or 64; every pointer 16-byte aligned.  A row whose keys are all blocked yields zeros (the library yields NaN). */
int tamtr_self_attention_supported(int L, int H, int Dh);
int tamtr_self_attention_padded_len(int L);
int tamtr_self_attention_mask_words(int L);
int tamtr_self_attention_pack_mask(const uint8_t *blocked, unsigned long long *mask_bits, int L, void *stream);
int tamtr_self_attention_forward(const void *qk, const void *v, const unsigned long long *mask_bits, void *o, float *lse2, int Bn,
                                 int L, int H, int Dh, void *stream);
int tamtr_self_attention_backward(const void *qk, const void *v, const unsigned long long *mask_bits, const void *o,
                                  const void *d_o, const float *lse2, void *d_qk, void *d_v, void *scratch, int Bn, int L, int H,
                                  int Dh, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Query <-> ground-truth matching on the device: scipy.optimize.linear_sum_assignment as the reference calls it per
 * image on the host (ultralytics/models/utils/ops.py:116-121), same algorithm (Crouse / Jonker-Volgenant shortest
 * augmenting path), fp64 arithmetic and tie order, one warp per (layer, image).
 *   C            f32 [n_layers, bs, nq, c_cols]  cost matrices (finite values).  padded = 0: c_cols = total number of
 *                gts and image b owns the columns [gt_start[b], gt_start[b+1]) (the reference's layout, ops.py:104-116);
 *                padded = 1: c_cols >= max_gt and image b owns the columns [0, n_gt[b]) of its own matrix
 *   gt_start_dev int32 [bs+1] (device), out_start_dev int32 [bs] (device): first output slot of image b
 *                = sum_{b' < b} min(nq, n_gt[b'])
 *   out_q, out_g int64 [n_layers, out_layer_stride]: matched (query index, GLOBAL gt index) pairs of each image in
 *                ascending query order, min(nq, n_gt[b]) pairs per image
 *   max_gt       the largest n_gt[b] (host) */
int tamtr_linear_sum_assignment(const float *C, const int *gt_start_dev, const int *out_start_dev, long long *out_q,
                                long long *out_g, int n_layers, int bs, int nq, int c_cols, int padded, int max_gt,
                                long long out_layer_stride, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Selective scan (S6) of the VSSBlocks in front of the MEH head (ultralytics/nn/modules/head.py:1092-1098,1134 ->
 * nn/extra_modules/VManba/vmamba.py:962-990 -> csms6s.py:252-270).  The reference calls the third-party extension
 * `selective_scan_cuda_core` (fwd/bwd) there; that extension is not in the reference tree, so these entry points follow
 * its call signature and the published recurrence (Gu & Dao, "Mamba", 2023):
 *   delta = softplus(dt + bias) (identity above 20); h_t = exp(delta_t*A)*h_{t-1} + delta_t*B_t*u_t; y_t = <C_t,h_t> + D*u_t
 * u, dt, g_u, g_dt: f32 | bf16 (in_dtype; bf16 needs an even L -- converted on load, the arithmetic is fp32 either way, as
 * vmamba.py:985-986 forces) [Bn, KD, L]; y, dy: f32 [Bn, KD, L]; A, g_A: f32 [KD, N]; Bm, Cm, g_B, g_C: f32 [Bn, KD/Dg, N, L];
 * D, bias, g_D, g_bias: f32 [KD] (D / bias may be NULL).  N = 16, Dg (channels per scan direction) % 32 == 0.
 * ckpt: f32 [Bn, KD, tamtr_selective_scan_segments(L), N], written by the forward (may be NULL at inference), read by the
 * backward.  g_A, g_B, g_C, g_D, g_bias are zeroed by the call and accumulated (fp32 reductions over batch / channels). */
int tamtr_selective_scan_segments(int L);
int tamtr_selective_scan_forward(const void *u, const void *dt, int in_dtype, const float *A, const float *Bm,
                                 const float *Cm, const float *D, const float *bias, float *y, float *ckpt, int Bn, int KD,
                                 int Dg, int N, int L, void *stream);
int tamtr_selective_scan_backward(const void *u, const void *dt, int in_dtype, const float *A, const float *Bm,
                                  const float *Cm, const float *D, const float *bias, const float *dy, const float *ckpt,
                                  void *g_u, void *g_dt, float *g_A, float *g_B, float *g_C, float *g_D, float *g_bias,
                                  int Bn, int KD, int Dg, int N, int L, void *stream);

/* Chunk-parallel forward for inference (no checkpoints), for (channel block, image) grids too small to fill the GPU --
 * batch 1 at 1280x1280 is 32 CTAs walking 102 400 positions.  The sequence is cut into n_chunks pieces; a state pass runs
 * the recurrence without outputs on every piece but the last from h = 0 and records its end state and sum of delta, the
 * output pass folds h_in of each piece from the pieces before it (the recurrence is linear in h and the decay over a piece
 * is exp(A * sum of its deltas)) and scans it.  Same arguments as tamtr_selective_scan_forward, plus
 *   carry f32 [Bn * KD * n_chunks * 17] workspace; n_chunks >= 2 (tamtr_selective_scan_chunks() recommends a value, 1 =
 *   use the plain forward). */
int tamtr_selective_scan_chunks(int Bn, int KD, int L);
int tamtr_selective_scan_forward_chunked(const void *u, const void *dt, int in_dtype, const float *A, const float *B,
                                         const float *C, const float *D, const float *bias, float *y, float *carry,
                                         int n_chunks, int Bn, int KD, int Dg, int N, int L, void *stream);

/* SS2D tail (ultralytics/nn/extra_modules/VManba/vmamba.py:1011-1014 out_norm, :1029-1031 gate): LayerNorm over the CHANNEL
 * dimension of a position-major tensor, times the SiLU-activated gate, one kernel each way (csrc/vssfuse.cu):
 *   out[b,c,l] = (LN_c(y[b,:,l]) * gamma[c] + beta[c]) * silu(z[b,c,l])
 * y, d_y: f32 [Bn, D, L]; z, d_z: f32 | bf16 [Bn, D, L] (the raw gate half of in_proj, before the activation); out: f32 | bf16;
 * dout: f32 | bf16; gamma, beta, d_gamma, d_beta: f32 [D]; mean, rstd: f32 [Bn, L] (written by the forward unless NULL,
 * read by the backward).  d_gamma / d_beta are zeroed by the call and accumulated in fp32. */
int tamtr_colnorm_gate_forward(const float *y, const void *z, int z_dtype, const float *gamma, const float *beta, void *out,
                               int out_dtype, float *mean, float *rstd, int Bn, int D, int L, float eps, void *stream);
int tamtr_colnorm_gate_backward(const void *dout, int dout_dtype, const float *y, const void *z, int z_dtype,
                                const float *gamma, const float *beta, const float *mean, const float *rstd, float *d_y,
                                void *d_z, float *d_gamma, float *d_beta, int Bn, int D, int L, void *stream);

/* The four scan orders of SS2D (ultralytics/nn/extra_modules/VManba/csms6s.py:4-47), one pass each way:
 *   tamtr_cross_scan : x [Bn, D, H, W] -> xs [Bn, 4, D, H*W]  (row-major, column-major, both reversed)   = CrossScan.forward
 *                                                                                                       = CrossMerge.backward
 *   tamtr_cross_merge: ys [Bn, 4, D, H*W] -> y [Bn, D, H*W] = (ys0 + flip(ys2)) + transpose(ys1 + flip(ys3))
 *                                                                              = CrossMerge.forward = CrossScan.backward
 * f32 | bf16 (same dtype in and out; bf16 sums are rounded where the reference's tensor adds round). */
int tamtr_cross_scan(const void *x, void *xs, int dtype, int Bn, int D, int H, int W, void *stream);
int tamtr_cross_merge(const void *ys, void *y, int dtype, int Bn, int D, int H, int W, void *stream);

/* Depth-wise 3x3 convolution (padding 1, stride 1) + bias + SiLU of SS2D, NCHW, one kernel each way
 * (ultralytics/nn/extra_modules/VManba/vmamba.py:1026-1027: x = act(conv2d(x)); conv2d = nn.Conv2d(d, d, 3, padding=1,
 * groups=d), act = SiLU).
 *   x, y, grad_y, grad_x [Bn, D, H, W] f32 | bf16 (dtype); weight f32 [D, 3, 3]; bias f32 [D] or NULL
 *   backward: pre-activations are recomputed from x; grad_weight f32 [D, 3, 3] and grad_bias f32 [D] (or NULL) are
 *   zeroed by the call and accumulated with fp32 atomics. */
int tamtr_dwconv3x3_silu_forward(const void *x, const float *weight, const float *bias, void *y, int dtype, int Bn, int D,
                                 int H, int W, void *stream);
int tamtr_dwconv3x3_silu_backward(const void *grad_y, const void *x, const float *weight, const float *bias, void *grad_x,
                                  float *grad_weight, float *grad_bias, int dtype, int Bn, int D, int H, int W, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Training-side glue on the device with FIXED shapes (one captured CUDA graph serves every batch): the batch's ground
 * truth is padded -- gt_box [B, G, 4] f32 (cx, cy, w, h), gt_cls [B, G] int64, count [B] int32 on the device -- and
 * everything that depends on the counts is decided inside the kernels.
 *
 * tamtr_cdn_group: contrastive denoising group, ultralytics/models/utils/ops.py:152-291.
 *   uniforms [B, Dmax, 10] f32 in [0,1): per slot (label flip test, new label, 4 box-noise signs, 4 magnitudes) -- the
 *     caller draws them (torch.rand: graph-safe Philox), the kernel applies the reference's formulas to them
 *   dn_cls [B, Dmax] int64 (noised labels; 0 in empty slots), dn_box [B, Dmax, 4] f32 in logit space (0 in empty slots),
 *   dn_valid [B, Dmax] f32 (1 = the slot holds a query: multiply the class embedding by it), attn_mask [Dmax+nq, Dmax+nq]
 *   uint8 (1 = blocked).  Slot layout as the reference (2 * num_group copies in slots of max(count); copies
 *   [0, num_group) positive, the rest negative; attention groups = pairs of copies); Dmax >= 2 * max(count) * num_group
 *   is the bucket's capacity: the slots beyond are padding (blocked both ways, ignored by the loss).
 * tamtr_match_cost: cost [NL, B, Q, G] f32 of query q against ground truth g of ITS image (ops.py:77-112, focal class
 *   cost + L1 + (1 - RIoU); pred_box [NL, B, Q, 4] post-sigmoid, pred_score [NL, B, Q, nc] logits).
 * tamtr_linear_sum_assignment_padded: scipy-exact assignment per (layer, image) on the first count[b] columns;
 *   match [NL, B, Q] int32 = ground-truth index or -1.
 * tamtr_detection_loss: losses of NL layers (loss.py:85-167, 232-326, 376-443; VFL / focal + L1 + RIoU) and the
 *   derivatives of the three un-normalised sums w.r.t. the predictions, in one pass.
 *   dn_group = 0: matched group (match required); 1: denoising group (targets follow from the slot layout, num_dn)
 *   partial [NL, B, 3] scratch; out [NL*3 + 1]: (class, bbox, giou) per layer with gains and the 1/pairs normaliser
 *   applied, then 1/pairs itself; d_l1, d_giou [NL, B, Q, 4], d_cls [NL, B, Q, nc] (multiply by gain/pairs and the
 *   upstream gradient of the corresponding loss). */
int tamtr_cdn_group(const float *gt_box, const long long *gt_cls, const int *count, const float *uniforms,
                    long long *dn_cls, float *dn_box, float *dn_valid, unsigned char *attn_mask, int B, int G, int Dmax,
                    int nq, int nc, int num_dn, float cls_noise_ratio, float box_noise_scale, void *stream);
int tamtr_match_cost(const float *pred_box, const float *pred_score, const float *gt_box, const long long *gt_cls,
                     float *cost, int NL, int B, int Q, int G, int nc, float alpha, float gamma, float gain_class,
                     float gain_bbox, float gain_giou, void *stream);
int tamtr_linear_sum_assignment_padded(const float *C, const int *count_dev, int *match, int n_layers, int bs, int nq,
                                       int max_gt, void *stream);
int tamtr_detection_loss(const float *pred_box, const float *pred_score, const float *gt_box, const long long *gt_cls,
                         const int *count, const int *match, float *partial, float *out, float *d_l1, float *d_giou,
                         float *d_cls, int NL, int B, int Q, int G, int nc, int dn_group, int num_dn, int use_vfl,
                         float gain_class, float gain_bbox, float gain_giou, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Optimizer step on flat buffers: clip_grad_norm_(max_norm) + AdamW, the reference's optimizer_step
 * (ultralytics/engine/trainer.py:471-477; parameter groups of build_optimizer, :654-677: weights with decay, biases
 * and normalisation weights without).
 *   param, grad, exp_avg, exp_avg_sq [n] f32 (n a multiple of 4); decay4 [n/4] uint8 or NULL: non-zero = the four
 *   elements take weight decay (parameters start at multiples of 4); partial [tamtr_optim_partials(n)] f32 scratch;
 *   step [1] f32 on the device: number of steps taken so far (incremented by the call)
 *   max_norm <= 0: no clipping.  Same arithmetic as torch.optim.AdamW (decoupled decay, bias correction) with the global
 *   gradient norm computed in a fixed order (deterministic).  No host synchronisation; CUDA-graph capturable. */
int tamtr_optim_partials(long long n);
int tamtr_adamw_flat(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, long long n,
                     const unsigned char *decay4, float *partial, float *step, float lr, float beta1, float beta2, float eps, float weight_decay,
                     float max_norm, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Folded encoder-side projections (tcgen05 / TMEM / TMA; csrc/tokgemm.cu; algebra in tamtr_b200/fold.py).
 * Replaces, per pyramid level l, the chain  input_proj[l] = Conv2d(1x1, bias=False) + BatchNorm2d
 * (ultralytics/nn/modules/head.py:1202-1218)  ->  value_proj of every decoder layer (transformer.py:273)  and
 * enc_output.0 / enc_score_head for the query-selection ranking (head.py:1229-1237): BatchNorm is affine per channel, so
 * each consumer's first Linear folds with it and the conv into ONE weight W_fold [N, C_l] + bias [N] per level.
 *
 *   tamtr_tok_project:  out[b, tok, n] = sum_c x[b, c, tok] * w[n, c] + bias[n]
 *     x     [B, C, HW] bf16 (an NCHW feature map as the backbone wrote it), C % 64 == 0, C <= 512, HW % 8 == 0
 *     w     [N0 + N1 + NT, C] bf16,  bias [N0 + N1 + NT] f32
 *     out0  bf16, columns [0, N0): element (b, tok, n) at out0 + b * out0_img + tok * out0_row + n  (strides in elements;
 *           the caller passes the level's first token of a [B, Lv, N0] tensor), N0 % 64 == 0
 *     zero0 NULL, or a second bf16 tensor laid out like out0: the same (token, column) range is filled with zeros (the
 *           gradient arena the samplers' backward accumulates into: saves the 1.65 GB memset node of the training step)
 *     out1  bf16, columns [N0, N0 + N1) likewise (NULL when N1 == 0), N1 % 64 == 0
 *     raw   f32, the last NT columns (NT % 4 == 0; NULL when NT == 0)
 *   tamtr_tok_reduce:   D[m, c] = sum_{b, tok} a[m; b, tok] * x[b, c, tok],   rs[m] = sum_{b, tok} a[m; b, tok]
 *     a_token_major = 1: a is [B, HW, M] bf16 with strides a_img / a_row (grad_value of one level: the weight gradient of
 *                        the folded projection and, through rs, of its bias);  = 0: a is [B, M, HW] (a = x: second moments
 *                        and channel sums of the level, from which the BatchNorm batch statistics follow)
 *     part_d [S, M, C] f32, part_rs [S, M] f32 with S = tamtr_tok_reduce_splits(...): per-split partial results in a fixed
 *     order (deterministic); the caller adds them.
 * Both only enqueue one kernel; CUDA-graph capturable. */
int tamtr_tok_project_supported(int B, int C, int HW, int N0, int N1, int NT);
int tamtr_tok_project(const void *x_bf16, const void *w_bf16, const float *bias, void *out0, void *zero0, long out0_row,
                      long out0_img, void *out1, long out1_row, long out1_img, float *raw, long raw_row, long raw_img, int B,
                      int C, int HW, int N0, int N1, int NT, void *stream);
/* tamtr_tok_project with the query-selection ranking (head.py:1229-1237) finished in its epilogue: the N1 = d columns are
 * E = enc_output.0(feats) without its bias, the NT tail columns are E @ (enc_score_head.weight * ln.weight)^T for the nc
 * classes (columns [0, nc)) and E . enc_bias (column NT - 1); neither is stored.  Per token:
 *   (mean, rstd) = LayerNorm statistics of E + enc_bias over d, from sum E, sum E^2 and E . enc_bias
 *   rank[b, tok] = max_k rstd * (tail[k] + bw[k] - mean * sw[k]) + ck[k]     (E and the tail count as 0 where valid[tok] == 0)
 *   rank_consts = { sum enc_bias, sum enc_bias^2, bw[NT], sw[NT], ck[NT] } f32 (the constants of tamtr_rank_tokens)
 *   rank f32: element (b, tok) at rank + b * rank_img + tok; valid u8 [HW] (the level's slice of the anchor validity mask)
 * nc < NT <= 64. */
int tamtr_tok_project_rank(const void *x_bf16, const void *w_bf16, const float *bias, void *out0, void *zero0,
                           long out0_row, long out0_img, float *rank, long rank_img, const uint8_t *valid, const float *rank_consts, int nc,
                           float eps, int B, int C, int HW, int N0, int N1, int NT, void *stream);
int tamtr_tok_reduce_supported(int B, int C, int HW, int M, int a_token_major);
int tamtr_tok_reduce_splits(int B, int C, int HW, int M, int a_token_major);
int tamtr_tok_reduce(const void *a_bf16, long a_row, long a_img, int a_token_major, const void *x_bf16, float *part_d,
                     float *part_rs, int B, int C, int HW, int M, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Glue of the folded encoder side (csrc/foldglue.cu; the algebra and its derivation are in tamtr_b200/fold.py): the
 * element-wise work between the token reductions / projections above and the small dense products that stay library
 * GEMMs, one launch for all pyramid levels each.  Level tables (C, S, n_tok, pointer arrays) are host arrays of L <= 8
 * entries; every tensor is f32 unless noted.  K = max_l C_l + 1 rounded up to a multiple of 8 (column max_l C_l carries the
 * bias part, the columns after it are zero).
 *   tamtr_fold_stats   partials of tamtr_tok_reduce(x, x) (part_d [S_l, C_l, C_l], part_rs [S_l, C_l]) and the token
 *                      count n_tok -> mean_x [C_l], cov [C_l, C_l] = E[x x^T] - mean mean^T
 *   tamtr_fold_bn      BatchNorm2d (torch/nn/modules/batchnorm.py:155-193) of y = wc x from P = wc cov, wc [d, C_l] and
 *                      mean_x: mu = wc mean, var = diag(P wc^T) (batch_stats) or the running statistics; updates
 *                      run_mean / run_var / n_batches (update_running) with the unbiased variance; s = gamma * rstd,
 *                      t = beta - mu * s; a_ext [L, d, K] = [s * wc | 0 | t], a_ext_t [L, K, d] its transpose,
 *                      stats [L, d, 4] = (mu, rstd, s, t)
 *   tamtr_fold_pack    Fv [L, N0, K], Fe [L, NE, K] (folded weights, column K-1 = folded bias part), bv [N0] ->
 *                      w_out[l] bf16 [N0 + NE, C_l] and bias [L, N0 + NE]: the operands of tamtr_tok_project
 *   tamtr_fold_unpack  partials of tamtr_tok_reduce(grad_value, x) per level -> dF [L, N0, K], dF_t [N0, L, K], d_bv [N0]
 *   tamtr_fold_bn_bwd  dA [L, d, K] (+ dAt [L, K, d] or NULL, added) -> d_wc[l] [d, C_l], d_gamma[l], d_beta[l]; d_stat [L, d, 2]
 *                      (or NULL) = (d mu, d var) per channel, for callers that carry the gradient on to the feature maps
 *   tamtr_fold_gather  xcat [R, L * K]: for the (image, token) pair flat_idx[r] = image * Lv + token the bf16 column
 *                      x_l[b, :, token - start_l] of its level l in block l (1 in the block's last column), 0 elsewhere:
 *                      the rows head.py:1240 gathers from `feats` are xcat @ a_ext_t.view(L * K, d) */
/*   tamtr_fold_rank_consts  the ranking branch's operand and constants from the parameters (enc_output.0 weight We [d, d]
 *                      and bias eb, enc_score_head weight [nc, d] and bias, all of dtype lin_dtype; LayerNorm weight /
 *                      bias f32): we_all [d + NT, d] = We (f32), then the NT COEFFICIENT rows of the tail (class k:
 *                      score_w[k] * ln_w; the last row, fused mode: eb) -- the caller multiplies them by We in place, one
 *                      small GEMM; consts [2 + 3 * NT] = { sum eb, sum eb^2, bw, sw, ck } of tamtr_tok_project_rank
 *                      (written in fused mode only) */
int tamtr_fold_rank_consts(const void *We, const void *eb, const void *score_w, const void *score_b, const float *ln_w,
                           const float *ln_b, float *we_all, float *consts, int d, int nc, int NT, int fused, int lin_dtype,
                           void *stream);
int tamtr_fold_stats(int L, const int *C, const int *S, const float *n_tok, const float *const *part_d,
                     const float *const *part_rs, float *const *mean_x, float *const *cov, void *stream);
int tamtr_fold_bn(int L, int d, const int *C, const float *n_tok, const float *const *wc, const float *const *P,
                  const float *const *mean_x, const float *const *gamma, const float *const *beta, float *const *run_mean,
                  float *const *run_var, long long *const *n_batches, const float *momentum, const float *eps,
                  int batch_stats, int update_running, float *a_ext, float *a_ext_t, float *stats, void *stream);
int tamtr_fold_pack(int L, const int *C, const float *Fv, const float *Fe, const float *bv, void *const *w_out, float *bias,
                    int N0, int NE, void *stream);
int tamtr_fold_unpack(int L, const int *C, const int *S, const float *const *part_d, const float *const *part_rs, float *dF,
                      float *dF_t, float *d_bv, int N0, void *stream);
int tamtr_fold_bn_bwd(int L, int d, const int *C, const float *const *wc, const float *const *P, const float *const *mean_x,
                      const float *const *gamma, const float *dA, const float *dAt, const float *stats, int batch_stats,
                      float *const *d_wc, float *const *d_gamma, float *const *d_beta, float *d_stat, void *stream);
int tamtr_fold_gather(int L, int Lv, const int *C, const int *start, const int *hw, const void *const *x_bf16,
                      const long long *flat_idx, float *xcat, int R, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TAMTR_B200_H */
