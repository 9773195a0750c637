// Sampling-location / attention-weight epilogue of MSDeformAttn, forward + backward (sm_100a).
//
// Replaces /root/reference ultralytics/nn/modules/transformer.py:278-293: the two nn.Linear outputs are produced by
// ONE GEMM (weights concatenated to [H*L*P*3, d], see ops.py) whose fp32 result `raw` this kernel turns into
//   attn = softmax over the L*P logits of each (query, head)                        (transformer.py:279-280)
//   loc  = ref_xy + off / P * ref_wh * 0.5          (4-d reference boxes, :292-293)
//   loc  = ref_xy + off / (W_l, H_l)                (2-d reference points, :284-286)
// with the bias add fused in.  The location arithmetic reproduces the reference's op order and per-op fp32 rounding
// (torch evaluates `/ n_points`, `* wh`, `* 0.5`, `+ xy` as four separate kernels -- no FMA contraction), so the
// sampler sees the same bits it would see behind the reference's own projection.
//
// raw  [M, 3*H*S]  S = L*P; columns [0, 2*H*S) = offsets laid out (h, l, p, xy); columns [2*H*S, 3*H*S) = logits (h, s)
// bias [3*H*S]     same column order (sampling_offsets.bias ++ attention_weights.bias)
// ref  [M, RL, RD] RL in {1, L}; RD in {2, 4}
// loc  [M, H, L, P, 2], attn [M, H, L, P]   (fp32)
#include "common.cuh"

namespace tamtr {

__global__ void __launch_bounds__(128)
locw_fwd_kernel(const float *__restrict__ raw, const float *__restrict__ bias, const float *__restrict__ ref,
                float *__restrict__ loc, float *__restrict__ attn, const Levels lv, int M, int H, int RL, int RD) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (m, h)
    if (idx >= M * H) return;
    const int m = idx / H, h = idx % H;
    const int L = lv.n, P = lv.P, S = L * P;
    const float *off = raw + (size_t)m * 3 * H * S + (size_t)h * 2 * S;
    const float *lg = raw + (size_t)m * 3 * H * S + (size_t)2 * H * S + (size_t)h * S;
    const float *boff = bias + (size_t)h * 2 * S;
    const float *blg = bias + (size_t)2 * H * S + (size_t)h * S;
    float *loc_o = loc + (size_t)idx * 2 * S;
    float *attn_o = attn + (size_t)idx * S;

    // softmax over S logits (three cheap passes over L1-resident data instead of a register array)
    float zmax = -INFINITY;
    for (int s = 0; s < S; ++s) zmax = fmaxf(zmax, lg[s] + blg[s]);
    float sum = 0.0f;
    for (int s = 0; s < S; ++s) sum += expf((lg[s] + blg[s]) - zmax);
    for (int s = 0; s < S; ++s) attn_o[s] = __fdiv_rn(expf((lg[s] + blg[s]) - zmax), sum);

    const float fP = (float)P;
    for (int l = 0; l < L; ++l) {
        const float *r = ref + ((size_t)m * RL + (RL == 1 ? 0 : l)) * RD;
        const float rx = r[0], ry = r[1];
        const float sx = RD == 4 ? r[2] : (float)lv.w[l];  // 4-d: box (w, h); 2-d: normaliser (W_l, H_l)
        const float sy = RD == 4 ? r[3] : (float)lv.h[l];
        for (int p = 0; p < P; ++p) {
            const int s = l * P + p;
            const float ox = __fadd_rn(off[2 * s], boff[2 * s]);
            const float oy = __fadd_rn(off[2 * s + 1], boff[2 * s + 1]);
            float ax, ay;
            if (RD == 4) {
                ax = __fmul_rn(__fmul_rn(__fdiv_rn(ox, fP), sx), 0.5f);
                ay = __fmul_rn(__fmul_rn(__fdiv_rn(oy, fP), sy), 0.5f);
            } else {
                ax = __fdiv_rn(ox, sx);
                ay = __fdiv_rn(oy, sy);
            }
            loc_o[2 * s] = __fadd_rn(rx, ax);
            loc_o[2 * s + 1] = __fadd_rn(ry, ay);
        }
    }
}

// grad_raw [M, 3*H*S] (= gradient w.r.t. the GEMM output; also the per-row contribution to the bias gradient),
// grad_ref [M, RL, RD] or null.
__global__ void __launch_bounds__(128)
locw_bwd_kernel(const float *__restrict__ grad_loc, const float *__restrict__ grad_attn,
                const float *__restrict__ attn, const float *__restrict__ raw, const float *__restrict__ bias,
                const float *__restrict__ ref, float *__restrict__ grad_raw, float *__restrict__ grad_ref,
                const Levels lv, int M, int H, int RL, int RD) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * H) return;
    const int m = idx / H, h = idx % H;
    const int L = lv.n, P = lv.P, S = L * P;
    const float *gl = grad_loc + (size_t)idx * 2 * S;
    const float *ga = grad_attn + (size_t)idx * S;
    const float *a = attn + (size_t)idx * S;
    float *g_off = grad_raw + (size_t)m * 3 * H * S + (size_t)h * 2 * S;
    float *g_lg = grad_raw + (size_t)m * 3 * H * S + (size_t)2 * H * S + (size_t)h * S;

    // softmax backward: dz_s = a_s * (g_s - sum_j a_j g_j)
    float dot = 0.0f;
#pragma unroll 4
    for (int s = 0; s < S; ++s) dot = fmaf(a[s], ga[s], dot);
#pragma unroll 4
    for (int s = 0; s < S; ++s) g_lg[s] = a[s] * (ga[s] - dot);

    const float *off = raw + (size_t)m * 3 * H * S + (size_t)h * 2 * S;
    const float *boff = bias + (size_t)h * 2 * S;
    const float invP = 1.0f / (float)P;
    for (int l = 0; l < L; ++l) {
        const int rl = RL == 1 ? 0 : l;
        const float *r = ref + ((size_t)m * RL + rl) * RD;
        const float sx = RD == 4 ? r[2] * 0.5f * invP : 1.0f / (float)lv.w[l];
        const float sy = RD == 4 ? r[3] * 0.5f * invP : 1.0f / (float)lv.h[l];
        float gx = 0.f, gy = 0.f, gw = 0.f, gh = 0.f;
        for (int p = 0; p < P; ++p) {
            const int s = l * P + p;
            const float dx = gl[2 * s], dy = gl[2 * s + 1];
            g_off[2 * s] = dx * sx;
            g_off[2 * s + 1] = dy * sy;
            gx += dx;
            gy += dy;
            if (RD == 4 && grad_ref) {               // `raw` may be null when the reference boxes need no gradient
                gw = fmaf(dx, (off[2 * s] + boff[2 * s]) * (0.5f * invP), gw);
                gh = fmaf(dy, (off[2 * s + 1] + boff[2 * s + 1]) * (0.5f * invP), gh);
            }
        }
        if (grad_ref) {  // heads of one query accumulate into the same box
            float *gr = grad_ref + ((size_t)m * RL + rl) * RD;
            atomicAdd(gr + 0, gx);
            atomicAdd(gr + 1, gy);
            if (RD == 4) {
                atomicAdd(gr + 2, gw);
                atomicAdd(gr + 3, gh);
            }
        }
    }
}

static int check_locw(int M, int H, int L, int P, int RL, int RD, const int32_t *shapes, Levels &lv) {
    TAMTR_CHECK_ARG(M > 0 && H > 0, TAMTR_E_BADARG, "locw: non-positive size");
    TAMTR_CHECK_ARG(RD == 2 || RD == 4, TAMTR_E_BADARG,
                    "Last dim of reference_points must be 2 or 4, but got %d.", RD);  // transformer.py:295
    TAMTR_CHECK_ARG(RL == 1 || RL == L, TAMTR_E_BADARG, "locw: reference level dim %d must be 1 or n_levels=%d", RL, L);
    int32_t ones[2 * kMaxLevels];
    for (int i = 0; i < 2 * kMaxLevels; ++i) ones[i] = 1;
    const int rc = fill_levels(lv, L, P, shapes ? shapes : ones, -1);
    TAMTR_CHECK_ARG(rc == 0, rc, "locw: bad levels (L=%d, P=%d)", L, P);
    TAMTR_CHECK_ARG(RD == 4 || shapes, TAMTR_E_BADARG, "locw: 2-d reference points need level shapes");
    TAMTR_CHECK_ARG((long)M * H < (1L << 31), TAMTR_E_UNSUPPORTED, "locw: too many (query, head) pairs");
    return 0;
}

}  // namespace tamtr

using namespace tamtr;

extern "C" int tamtr_locw_forward(const float *raw, const float *bias, const float *ref, float *loc, float *attn,
                                  int M, int H, int L, int P, int RL, int RD, const int32_t *level_shapes_host,
                                  void *stream) {
    TAMTR_CHECK_ARG(raw && bias && ref && loc && attn, TAMTR_E_BADARG, "locw_forward: null pointer");
    Levels lv;
    const int rc = check_locw(M, H, L, P, RL, RD, level_shapes_host, lv);
    if (rc) return rc;
    const int n = M * H;
    KernelTimer timer(K_LOCW_FWD, (cudaStream_t)stream);
    locw_fwd_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(raw, bias, ref, loc, attn, lv, M, H, RL, RD);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_locw_backward(const float *grad_loc, const float *grad_attn, const float *attn, const float *raw,
                                   const float *bias, const float *ref, float *grad_raw, float *grad_ref, int M, int H,
                                   int L, int P, int RL, int RD, const int32_t *level_shapes_host, void *stream) {
    TAMTR_CHECK_ARG(grad_loc && grad_attn && attn && ref && grad_raw && ((raw && bias) || !grad_ref), TAMTR_E_BADARG,
                    "locw_backward: null pointer");
    Levels lv;
    const int rc = check_locw(M, H, L, P, RL, RD, level_shapes_host, lv);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (grad_ref) {
        TAMTR_CUDA_OK(cudaMemsetAsync(grad_ref, 0, sizeof(float) * (size_t)M * RL * RD, st));
        count_launch();
    }
    const int n = M * H;
    KernelTimer timer(K_LOCW_BWD, st);
    locw_bwd_kernel<<<(n + 127) / 128, 128, 0, st>>>(grad_loc, grad_attn, attn, raw, bias, ref, grad_raw, grad_ref, lv,
                                                      M, H, RL, RD);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Iterative box refinement of the decoder (transformer.py:875, 882 / 699, 706):
//   y = sigmoid(bbox + inverse_sigmoid(ref)),  inverse_sigmoid(x) = log(clamp(x,0,1).clamp(min=eps) / (1-clamp(x,0,1)).clamp(min=eps))
// The reference spends 8 elementwise launches forward and ~12 backward on B*Lq*4 numbers, five times per step.
namespace tamtr {

__global__ void box_refine_fwd_kernel(const float *__restrict__ bbox, const float *__restrict__ ref,
                                      float *__restrict__ out, int n, float eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = fminf(fmaxf(ref[i], 0.f), 1.f);
    const float x1 = fmaxf(x, eps), x2 = fmaxf(1.f - x, eps);
    const float z = bbox[i] + logf(x1 / x2);
    out[i] = 1.f / (1.f + expf(-z));
}

__global__ void box_refine_bwd_kernel(const float *__restrict__ grad_out, const float *__restrict__ out,
                                      const float *__restrict__ ref, float *__restrict__ grad_bbox,
                                      float *__restrict__ grad_ref, int n, float eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float y = out[i];
    const float gz = grad_out[i] * y * (1.f - y);
    grad_bbox[i] = gz;
    if (grad_ref) {
        const float r = ref[i];
        const float x = fminf(fmaxf(r, 0.f), 1.f);
        const float x1 = fmaxf(x, eps), x2 = fmaxf(1.f - x, eps);
        // clamp gradients as autograd defines them: pass-through where the input lies inside [min, max] (inclusive)
        const float inside = (r >= 0.f && r <= 1.f) ? 1.f : 0.f;
        const float d1 = (x >= eps) ? 1.f / x1 : 0.f;
        const float d2 = (1.f - x >= eps) ? 1.f / x2 : 0.f;
        grad_ref[i] = gz * inside * (d1 + d2);
    }
}

}  // namespace tamtr

extern "C" int tamtr_box_refine_forward(const float *bbox, const float *ref, float *out, int n, float eps, void *stream) {
    TAMTR_CHECK_ARG(bbox && ref && out && n > 0, TAMTR_E_BADARG, "box_refine_forward: bad argument");
    box_refine_fwd_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(bbox, ref, out, n, eps);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_box_refine_backward(const float *grad_out, const float *out, const float *ref, float *grad_bbox,
                                         float *grad_ref, int n, float eps, void *stream) {
    TAMTR_CHECK_ARG(grad_out && out && ref && grad_bbox && n > 0, TAMTR_E_BADARG, "box_refine_backward: bad argument");
    box_refine_bwd_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(grad_out, out, ref, grad_bbox, grad_ref, n,
                                                                             eps);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}
