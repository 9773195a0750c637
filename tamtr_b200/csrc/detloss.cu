// Training-side glue of the detection heads on the device, with FIXED shapes so that one captured CUDA graph serves
// every batch: the ground truth of a batch lives in padded device tensors (boxes [B, G, 4], classes [B, G], counts [B]);
// everything that depends on the counts -- the denoising-group layout, which queries are matched, the normalisers of
// the losses -- is decided inside the kernels from the count array, never on the host.
//
//   tamtr_cdn_group        contrastive denoising queries + attention mask      ultralytics/models/utils/ops.py:152-291
//   tamtr_match_cost       query x ground-truth cost matrices                  ultralytics/models/utils/ops.py:77-112
//   (csrc/assign.cu)       linear_sum_assignment per (layer, image)            ultralytics/models/utils/ops.py:116-121
//   tamtr_detection_loss   VFL / focal + L1 + RIoU losses of all layers, forward values and the gradients w.r.t. the
//                          predictions in the same pass        ultralytics/models/utils/loss.py:85-167, 232-326, 376-443;
//                          utils/loss.py:135-178 (VarifocalLoss / FocalLoss); utils/metrics.py:71-130 (bbox_iou, RIOU)
//
// The reference runs these as ~150 small eager ops per step around four host round trips (scipy); here they are five
// launches.  Derivatives are taken with forward-mode dual numbers over the reference's own formulas (same op order),
// so there is no hand-derived gradient to get wrong.
#include <math_constants.h>

#include "common.cuh"

namespace tamtr {

// ------------------------------------------------------------------------------------------- dual numbers
template <int N> struct Dual {
    float v;
    float d[N];
};
template <int N> __device__ __forceinline__ Dual<N> dconst(float c) {
    Dual<N> r;
    r.v = c;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = 0.0f;
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> dvar(float c, int i) {
    Dual<N> r = dconst<N>(c);
    r.d[i] = 1.0f;
    return r;
}
#define TAMTR_DUAL_BIN(NAME, VAL, DA, DB)                                                          \
    template <int N> __device__ __forceinline__ Dual<N> NAME(const Dual<N> &a, const Dual<N> &b) { \
        Dual<N> r;                                                                                  \
        r.v = (VAL);                                                                                \
        const float da = (DA), db = (DB);                                                           \
        _Pragma("unroll") for (int i = 0; i < N; ++i) r.d[i] = da * a.d[i] + db * b.d[i];           \
        return r;                                                                                   \
    }
TAMTR_DUAL_BIN(operator+, a.v + b.v, 1.0f, 1.0f)
TAMTR_DUAL_BIN(operator-, a.v - b.v, 1.0f, -1.0f)
TAMTR_DUAL_BIN(operator*, a.v * b.v, b.v, a.v)
TAMTR_DUAL_BIN(operator/, a.v / b.v, 1.0f / b.v, -a.v / (b.v * b.v))
// torch.minimum / maximum: the gradient follows the selected operand (ties have measure zero here)
TAMTR_DUAL_BIN(dmin, fminf(a.v, b.v), a.v <= b.v ? 1.0f : 0.0f, a.v <= b.v ? 0.0f : 1.0f)
TAMTR_DUAL_BIN(dmax, fmaxf(a.v, b.v), a.v >= b.v ? 1.0f : 0.0f, a.v >= b.v ? 0.0f : 1.0f)
#undef TAMTR_DUAL_BIN
template <int N> __device__ __forceinline__ Dual<N> dscale(const Dual<N> &a, float s) {
    Dual<N> r;
    r.v = a.v * s;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * s;
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> dshift(const Dual<N> &a, float s) {
    Dual<N> r = a;
    r.v += s;
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> dchain(const Dual<N> &a, float value, float slope) {
    Dual<N> r;
    r.v = value;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * slope;
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> dclamp0(const Dual<N> &a) { return dchain(a, fmaxf(a.v, 0.0f), a.v > 0.0f ? 1.0f : 0.0f); }
template <int N> __device__ __forceinline__ Dual<N> dsqrt(const Dual<N> &a) {
    const float s = sqrtf(a.v);
    return dchain(a, s, s > 0.0f ? 0.5f / s : 0.0f);
}
template <int N> __device__ __forceinline__ Dual<N> datan(const Dual<N> &a) { return dchain(a, atanf(a.v), 1.0f / (1.0f + a.v * a.v)); }
template <int N> __device__ __forceinline__ Dual<N> dsquare(const Dual<N> &a) { return dchain(a, a.v * a.v, 2.0f * a.v); }

// ------------------------------------------------------------------------------------------- box arithmetic
constexpr float kIouEps = 1e-7f;

struct Box { float x, y, w, h; };

// bbox_iou(box1, box2, xywh=True) (utils/metrics.py:94-111): plain IoU, no gradient needed (it only feeds gt_score and,
// detached, alpha)
__device__ __forceinline__ float iou_xywh(const Box &a, const Box &b) {
    const float aw = a.w / 2, ah = a.h / 2, bw = b.w / 2, bh = b.h / 2;
    const float ax1 = a.x - aw, ax2 = a.x + aw, ay1 = a.y - ah, ay2 = a.y + ah;
    const float bx1 = b.x - bw, bx2 = b.x + bw, by1 = b.y - bh, by2 = b.y + bh;
    const float inter = fmaxf(fminf(ax2, bx2) - fmaxf(ax1, bx1), 0.0f) * fmaxf(fminf(ay2, by2) - fmaxf(ay1, by1), 0.0f);
    const float uni = a.w * a.h + b.w * b.h - inter + kIouEps;
    return inter / uni;
}

// bbox_iou(box1, box2, xywh=True, RIOU=True) (utils/metrics.py:94-130) with box1 = the prediction as the variable:
// returns the value and d/d(cx, cy, w, h) of box1.  alpha is formed from detached values, as under torch.no_grad().
template <int N>
__device__ __forceinline__ Dual<N> riou_xywh(const Dual<N> &x1, const Dual<N> &y1, const Dual<N> &w1, const Dual<N> &h1,
                                             const Box &g) {
    const Dual<N> w1h = dscale(w1, 0.5f), h1h = dscale(h1, 0.5f);
    const Dual<N> b1x1 = x1 - w1h, b1x2 = x1 + w1h, b1y1 = y1 - h1h, b1y2 = y1 + h1h;
    const float w2h = g.w / 2, h2h = g.h / 2;
    const Dual<N> b2x1 = dconst<N>(g.x - w2h), b2x2 = dconst<N>(g.x + w2h), b2y1 = dconst<N>(g.y - h2h),
                  b2y2 = dconst<N>(g.y + h2h);
    const Dual<N> inter = dclamp0(dmin(b1x2, b2x2) - dmax(b1x1, b2x1)) * dclamp0(dmin(b1y2, b2y2) - dmax(b1y1, b2y1));
    const Dual<N> uni = dshift(w1 * h1 + dconst<N>(g.w * g.h) - inter, kIouEps);
    const Dual<N> iou = inter / uni;
    const Dual<N> rho2 = dscale(dsquare(b2x1 + b2x2 - b1x1 - b1x2) + dsquare(b2y1 + b2y2 - b1y1 - b1y2), 0.25f);
    const Dual<N> maxwh1 = dmax(w1, h1);
    const float maxwh2 = fmaxf(g.w, g.h);
    const Dual<N> c2 = dsquare(dshift(maxwh1 + dsqrt(rho2), maxwh2 + kIouEps));
    const float four_pi2 = 4.0f / (CUDART_PI_F * CUDART_PI_F);
    const Dual<N> v = dscale(dsquare(dconst<N>(atanf(g.w / g.h)) - datan(w1 / h1)), four_pi2);
    const float alpha = v.v / (v.v - iou.v + (1.0f + kIouEps));
    return iou - (rho2 / c2 + dscale(v, alpha));
}

// ------------------------------------------------------------------------------------------- denoising-group layout
// ops.py:193-196, 242-262: max_nums = the largest ground-truth count of the batch, num_group = max(1, num_dn // max_nums),
// 2 * num_group copies of every image's ground truths laid out copy after copy in slots of max_nums; copies
// [0, num_group) are the positives, [num_group, 2 * num_group) the negatives (neg_idx, ops.py:214); attention groups are
// PAIRS of consecutive copies (ops.py:278-288).
struct DnGeom {
    int max_gt, num_group, n_dn;
};
__device__ __forceinline__ DnGeom dn_geom(const int *__restrict__ count, int B, int num_dn_cfg) {
    int m = 0;
    for (int b = 0; b < B; ++b) m = max(m, __ldg(count + b));
    DnGeom g;
    g.max_gt = m;
    // num_dn_cfg > 0: the head's `num_denoising`; < 0: the number of groups itself, negated (a reference-built group)
    g.num_group = m > 0 ? (num_dn_cfg < 0 ? -num_dn_cfg : max(1, num_dn_cfg / m)) : 0;
    g.n_dn = 2 * m * g.num_group;
    return g;
}

__global__ void cdn_group_kernel(const float *__restrict__ gt_box, const long long *__restrict__ gt_cls,
                                 const int *__restrict__ count, const float *__restrict__ uni, long long *__restrict__ dn_cls,
                                 float *__restrict__ dn_box, float *__restrict__ dn_valid, unsigned char *__restrict__ mask,
                                 int B, int G, int Dmax, int nq, int nc, int num_dn_cfg, float cls_noise_ratio,
                                 float box_noise_scale) {
    const DnGeom gm = dn_geom(count, B, num_dn_cfg);
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int Lq = Dmax + nq;
    if (idx < (long)Lq * Lq) {                                  // attention mask, True = blocked (ops.py:273-288)
        const int row = (int)(idx / Lq), col = (int)(idx % Lq);
        bool blocked;
        if (row >= Dmax) blocked = col < Dmax;                  // matching queries never see denoising (or padding) slots
        else if (row >= gm.n_dn) blocked = col != row;          // padding slot of this bucket: sees only itself
        else if (col >= Dmax) blocked = false;                  // denoising queries do see the matching queries
        else if (col >= gm.n_dn) blocked = true;
        else blocked = row / (2 * gm.max_gt) != col / (2 * gm.max_gt);
        mask[idx] = blocked ? 1 : 0;
    }
    if (idx >= (long)B * Dmax) return;
    const int b = (int)(idx / Dmax), s = (int)(idx % Dmax);
    long long cls = 0;
    float out[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    float valid = 0.0f;
    if (s < gm.n_dn) {
        const int copy = s / gm.max_gt, r = s % gm.max_gt;
        if (r < __ldg(count + b)) {
            valid = 1.0f;
            const float *u = uni + idx * 10;
            cls = __ldg(gt_cls + (long)b * G + r);
            if (cls_noise_ratio > 0.0f && u[0] < cls_noise_ratio * 0.5f) cls = min((int)(u[1] * (float)nc), nc - 1);
            const float4 bx = __ldg(reinterpret_cast<const float4 *>(gt_box) + (long)b * G + r);
            float box[4] = {bx.x, bx.y, bx.z, bx.w};
            if (box_noise_scale > 0.0f) {
                // (explicit round-to-nearest intrinsics: the reference rounds after every op, ops.py:217-231 -- no FMA)
                const float dw = bx.z / 2, dh = bx.w / 2;                         // xywh2xyxy
                float xyxy[4] = {__fsub_rn(bx.x, dw), __fsub_rn(bx.y, dh), __fadd_rn(bx.x, dw), __fadd_rn(bx.y, dh)};
                const float dfw = __fmul_rn(__fmul_rn(bx.z, 0.5f), box_noise_scale);
                const float dfh = __fmul_rn(__fmul_rn(bx.w, 0.5f), box_noise_scale);
                const float diff[4] = {dfw, dfh, dfw, dfh};
                const float neg = copy >= gm.num_group ? 1.0f : 0.0f;            // negatives: 1x..2x the half size away
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float sign = u[2 + k] < 0.5f ? -1.0f : 1.0f;            // randint(0, 2) * 2 - 1
                    const float part = __fmul_rn(__fadd_rn(u[6 + k], neg), sign);
                    xyxy[k] = fminf(fmaxf(__fadd_rn(xyxy[k], __fmul_rn(part, diff[k])), 0.0f), 1.0f);
                }
                box[0] = __fadd_rn(xyxy[0], xyxy[2]) / 2;                         // xyxy2xywh
                box[1] = __fadd_rn(xyxy[1], xyxy[3]) / 2;
                box[2] = __fsub_rn(xyxy[2], xyxy[0]);
                box[3] = __fsub_rn(xyxy[3], xyxy[1]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {                                     // torch.logit(x, eps=1e-6)
                    const float z = fminf(fmaxf(box[k], 1e-6f), 1.0f - 1e-6f);
                    box[k] = logf(z / (1.0f - z));
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) out[k] = box[k];
        }
    }
    dn_cls[idx] = cls;
    dn_valid[idx] = valid;
    reinterpret_cast<float4 *>(dn_box)[idx] = make_float4(out[0], out[1], out[2], out[3]);
}

// ------------------------------------------------------------------------------------------- matching cost
// ops.py:77-112 (use_fl): focal classification cost + L1 + (1 - RIoU), per (layer, image, query, own ground truth).
__global__ void match_cost_kernel(const float *__restrict__ pred_box, const float *__restrict__ pred_score,
                                  const float *__restrict__ gt_box, const long long *__restrict__ gt_cls, float *__restrict__ C,
                                  long total, int B, int Q, int G, int nc, float alpha, float gamma, float g_class,
                                  float g_bbox, float g_giou) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int g = (int)(idx % G);
    const long lbq = idx / G;
    const int b = (int)((lbq / Q) % B);
    const float4 pb = __ldg(reinterpret_cast<const float4 *>(pred_box) + lbq);
    const float4 gb4 = __ldg(reinterpret_cast<const float4 *>(gt_box) + (long)b * G + g);
    int cls = (int)__ldg(gt_cls + (long)b * G + g);
    cls = min(max(cls, 0), nc - 1);                                  // (padding columns hold zeros; never read by the matcher)
    const float x = __ldg(pred_score + lbq * nc + cls);
    const float p = 1.0f / (1.0f + expf(-x));
    const float pg = gamma == 2.0f ? p * p : powf(p, gamma), qg = gamma == 2.0f ? (1.0f - p) * (1.0f - p) : powf(1.0f - p, gamma);
    const float neg = (1.0f - alpha) * pg * (-logf(1.0f - p + 1e-8f));
    const float pos = alpha * qg * (-logf(p + 1e-8f));
    const float c_class = pos - neg;
    const float c_bbox = fabsf(pb.x - gb4.x) + fabsf(pb.y - gb4.y) + fabsf(pb.z - gb4.z) + fabsf(pb.w - gb4.w);
    const Box gb = {gb4.x, gb4.y, gb4.z, gb4.w};
    const Dual<1> r = riou_xywh<1>(dconst<1>(pb.x), dconst<1>(pb.y), dconst<1>(pb.z), dconst<1>(pb.w), gb);
    const float c = g_class * c_class + g_bbox * c_bbox + g_giou * (1.0f - r.v);
    C[idx] = isfinite(c) ? c : 0.0f;                                 // ops.py:112
}

// ------------------------------------------------------------------------------------------- losses
// One CTA per (layer, image): every thread walks queries, the CTA writes three partial sums (class, L1, 1 - RIoU) and the
// per-prediction derivatives of those three sums (not yet divided by the number of matched pairs: the final kernel forms
// the normalisers from the counts and the autograd wrapper scales by them).
//   MATCHED group: match[l, b, q] = ground-truth index or -1 (from the assignment kernel)
//   DN group     : the target follows from the slot (positive copies of the denoising layout), nothing is read
// VFL (loss.py:108-112 -> utils/loss.py:135-147, alpha 0.75, gamma 2) when the batch has ground truths, focal loss
// (utils/loss.py:150-178, gamma 1.5, alpha 0.25) when it has none -- the reference's `if num_gts and self.vfl`.
constexpr int kLossThreads = 128;

__device__ __forceinline__ float softplus_neg_abs(float x) { return log1pf(expf(-fabsf(x))); }

template <bool DN>
__global__ void __launch_bounds__(kLossThreads)
det_loss_kernel(const float *__restrict__ pred_box, const float *__restrict__ pred_score, const float *__restrict__ gt_box,
                const long long *__restrict__ gt_cls, const int *__restrict__ count, const int *__restrict__ match,
                float *__restrict__ partial, float *__restrict__ d_l1, float *__restrict__ d_giou, float *__restrict__ d_cls,
                int B, int Q, int G, int nc, int num_dn_cfg, int use_vfl) {
    const int b = blockIdx.x, l = blockIdx.y;
    const DnGeom gm = dn_geom(count, B, num_dn_cfg);
    int total_gt = 0;
    for (int i = 0; i < B; ++i) total_gt += __ldg(count + i);
    const bool vfl = use_vfl && total_gt > 0;
    const int nb = __ldg(count + b);
    float s_cls = 0.0f, s_l1 = 0.0f, s_giou = 0.0f;
    for (int q = threadIdx.x; q < Q; q += kLossThreads) {
        const long lbq = ((long)l * B + b) * Q + q;
        int g = -1;
        bool counted = true;                                     // does this query take part in the class loss at all?
        if (DN) {
            counted = q < gm.n_dn;                               // bucket padding beyond the real denoising queries: ignored
            if (counted) {
                const int copy = q / gm.max_gt, r = q % gm.max_gt;
                if (copy < gm.num_group && r < nb) g = r;        // positives (loss.py:424-443 get_dn_match_indices)
            }
        } else {
            g = __ldg(match + lbq);
        }
        float4 dl1 = make_float4(0.f, 0.f, 0.f, 0.f), dgi = dl1;
        float gt_score = 0.0f;
        int cls = -1;
        if (g >= 0) {
            const float4 pb = __ldg(reinterpret_cast<const float4 *>(pred_box) + lbq);
            const float4 gb4 = __ldg(reinterpret_cast<const float4 *>(gt_box) + (long)b * G + g);
            const Box gb = {gb4.x, gb4.y, gb4.z, gb4.w}, pbx = {pb.x, pb.y, pb.z, pb.w};
            cls = (int)__ldg(gt_cls + (long)b * G + g);
            s_l1 += fabsf(pb.x - gb4.x) + fabsf(pb.y - gb4.y) + fabsf(pb.z - gb4.z) + fabsf(pb.w - gb4.w);
            dl1 = make_float4(pb.x > gb4.x ? 1.f : (pb.x < gb4.x ? -1.f : 0.f), pb.y > gb4.y ? 1.f : (pb.y < gb4.y ? -1.f : 0.f),
                              pb.z > gb4.z ? 1.f : (pb.z < gb4.z ? -1.f : 0.f), pb.w > gb4.w ? 1.f : (pb.w < gb4.w ? -1.f : 0.f));
            const Dual<4> r = riou_xywh<4>(dvar<4>(pb.x, 0), dvar<4>(pb.y, 1), dvar<4>(pb.z, 2), dvar<4>(pb.w, 3), gb);
            s_giou += 1.0f - r.v;
            dgi = make_float4(-r.d[0], -r.d[1], -r.d[2], -r.d[3]);
            gt_score = iou_xywh(pbx, gb);                       // bbox_iou(pred.detach(), gt) (loss.py:322)
        }
        reinterpret_cast<float4 *>(d_l1)[lbq] = dl1;
        reinterpret_cast<float4 *>(d_giou)[lbq] = dgi;
        for (int c = 0; c < nc; ++c) {
            const float x = __ldg(pred_score + lbq * nc + c);
            float loss = 0.0f, dx = 0.0f;
            if (counted) {
                const float s = 1.0f / (1.0f + expf(-x));
                const float label = c == cls ? 1.0f : 0.0f;
                if (vfl) {      // weight = 0.75 * sigmoid(x)^2 * (1 - label) + gt_score * label; loss = BCE(x, gt_score * label) * weight
                    const float t = gt_score * label;
                    const float w = 0.75f * s * s * (1.0f - label) + t;
                    const float dw = 0.75f * 2.0f * s * s * (1.0f - s) * (1.0f - label);
                    const float bce = fmaxf(x, 0.0f) - x * t + softplus_neg_abs(x);
                    loss = bce * w;
                    dx = (s - t) * w + bce * dw;
                } else {        // focal: BCE(x, label) * (1 - p_t)^1.5 * (label * 0.25 + (1 - label) * 0.75)
                    const float bce = fmaxf(x, 0.0f) - x * label + softplus_neg_abs(x);
                    const float pt = label * s + (1.0f - label) * (1.0f - s);
                    const float om = 1.0f - pt, mod = om * sqrtf(om);
                    const float af = label * 0.25f + (1.0f - label) * 0.75f;
                    const float dpt = (2.0f * label - 1.0f) * s * (1.0f - s);
                    loss = bce * mod * af;
                    dx = ((s - label) * mod - bce * 1.5f * sqrtf(om) * dpt) * af;
                }
            }
            s_cls += loss;
            d_cls[lbq * nc + c] = dx;
        }
    }
    __shared__ float red[3][kLossThreads / 32];
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        s_cls += __shfl_xor_sync(0xffffffffu, s_cls, m);
        s_l1 += __shfl_xor_sync(0xffffffffu, s_l1, m);
        s_giou += __shfl_xor_sync(0xffffffffu, s_giou, m);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = s_cls;
        red[1][threadIdx.x >> 5] = s_l1;
        red[2][threadIdx.x >> 5] = s_giou;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < kLossThreads / 32; ++w) t += red[threadIdx.x][w];
        partial[((long)l * B + b) * 3 + threadIdx.x] = t;
    }
}

// out[l, 0..2] = (class, bbox, giou) loss of layer l, gains applied, divided by the number of matched pairs
// (loss.py:112, 128-131: the class loss' mean over queries and its nq / num_gts factor cancel); out[NL, 0] = 1 / pairs
template <bool DN>
__global__ void det_loss_finish_kernel(const float *__restrict__ partial, const int *__restrict__ count, float *__restrict__ out,
                                       int NL, int B, int Q, int num_dn_cfg, float g_class, float g_bbox, float g_giou) {
    const DnGeom gm = dn_geom(count, B, num_dn_cfg);
    long pairs = 0;
    for (int b = 0; b < B; ++b) pairs += DN ? (long)gm.num_group * __ldg(count + b) : (long)min(Q, __ldg(count + b));
    const float inv = 1.0f / (float)(pairs > 0 ? pairs : 1);
    const int l = threadIdx.x;
    if (l < NL) {
        float s[3] = {0.f, 0.f, 0.f};
        for (int b = 0; b < B; ++b)
#pragma unroll
            for (int k = 0; k < 3; ++k) s[k] += partial[((long)l * B + b) * 3 + k];
        out[l * 3 + 0] = g_class * s[0] * inv;
        out[l * 3 + 1] = g_bbox * s[1] * inv;
        out[l * 3 + 2] = g_giou * s[2] * inv;
    }
    if (threadIdx.x == 0) out[NL * 3] = inv;
}

}  // namespace tamtr

using namespace tamtr;

extern "C" int tamtr_cdn_group(const float *gt_box, const long long *gt_cls, const int *count, const float *uniforms,
                               long long *dn_cls, float *dn_box, float *dn_valid, unsigned char *attn_mask, int B, int G,
                               int Dmax, int nq, int nc, int num_dn, float cls_noise_ratio, float box_noise_scale,
                               void *stream) {
    TAMTR_CHECK_ARG(gt_box && gt_cls && count && uniforms && dn_cls && dn_box && dn_valid && attn_mask, TAMTR_E_BADARG,
                    "cdn_group: null pointer");
    TAMTR_CHECK_ARG(B > 0 && B <= 4096 && G > 0 && Dmax > 0 && nq > 0 && nc > 0 && num_dn > 0, TAMTR_E_BADARG,
                    "cdn_group: bad sizes");
    const long Lq = (long)Dmax + nq;
    const long n = Lq * Lq > (long)B * Dmax ? Lq * Lq : (long)B * Dmax;
    cdn_group_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        gt_box, gt_cls, count, uniforms, dn_cls, dn_box, dn_valid, attn_mask, B, G, Dmax, nq, nc, num_dn, cls_noise_ratio,
        box_noise_scale);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_match_cost(const float *pred_box, const float *pred_score, const float *gt_box, const long long *gt_cls,
                                float *cost, int NL, int B, int Q, int G, int nc, float alpha, float gamma, float gain_class,
                                float gain_bbox, float gain_giou, void *stream) {
    TAMTR_CHECK_ARG(pred_box && pred_score && gt_box && gt_cls && cost, TAMTR_E_BADARG, "match_cost: null pointer");
    TAMTR_CHECK_ARG(NL > 0 && B > 0 && Q > 0 && G > 0 && nc > 0, TAMTR_E_BADARG, "match_cost: bad sizes");
    const long total = (long)NL * B * Q * G;
    match_cost_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        pred_box, pred_score, gt_box, gt_cls, cost, total, B, Q, G, nc, alpha, gamma, gain_class, gain_bbox, gain_giou);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_detection_loss(const float *pred_box, const float *pred_score, const float *gt_box,
                                    const long long *gt_cls, const int *count, const int *match, float *partial, float *out,
                                    float *d_l1, float *d_giou, float *d_cls, int NL, int B, int Q, int G, int nc,
                                    int dn_group, int num_dn, int use_vfl, float gain_class, float gain_bbox, float gain_giou,
                                    void *stream) {
    TAMTR_CHECK_ARG(pred_box && pred_score && gt_box && gt_cls && count && partial && out && d_l1 && d_giou && d_cls,
                    TAMTR_E_BADARG, "detection_loss: null pointer");
    TAMTR_CHECK_ARG(dn_group || match, TAMTR_E_BADARG, "detection_loss: the matched group needs the match array");
    TAMTR_CHECK_ARG(NL > 0 && NL <= 32 && B > 0 && B <= 65535 && Q > 0 && G > 0 && nc > 0, TAMTR_E_BADARG,
                    "detection_loss: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    if (dn_group) {
        det_loss_kernel<true><<<dim3(B, NL), kLossThreads, 0, st>>>(pred_box, pred_score, gt_box, gt_cls, count, nullptr,
                                                                    partial, d_l1, d_giou, d_cls, B, Q, G, nc, num_dn, use_vfl);
        det_loss_finish_kernel<true><<<1, 32, 0, st>>>(partial, count, out, NL, B, Q, num_dn, gain_class, gain_bbox, gain_giou);
    } else {
        det_loss_kernel<false><<<dim3(B, NL), kLossThreads, 0, st>>>(pred_box, pred_score, gt_box, gt_cls, count, match,
                                                                     partial, d_l1, d_giou, d_cls, B, Q, G, nc, num_dn, use_vfl);
        det_loss_finish_kernel<false><<<1, 32, 0, st>>>(partial, count, out, NL, B, Q, num_dn, gain_class, gain_bbox, gain_giou);
    }
    count_launch(2);
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}
