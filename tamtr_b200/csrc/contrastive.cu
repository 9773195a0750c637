// Text-guided classification branch: fused region-text cosine-similarity head, forward + backward (sm_100a).
//
// Replaces /root/reference ultralytics/nn/modules/block.py:534-541 (ContrastiveHeadMLP.forward: two permutes, two
// F.normalize, an einsum "bch,bkc->bkh", an affine and a permute back: ~7 launches for 49 MFLOP) with one kernel:
//   out[b,q,k] = <x[b,q,:]/max(|x|,eps), w[b,k,:]/max(|w|,eps)> * exp(logit_scale) + bias        eps = 1e-12
// One warp per query row; the K text rows of image b are read through L1 by all warps of the CTA; their inverse
// norms are computed once per CTA into shared memory.  HBM-bound on reading x (B*Lq*C elements) -- launch-bound in
// practice.
//
// x [B, Lq, C] f32|bf16, w [B, K, C] f32, logit_scale/bias: device scalars (fp32), out [B, Lq, K] f32.
#include "common.cuh"

namespace tamtr {

constexpr int kCtrWarps = 8;
constexpr int kCtrMaxK = 256;
constexpr int kCtrMaxChunks = 8;  // C <= 1024, C % 128 == 0

template <typename T> __device__ __forceinline__ float4 load4(const T *p);
template <> __device__ __forceinline__ float4 load4<float>(const float *p) {
    return __ldg(reinterpret_cast<const float4 *>(p));
}
template <> __device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16 *p) {
    const uint2 u = __ldg(reinterpret_cast<const uint2 *>(p));
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                       __uint_as_float(u.y & 0xffff0000u));
}
template <typename T> __device__ __forceinline__ void store4(T *p, float4 v);
template <> __device__ __forceinline__ void store4<float>(float *p, float4 v) {
    *reinterpret_cast<float4 *>(p) = v;
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16 *p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t *>(&a);
    u.y = *reinterpret_cast<uint32_t *>(&b);
    *reinterpret_cast<uint2 *>(p) = u;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

// inverse L2 norms of the K text rows of image b -> smem
__device__ __forceinline__ void text_inv_norms(float *s_winv, const float *wb, int K, int C, int warp, int lane) {
    for (int k = warp; k < K; k += kCtrWarps) {
        float ss = 0.f;
        for (int c = lane * 4; c < C; c += 128) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(wb + (size_t)k * C + c));
            ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        ss = warp_sum(ss);
        if (lane == 0) s_winv[k] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    }
    __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(kCtrWarps * 32)
contrastive_fwd_kernel(const T *__restrict__ x, const float *__restrict__ w, const float *__restrict__ logit_scale,
                       const float *__restrict__ bias, float *__restrict__ out, int Lq, int K, int C) {
    __shared__ float s_winv[kCtrMaxK];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const float *wb = w + (size_t)b * K * C;
    text_inv_norms(s_winv, wb, K, C, warp, lane);
    const int q = blockIdx.x * kCtrWarps + warp;
    if (q >= Lq) return;
    const int nchunk = C / 128;
    const T *xr_p = x + ((size_t)b * Lq + q) * C + lane * 4;
    float4 xr[kCtrMaxChunks];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kCtrMaxChunks; ++i) {
        xr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < nchunk) {
            xr[i] = load4<T>(xr_p + i * 128);
            ss += xr[i].x * xr[i].x + xr[i].y * xr[i].y + xr[i].z * xr[i].z + xr[i].w * xr[i].w;
        }
    }
    const float xinv = 1.0f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
    const float scale = expf(__ldg(logit_scale)), bs = __ldg(bias);
    float *orow = out + ((size_t)b * Lq + q) * K;
    for (int k = 0; k < K; ++k) {
        const float *wk = wb + (size_t)k * C + lane * 4;
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < kCtrMaxChunks; ++i)
            if (i < nchunk) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(wk + i * 128));
                d += xr[i].x * v.x + xr[i].y * v.y + xr[i].z * v.z + xr[i].w * v.w;
            }
        d = warp_sum(d);
        if (lane == 0) orow[k] = d * xinv * s_winv[k] * scale + bs;
    }
}

// grad_x [B,Lq,C] (dtype of x); grad_scalars[0] += d/d logit_scale, grad_scalars[1] += d/d bias (pre-zeroed).
template <typename T>
__global__ void __launch_bounds__(kCtrWarps * 32)
contrastive_bwd_kernel(const float *__restrict__ grad_out, const T *__restrict__ x, const float *__restrict__ w,
                       const float *__restrict__ logit_scale, T *__restrict__ grad_x,
                       float *__restrict__ grad_scalars, int Lq, int K, int C) {
    __shared__ float s_winv[kCtrMaxK];
    __shared__ float s_red[2][kCtrWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const float *wb = w + (size_t)b * K * C;
    text_inv_norms(s_winv, wb, K, C, warp, lane);
    const int q = blockIdx.x * kCtrWarps + warp;
    float part_ls = 0.f, part_b = 0.f;
    if (q < Lq) {
        const int nchunk = C / 128;
        const size_t row = ((size_t)b * Lq + q);
        const T *xr_p = x + row * C + lane * 4;
        float4 xr[kCtrMaxChunks], gx[kCtrMaxChunks];
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < kCtrMaxChunks; ++i) {
            xr[i] = gx[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < nchunk) {
                xr[i] = load4<T>(xr_p + i * 128);
                ss += xr[i].x * xr[i].x + xr[i].y * xr[i].y + xr[i].z * xr[i].z + xr[i].w * xr[i].w;
            }
        }
        const float xinv = 1.0f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
        const float scale = expf(__ldg(logit_scale));
        const float *grow = grad_out + row * K;
        for (int k = 0; k < K; ++k) {
            const float g = __ldg(grow + k);
            const float gk = g * scale * s_winv[k];   // d out / d <x_hat, w_k> times 1/|w_k|
            const float *wk = wb + (size_t)k * C + lane * 4;
            float d = 0.f;
#pragma unroll
            for (int i = 0; i < kCtrMaxChunks; ++i)
                if (i < nchunk) {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(wk + i * 128));
                    d += xr[i].x * v.x + xr[i].y * v.y + xr[i].z * v.z + xr[i].w * v.w;
                    gx[i].x = fmaf(gk, v.x, gx[i].x); gx[i].y = fmaf(gk, v.y, gx[i].y);
                    gx[i].z = fmaf(gk, v.z, gx[i].z); gx[i].w = fmaf(gk, v.w, gx[i].w);
                }
            part_ls = fmaf(gk * xinv, d, part_ls);    // lane-partial of g * c_k * scale
            part_b += (lane == 0) ? g : 0.f;
        }
        // grad_x = (grad_xhat - xhat * <xhat, grad_xhat>) / |x|     with xhat = x * xinv
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < kCtrMaxChunks; ++i)
            t += xr[i].x * gx[i].x + xr[i].y * gx[i].y + xr[i].z * gx[i].z + xr[i].w * gx[i].w;
        t = warp_sum(t) * xinv * xinv;  // <xhat, g> * xinv (one xinv folded for the x -> xhat conversion below)
        T *go = grad_x + row * C + lane * 4;
#pragma unroll
        for (int i = 0; i < kCtrMaxChunks; ++i)
            if (i < nchunk) {
                float4 r;
                r.x = (gx[i].x - xr[i].x * t) * xinv; r.y = (gx[i].y - xr[i].y * t) * xinv;
                r.z = (gx[i].z - xr[i].z * t) * xinv; r.w = (gx[i].w - xr[i].w * t) * xinv;
                store4<T>(go + i * 128, r);
            }
    }
    part_ls = warp_sum(part_ls);
    part_b = warp_sum(part_b);
    if (lane == 0) { s_red[0][warp] = part_ls; s_red[1][warp] = part_b; }
    __syncthreads();
    if (threadIdx.x < 2) {
        float s = 0.f;
        for (int i = 0; i < kCtrWarps; ++i) s += s_red[threadIdx.x][i];
        atomicAdd(grad_scalars + threadIdx.x, s);
    }
}

static int check_ctr(int dtype, int B, int Lq, int K, int C) {
    TAMTR_CHECK_ARG(dtype == TAMTR_F32 || dtype == TAMTR_BF16, TAMTR_E_UNSUPPORTED, "contrastive: dtype %d", dtype);
    TAMTR_CHECK_ARG(B > 0 && Lq > 0 && K > 0 && C > 0, TAMTR_E_BADARG, "contrastive: non-positive size");
    TAMTR_CHECK_ARG(C % 128 == 0 && C <= 128 * kCtrMaxChunks, TAMTR_E_UNSUPPORTED,
                    "contrastive: channel dim %d unsupported (need C %% 128 == 0 and C <= %d)", C, 128 * kCtrMaxChunks);
    TAMTR_CHECK_ARG(K <= kCtrMaxK, TAMTR_E_UNSUPPORTED, "contrastive: %d text tokens > %d", K, kCtrMaxK);
    TAMTR_CHECK_ARG(B <= 65535, TAMTR_E_UNSUPPORTED, "contrastive: batch %d > 65535", B);
    return 0;
}

}  // namespace tamtr

using namespace tamtr;

extern "C" int tamtr_contrastive_forward(const void *x, const float *w, const float *logit_scale, const float *bias,
                                         float *out, int dtype, int B, int Lq, int K, int C, void *stream) {
    TAMTR_CHECK_ARG(x && w && logit_scale && bias && out, TAMTR_E_BADARG, "contrastive_forward: null pointer");
    const int rc = check_ctr(dtype, B, Lq, K, C);
    if (rc) return rc;
    const dim3 grid((Lq + kCtrWarps - 1) / kCtrWarps, B);
    cudaStream_t st = (cudaStream_t)stream;
    KernelTimer timer(K_CTR_FWD, st);
    if (dtype == TAMTR_F32)
        contrastive_fwd_kernel<float><<<grid, kCtrWarps * 32, 0, st>>>((const float *)x, w, logit_scale, bias, out, Lq,
                                                                        K, C);
    else
        contrastive_fwd_kernel<__nv_bfloat16><<<grid, kCtrWarps * 32, 0, st>>>((const __nv_bfloat16 *)x, w,
                                                                                logit_scale, bias, out, Lq, K, C);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_contrastive_backward(const float *grad_out, const void *x, const float *w,
                                          const float *logit_scale, void *grad_x, float *grad_scalars, int dtype, int B,
                                          int Lq, int K, int C, void *stream) {
    TAMTR_CHECK_ARG(grad_out && x && w && logit_scale && grad_x && grad_scalars, TAMTR_E_BADARG,
                    "contrastive_backward: null pointer");
    const int rc = check_ctr(dtype, B, Lq, K, C);
    if (rc) return rc;
    const dim3 grid((Lq + kCtrWarps - 1) / kCtrWarps, B);
    cudaStream_t st = (cudaStream_t)stream;
    TAMTR_CUDA_OK(cudaMemsetAsync(grad_scalars, 0, 2 * sizeof(float), st));
    KernelTimer timer(K_CTR_BWD, st);
    if (dtype == TAMTR_F32)
        contrastive_bwd_kernel<float><<<grid, kCtrWarps * 32, 0, st>>>(grad_out, (const float *)x, w, logit_scale,
                                                                        (float *)grad_x, grad_scalars, Lq, K, C);
    else
        contrastive_bwd_kernel<__nv_bfloat16><<<grid, kCtrWarps * 32, 0, st>>>(
            grad_out, (const __nv_bfloat16 *)x, w, logit_scale, (__nv_bfloat16 *)grad_x, grad_scalars, Lq, K, C);
    count_launch(2);
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}
