// SS2D tail of VMamba's VSSBlock inside TAM-TR's MEH head (nn/extra_modules/VManba/vmamba.py:1003-1017 and 1029-1034):
//     y = out_norm(y_merged^T)            LayerNorm over the d_inner channels of every position   (vmamba.py:1011-1014)
//     y = y * act(z)                      SiLU gate                                                (vmamba.py:1029-1031)
// The scan works position-major ([b, d, L]: L contiguous) and the merged result of the four directions arrives in that
// layout; the reference transposes it to [b, L, d] (one pass), normalises (one pass), casts (one pass), multiplies by the
// activated gate (two passes).  Here: ONE kernel each way that normalises over the STRIDED channel dimension -- a thread
// owns a position, consecutive lanes consecutive positions, so every channel row is read and written as full 128-byte
// lines -- applies the gate, and leaves the result in [b, d, L], which the out-projection GEMM consumes as a transposed
// operand.  HBM-bound streaming: forward reads y (fp32) twice (the second time from L2) and z once, writes once.
#include "common.cuh"

namespace tamtr {

template <typename T> __device__ __forceinline__ float ldf(const T *p, size_t i);
template <> __device__ __forceinline__ float ldf<float>(const float *p, size_t i) { return __ldg(p + i); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16 *p, size_t i) {
    return __bfloat162float(__ldg(p + i));
}
template <typename T> __device__ __forceinline__ void stf(T *p, size_t i, float v);
template <> __device__ __forceinline__ void stf<float>(float *p, size_t i, float v) { p[i] = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16 *p, size_t i, float v) { p[i] = __float2bfloat16_rn(v); }

constexpr int kCnThreads = 128;

// statistics over the channels of one position, shifted by the first channel's value (one pass, no cancellation problem for
// d <= a few thousand in fp32)
template <typename TZ, typename TO>
__global__ void __launch_bounds__(kCnThreads)
colnorm_gate_fwd_kernel(const float *__restrict__ y, const TZ *__restrict__ z, const float *__restrict__ gamma,
                        const float *__restrict__ beta, TO *__restrict__ out, float *__restrict__ mean_o,
                        float *__restrict__ rstd_o, int d, int L, float eps) {
    const int l = blockIdx.x * kCnThreads + threadIdx.x, b = blockIdx.y;
    if (l >= L) return;
    const size_t base = (size_t)b * d * L + l;
    const float shift = __ldg(y + base);
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll 8
    for (int c = 0; c < d; ++c) {
        const float v = __ldg(y + base + (size_t)c * L) - shift;
        s1 += v;
        s2 = fmaf(v, v, s2);
    }
    const float inv_d = 1.0f / (float)d;
    const float m = s1 * inv_d;
    const float mean = shift + m;
    const float rstd = rsqrtf(fmaxf(fmaf(-m, m, s2 * inv_d), 0.0f) + eps);
    if (mean_o != nullptr) {
        mean_o[(size_t)b * L + l] = mean;
        rstd_o[(size_t)b * L + l] = rstd;
    }
#pragma unroll 8
    for (int c = 0; c < d; ++c) {
        const size_t i = base + (size_t)c * L;
        const float n = (__ldg(y + i) - mean) * rstd;
        const float ln = fmaf(n, __ldg(gamma + c), __ldg(beta + c));
        const float zz = ldf(z, i);
        const float s = __fdividef(zz, 1.0f + __expf(-zz));
        stf(out, i, ln * s);
    }
}

// d_y (LayerNorm backward over the channels) and d_z (SiLU gate backward), one thread per position
template <typename TZ, typename TG>
__global__ void __launch_bounds__(kCnThreads)
colnorm_gate_bwd_kernel(const TG *__restrict__ dout, const float *__restrict__ y, const TZ *__restrict__ z,
                        const float *__restrict__ gamma, const float *__restrict__ beta, const float *__restrict__ mean_i,
                        const float *__restrict__ rstd_i, float *__restrict__ d_y, TZ *__restrict__ d_z, int d, int L) {
    const int l = blockIdx.x * kCnThreads + threadIdx.x, b = blockIdx.y;
    if (l >= L) return;
    const size_t base = (size_t)b * d * L + l;
    const float mean = mean_i[(size_t)b * L + l], rstd = rstd_i[(size_t)b * L + l];
    float a1 = 0.0f, a2 = 0.0f;
#pragma unroll 4
    for (int c = 0; c < d; ++c) {
        const size_t i = base + (size_t)c * L;
        const float n = (__ldg(y + i) - mean) * rstd;
        const float zz = ldf(z, i);
        const float s = __fdividef(zz, 1.0f + __expf(-zz));
        const float dn = ldf(dout, i) * s * __ldg(gamma + c);
        a1 += dn;
        a2 = fmaf(dn, n, a2);
    }
    const float inv_d = 1.0f / (float)d;
    a1 *= inv_d;
    a2 *= inv_d;
#pragma unroll 4
    for (int c = 0; c < d; ++c) {
        const size_t i = base + (size_t)c * L;
        const float n = (__ldg(y + i) - mean) * rstd;
        const float g = __ldg(gamma + c);
        const float ln = fmaf(n, g, __ldg(beta + c));
        const float zz = ldf(z, i);
        const float sig = __fdividef(1.0f, 1.0f + __expf(-zz));
        const float go = ldf(dout, i);
        const float dn = go * (zz * sig) * g;
        d_y[i] = rstd * (dn - a1 - n * a2);
        stf(d_z, i, go * ln * sig * fmaf(zz, 1.0f - sig, 1.0f));
    }
}

// d_gamma / d_beta: sums over (image, position) of d_ln * n and d_ln; one CTA per (channel, image) row
template <typename TZ, typename TG>
__global__ void __launch_bounds__(256)
colnorm_gate_wgrad_kernel(const TG *__restrict__ dout, const float *__restrict__ y, const TZ *__restrict__ z,
                          const float *__restrict__ mean_i, const float *__restrict__ rstd_i, float *__restrict__ d_gamma,
                          float *__restrict__ d_beta, int d, int L) {
    const int c = blockIdx.x, b = blockIdx.y;
    const size_t row = ((size_t)b * d + c) * L, st = (size_t)b * L;
    float g1 = 0.0f, g2 = 0.0f;
    for (int l = threadIdx.x; l < L; l += 256) {
        const float n = (__ldg(y + row + l) - __ldg(mean_i + st + l)) * __ldg(rstd_i + st + l);
        const float zz = ldf(z, row + l);
        const float dl = ldf(dout, row + l) * __fdividef(zz, 1.0f + __expf(-zz));
        g1 = fmaf(dl, n, g1);
        g2 += dl;
    }
    __shared__ float red[2][8];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        g1 += __shfl_xor_sync(0xffffffffu, g1, o);
        g2 += __shfl_xor_sync(0xffffffffu, g2, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = g1; red[1][threadIdx.x >> 5] = g2; }
    __syncthreads();
    if (threadIdx.x < 8) {
        g1 = red[0][threadIdx.x];
        g2 = red[1][threadIdx.x];
#pragma unroll
        for (int o = 4; o >= 1; o >>= 1) {
            g1 += __shfl_xor_sync(0xffu, g1, o);
            g2 += __shfl_xor_sync(0xffu, g2, o);
        }
        if (threadIdx.x == 0) {
            atomicAdd(d_gamma + c, g1);
            atomicAdd(d_beta + c, g2);
        }
    }
}

}  // namespace tamtr

using namespace tamtr;

static int colnorm_check(int B, int d, int L, int dt_a, int dt_b) {
    TAMTR_CHECK_ARG(B > 0 && d > 0 && L > 0, TAMTR_E_BADARG, "colnorm_gate: non-positive size");
    TAMTR_CHECK_ARG(B <= 65535 && d <= 65535, TAMTR_E_UNSUPPORTED, "colnorm_gate: batch / channels too large");
    TAMTR_CHECK_ARG((dt_a == TAMTR_F32 || dt_a == TAMTR_BF16) && (dt_b == TAMTR_F32 || dt_b == TAMTR_BF16), TAMTR_E_UNSUPPORTED,
                    "colnorm_gate: dtypes %d / %d", dt_a, dt_b);
    return 0;
}

extern "C" int tamtr_colnorm_gate_forward(const float *y, const void *z, int z_dtype, const float *gamma, const float *beta,
                                          void *out, int out_dtype, float *mean, float *rstd, int B, int d, int L, float eps,
                                          void *stream) {
    TAMTR_CHECK_ARG(y && z && gamma && beta && out && ((mean == nullptr) == (rstd == nullptr)), TAMTR_E_BADARG,
                    "colnorm_gate_forward: null pointer");
    int rc = colnorm_check(B, d, L, z_dtype, out_dtype);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid((L + kCnThreads - 1) / kCnThreads, B);
#define TAMTR_CN_FWD(TZ, TO)                                                                                          \
    colnorm_gate_fwd_kernel<TZ, TO><<<grid, kCnThreads, 0, st>>>(y, (const TZ *)z, gamma, beta, (TO *)out, mean, rstd, d, L, eps)
    if (z_dtype == TAMTR_BF16 && out_dtype == TAMTR_BF16) TAMTR_CN_FWD(__nv_bfloat16, __nv_bfloat16);
    else if (z_dtype == TAMTR_BF16) TAMTR_CN_FWD(__nv_bfloat16, float);
    else if (out_dtype == TAMTR_BF16) TAMTR_CN_FWD(float, __nv_bfloat16);
    else TAMTR_CN_FWD(float, float);
#undef TAMTR_CN_FWD
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_colnorm_gate_backward(const void *dout, int dout_dtype, const float *y, const void *z, int z_dtype,
                                           const float *gamma, const float *beta, const float *mean, const float *rstd,
                                           float *d_y, void *d_z, float *d_gamma, float *d_beta, int B, int d, int L,
                                           void *stream) {
    TAMTR_CHECK_ARG(dout && y && z && gamma && beta && mean && rstd && d_y && d_z && d_gamma && d_beta, TAMTR_E_BADARG,
                    "colnorm_gate_backward: null pointer");
    int rc = colnorm_check(B, d, L, z_dtype, dout_dtype);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    TAMTR_CUDA_OK(cudaMemsetAsync(d_gamma, 0, (size_t)d * sizeof(float), st));
    TAMTR_CUDA_OK(cudaMemsetAsync(d_beta, 0, (size_t)d * sizeof(float), st));
    const dim3 grid((L + kCnThreads - 1) / kCnThreads, B), wgrid(d, B);
#define TAMTR_CN_BWD(TZ, TG)                                                                                          \
    do {                                                                                                              \
        colnorm_gate_bwd_kernel<TZ, TG><<<grid, kCnThreads, 0, st>>>((const TG *)dout, y, (const TZ *)z, gamma, beta, mean, \
                                                                     rstd, d_y, (TZ *)d_z, d, L);                     \
        colnorm_gate_wgrad_kernel<TZ, TG><<<wgrid, 256, 0, st>>>((const TG *)dout, y, (const TZ *)z, mean, rstd, d_gamma, \
                                                                 d_beta, d, L);                                       \
    } while (0)
    if (z_dtype == TAMTR_BF16 && dout_dtype == TAMTR_BF16) TAMTR_CN_BWD(__nv_bfloat16, __nv_bfloat16);
    else if (z_dtype == TAMTR_BF16) TAMTR_CN_BWD(__nv_bfloat16, float);
    else if (dout_dtype == TAMTR_BF16) TAMTR_CN_BWD(float, __nv_bfloat16);
    else TAMTR_CN_BWD(float, float);
#undef TAMTR_CN_BWD
    count_launch(2);
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}
