// BTA-PAN text-image attention gate (max-sigmoid attention), CUDA-core path, forward + backward (sm_100a).
//
// Replaces /root/reference ultralytics/nn/extra_modules/block.py:216-220:
//   aw[b,m,y,x] = sigmoid( max_n <embed[b, m*hc:(m+1)*hc, y, x], guide[b,n,m,:]> / sqrt(hc) + bias[m] )
// (einsum "bmchw,bnmc->bmhwn", max over text tokens, scale, bias, sigmoid: five launches and a materialised
// [B, nh, H, W, N] tensor) with one pass over `embed`.  This is the exact-fp32 path (fp32 or bf16 storage, fp32 FMA);
// the tcgen05 tensor-core path for bf16 lives in maxsig_tc.cu.  Thread = pixel: channel reads are coalesced along
// the NCHW pixel axis, the N x hc guide block of (b, m) sits in shared memory and is read as broadcasts.
//
// embed [B, nh*hc, HW] f32|bf16 (NCHW), guide [B, N, nh, hc] f32, bias [nh] f32
// aw [B, nh, HW] f32, amax [B, nh, HW] uint8 (arg max over n, kept for the backward)
#include "common.cuh"

namespace tamtr {

constexpr int kGateThreads = 128;
constexpr int kGateMaxHc = 64;

template <typename T> __device__ __forceinline__ float ld1(const T *p);
template <> __device__ __forceinline__ float ld1<float>(const float *p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld1<__nv_bfloat16>(const __nv_bfloat16 *p) {
    return __uint_as_float(((uint32_t)__ldg(reinterpret_cast<const unsigned short *>(p))) << 16);
}
template <typename T> __device__ __forceinline__ void st1(T *p, float v);
template <> __device__ __forceinline__ void st1<float>(float *p, float v) { *p = v; }
template <> __device__ __forceinline__ void st1<__nv_bfloat16>(__nv_bfloat16 *p, float v) {
    *p = __float2bfloat16_rn(v);
}

template <typename T, int HC>
__global__ void __launch_bounds__(kGateThreads)
gate_fwd_kernel(const T *__restrict__ embed, const float *__restrict__ guide, const float *__restrict__ bias,
                float *__restrict__ aw, uint8_t *__restrict__ amax, int nh, int HW, int N) {
    extern __shared__ float s_g[];  // [N][HC]
    const int m = blockIdx.y, b = blockIdx.z;
    for (int i = threadIdx.x; i < N * HC; i += kGateThreads) {
        const int n = i / HC, c = i % HC;
        s_g[i] = __ldg(guide + (((size_t)b * N + n) * nh + m) * HC + c);
    }
    __syncthreads();
    const int pix = blockIdx.x * kGateThreads + threadIdx.x;
    if (pix >= HW) return;
    const T *xp = embed + ((size_t)b * nh + m) * HC * HW + pix;
    float xr[HC];
#pragma unroll
    for (int c = 0; c < HC; ++c) xr[c] = ld1<T>(xp + (size_t)c * HW);
    float best = -INFINITY;
    int arg = 0;
    for (int n = 0; n < N; ++n) {
        const float4 *gn = reinterpret_cast<const float4 *>(s_g + n * HC);
        float d = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < HC / 4; ++c4) {
            const float4 g = gn[c4];
            d = fmaf(xr[4 * c4], g.x, d); d = fmaf(xr[4 * c4 + 1], g.y, d);
            d = fmaf(xr[4 * c4 + 2], g.z, d); d = fmaf(xr[4 * c4 + 3], g.w, d);
        }
        if (d > best) { best = d; arg = n; }   // first maximum wins, like torch.max(dim)
    }
    const float z = __fdiv_rn(best, sqrtf((float)HC)) + __ldg(bias + m);
    const size_t o = ((size_t)b * nh + m) * HW + pix;
    aw[o] = 1.0f / (1.0f + expf(-z));
    amax[o] = (uint8_t)arg;
}

// grad_embed [B, nh*hc, HW] (dtype of embed; fully written), grad_guide [B,N,nh,hc] f32 and grad_bias [nh] f32
// (both pre-zeroed, accumulated with atomics after a per-CTA shared-memory reduction).
template <typename T, int HC>
__global__ void __launch_bounds__(kGateThreads)
gate_bwd_kernel(const float *__restrict__ grad_aw, const float *__restrict__ aw, const uint8_t *__restrict__ amax,
                const T *__restrict__ embed, const float *__restrict__ guide, T *__restrict__ grad_embed,
                float *__restrict__ grad_guide, float *__restrict__ grad_bias, int nh, int HW, int N) {
    extern __shared__ float s_mem[];  // g [N][HC] then gg [N][HC]
    float *s_g = s_mem, *s_gg = s_mem + N * HC;
    __shared__ float s_gb;
    const int m = blockIdx.y, b = blockIdx.z;
    for (int i = threadIdx.x; i < N * HC; i += kGateThreads) {
        const int n = i / HC, c = i % HC;
        s_g[i] = __ldg(guide + (((size_t)b * N + n) * nh + m) * HC + c);
        s_gg[i] = 0.f;
    }
    if (threadIdx.x == 0) s_gb = 0.f;
    __syncthreads();
    const int pix = blockIdx.x * kGateThreads + threadIdx.x;
    if (pix < HW) {
        const size_t o = ((size_t)b * nh + m) * HW + pix;
        const float a = aw[o];
        const float gz = grad_aw[o] * a * (1.0f - a);
        const float gl = __fdiv_rn(gz, sqrtf((float)HC));
        const int n = amax[o];
        const T *xp = embed + ((size_t)b * nh + m) * HC * HW + pix;
        T *gp = grad_embed + ((size_t)b * nh + m) * HC * HW + pix;
#pragma unroll 8
        for (int c = 0; c < HC; ++c) {
            st1<T>(gp + (size_t)c * HW, gl * s_g[n * HC + c]);
            atomicAdd(s_gg + n * HC + c, gl * ld1<T>(xp + (size_t)c * HW));
        }
        atomicAdd(&s_gb, gz);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N * HC; i += kGateThreads) {
        const float v = s_gg[i];
        if (v != 0.f) {
            const int n = i / HC, c = i % HC;
            atomicAdd(grad_guide + (((size_t)b * N + n) * nh + m) * HC + c, v);
        }
    }
    if (threadIdx.x == 0) atomicAdd(grad_bias + m, s_gb);
}

static int check_gate(int dtype, int B, int nh, int hc, int HW, int N) {
    TAMTR_CHECK_ARG(dtype == TAMTR_F32 || dtype == TAMTR_BF16, TAMTR_E_UNSUPPORTED, "max_sigmoid: dtype %d", dtype);
    TAMTR_CHECK_ARG(B > 0 && nh > 0 && HW > 0 && N > 0, TAMTR_E_BADARG, "max_sigmoid: non-positive size");
    TAMTR_CHECK_ARG(hc == 16 || hc == 32 || hc == 64, TAMTR_E_UNSUPPORTED,
                    "max_sigmoid: head channels %d unsupported (16, 32, 64)", hc);
    TAMTR_CHECK_ARG(N <= 255, TAMTR_E_UNSUPPORTED, "max_sigmoid: %d text tokens > 255", N);
    TAMTR_CHECK_ARG(B <= 65535 && nh <= 65535, TAMTR_E_UNSUPPORTED, "max_sigmoid: grid too large");
    return 0;
}

}  // namespace tamtr

using namespace tamtr;

extern "C" int tamtr_max_sigmoid_forward(const void *embed, const float *guide, const float *bias, float *aw,
                                         uint8_t *amax, int dtype, int B, int nh, int hc, int HW, int N,
                                         void *stream) {
    TAMTR_CHECK_ARG(embed && guide && bias && aw && amax, TAMTR_E_BADARG, "max_sigmoid_forward: null pointer");
    const int rc = check_gate(dtype, B, nh, hc, HW, N);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = sizeof(float) * (size_t)N * hc;
    TAMTR_CHECK_ARG(smem <= 48 * 1024, TAMTR_E_UNSUPPORTED, "max_sigmoid: N*hc = %d too large for shared memory",
                    N * hc);
#define ARGS_F(T) (const T *)embed, guide, bias, aw, amax, nh, HW, N
    {
        KernelTimer timer(K_GATE_FWD, st);
        const dim3 grid((HW + kGateThreads - 1) / kGateThreads, nh, B);
        if (dtype == TAMTR_F32) {
            if (hc == 16) gate_fwd_kernel<float, 16><<<grid, kGateThreads, smem, st>>>(ARGS_F(float));
            else if (hc == 32) gate_fwd_kernel<float, 32><<<grid, kGateThreads, smem, st>>>(ARGS_F(float));
            else gate_fwd_kernel<float, 64><<<grid, kGateThreads, smem, st>>>(ARGS_F(float));
        } else {
            if (hc == 16) gate_fwd_kernel<__nv_bfloat16, 16><<<grid, kGateThreads, smem, st>>>(ARGS_F(__nv_bfloat16));
            else if (hc == 32) gate_fwd_kernel<__nv_bfloat16, 32><<<grid, kGateThreads, smem, st>>>(ARGS_F(__nv_bfloat16));
            else gate_fwd_kernel<__nv_bfloat16, 64><<<grid, kGateThreads, smem, st>>>(ARGS_F(__nv_bfloat16));
        }
    }
#undef ARGS_F
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_max_sigmoid_backward(const float *grad_aw, const float *aw, const uint8_t *amax,
                                          const void *embed, const float *guide, void *grad_embed, float *grad_guide,
                                          float *grad_bias, int dtype, int B, int nh, int hc, int HW, int N,
                                          void *stream) {
    TAMTR_CHECK_ARG(grad_aw && aw && amax && embed && guide && grad_embed && grad_guide && grad_bias, TAMTR_E_BADARG,
                    "max_sigmoid_backward: null pointer");
    const int rc = check_gate(dtype, B, nh, hc, HW, N);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = 2 * sizeof(float) * (size_t)N * hc;
    TAMTR_CHECK_ARG(smem <= 48 * 1024, TAMTR_E_UNSUPPORTED, "max_sigmoid: N*hc = %d too large for shared memory",
                    N * hc);
    TAMTR_CUDA_OK(cudaMemsetAsync(grad_guide, 0, sizeof(float) * (size_t)B * N * nh * hc, st));
    TAMTR_CUDA_OK(cudaMemsetAsync(grad_bias, 0, sizeof(float) * (size_t)nh, st));
#define ARGS_B(T) grad_aw, aw, amax, (const T *)embed, guide, (T *)grad_embed, grad_guide, grad_bias, nh, HW, N
    {
        KernelTimer timer(K_GATE_BWD, st);
        const dim3 grid((HW + kGateThreads - 1) / kGateThreads, nh, B);
        if (dtype == TAMTR_F32) {
            if (hc == 16) gate_bwd_kernel<float, 16><<<grid, kGateThreads, smem, st>>>(ARGS_B(float));
            else if (hc == 32) gate_bwd_kernel<float, 32><<<grid, kGateThreads, smem, st>>>(ARGS_B(float));
            else gate_bwd_kernel<float, 64><<<grid, kGateThreads, smem, st>>>(ARGS_B(float));
        } else {
            if (hc == 16) gate_bwd_kernel<__nv_bfloat16, 16><<<grid, kGateThreads, smem, st>>>(ARGS_B(__nv_bfloat16));
            else if (hc == 32) gate_bwd_kernel<__nv_bfloat16, 32><<<grid, kGateThreads, smem, st>>>(ARGS_B(__nv_bfloat16));
            else gate_bwd_kernel<__nv_bfloat16, 64><<<grid, kGateThreads, smem, st>>>(ARGS_B(__nv_bfloat16));
        }
    }
#undef ARGS_B
    count_launch(3);
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}
