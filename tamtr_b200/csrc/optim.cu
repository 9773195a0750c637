// Optimizer step of the data-parallel training loop on FLAT buffers.
//
// The reference steps torch.optim.AdamW over three parameter groups after clip_grad_norm_(max_norm = 0.1)
// (ultralytics/engine/trainer.py:471-477 optimizer_step, :624-681 build_optimizer: biases and normalisation weights
// without weight decay, every other weight with it).  tamtr_b200/dp.py keeps parameters, gradients and both moments in
// one flat fp32 buffer each (the gradient buffer is the one the NCCL all-reduce runs on; which elements take weight decay
// is one byte per group of four), so the whole step is
//   tamtr_sumsq_partials   one pass over the gradients -> per-CTA partial sums of squares (deterministic: fixed order)
//   tamtr_adamw_flat       every CTA folds the partials (a few hundred floats, L2-resident) into the global norm, forms
//                          the clip coefficient min(1, max_norm / (norm + 1e-6)) exactly as clip_grad_norm_ does, and
//                          updates its slice of (param, exp_avg, exp_avg_sq); the step counter lives on the device
//   (tick)                 step += 1
// three launches, no host synchronisation, CUDA-graph capturable.  HBM-bound: 7 x 4 bytes per parameter.
#include "common.cuh"

namespace tamtr {

constexpr int kOptThreads = 256;

__global__ void __launch_bounds__(kOptThreads)
sumsq_partials_kernel(const float *__restrict__ g, long n, float *__restrict__ partial) {
    __shared__ float s_red[kOptThreads / 32];
    float acc = 0.0f;
    const long n4 = n >> 2;
    const float4 *g4 = reinterpret_cast<const float4 *>(g);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(g4 + i);
        acc = fmaf(v.x, v.x, acc);
        acc = fmaf(v.y, v.y, acc);
        acc = fmaf(v.z, v.z, acc);
        acc = fmaf(v.w, v.w, acc);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const float v = g[(n4 << 2) + threadIdx.x];
        acc = fmaf(v, v, acc);
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < kOptThreads / 32; ++w) t += s_red[w];
        partial[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kOptThreads)
adamw_flat_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v, long n,
                  const unsigned char *__restrict__ decay4, const float *__restrict__ partial, int n_partial, const float *__restrict__ step, float lr,
                  float beta1, float beta2, float eps, float weight_decay, float max_norm) {
    __shared__ float s_coef;
    if (threadIdx.x < 32) {
        float t = 0.0f;     // fixed order: lane l takes partials l, l + 32, ...  (same in every CTA and every run)
        for (int i = threadIdx.x; i < n_partial; i += 32) t += partial[i];
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) t += __shfl_xor_sync(0xffffffffu, t, k);
        if (threadIdx.x == 0) {
            float c = 1.0f;
            if (max_norm > 0.0f) c = fminf(max_norm / (sqrtf(t) + 1e-6f), 1.0f);
            s_coef = c;
        }
    }
    __syncthreads();
    const float clip = s_coef;
    const float t = step[0] + 1.0f;
    const float bc1 = 1.0f - powf(beta1, t), bc2 = 1.0f - powf(beta2, t);
    const float step_size = lr / bc1, rs_bc2 = rsqrtf(bc2);
    const float keep = 1.0f - lr * weight_decay;
    const long n4 = n >> 2;     // every parameter starts at a multiple of 4 elements (dp.FlatGrads pads)
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        float4 pv = reinterpret_cast<float4 *>(p)[i];
        const float4 gv = __ldg(reinterpret_cast<const float4 *>(g) + i);
        float4 mv = reinterpret_cast<float4 *>(m)[i], vv = reinterpret_cast<float4 *>(v)[i];
        const float dk = (decay4 && decay4[i]) ? keep : 1.0f;
        float *pp = &pv.x, *mm = &mv.x, *vq = &vv.x;
        const float *gg = &gv.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = gg[k] * clip;
            mm[k] = beta1 * mm[k] + (1.0f - beta1) * gk;
            vq[k] = beta2 * vq[k] + (1.0f - beta2) * gk * gk;
            const float denom = sqrtf(vq[k]) * rs_bc2 + eps;
            pp[k] = pp[k] * dk - step_size * (mm[k] / denom);
        }
        reinterpret_cast<float4 *>(p)[i] = pv;
        reinterpret_cast<float4 *>(m)[i] = mv;
        reinterpret_cast<float4 *>(v)[i] = vv;
    }
}

__global__ void tick_kernel(float *step) { step[0] += 1.0f; }

}  // namespace tamtr

using namespace tamtr;

extern "C" int tamtr_optim_partials(long long n) {
    const long need = (n / 4 + kOptThreads - 1) / kOptThreads;
    const long wave = (long)sm_count() * 4;
    return (int)(need < 1 ? 1 : (need < wave ? need : wave));
}

extern "C" int tamtr_adamw_flat(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, long long n,
                                const unsigned char *decay4, float *partial, float *step, float lr, float beta1, float beta2,
                                float eps, float weight_decay, float max_norm, void *stream) {
    TAMTR_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && partial && step, TAMTR_E_BADARG, "adamw_flat: null pointer");
    TAMTR_CHECK_ARG(n > 0 && n % 4 == 0, TAMTR_E_BADARG, "adamw_flat: n = %lld must be a positive multiple of 4", n);
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = tamtr_optim_partials(n);
    sumsq_partials_kernel<<<grid, kOptThreads, 0, st>>>(grad, (long)n, partial);
    adamw_flat_kernel<<<grid, kOptThreads, 0, st>>>(param, grad, exp_avg, exp_avg_sq, (long)n, decay4, partial, grid,
                                                     step, lr, beta1, beta2, eps, weight_decay, max_norm);
    tick_kernel<<<1, 1, 0, st>>>(step);
    count_launch(3);
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}
