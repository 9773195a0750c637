// Residual add + LayerNorm of the decoder layers (ultralytics/nn/modules/transformer.py:548,553,537:
// `embed = self.normN(embed + self.dropoutN(tgt))`, dropout p = 0 in TAM-TR, head.py:1026) as one forward and one
// backward kernel.  The tensors are tiny ([B*Lq, d] = 4800 x 512): the cost of the library path is launches and, in the
// backward, the weight/bias-gradient reduction (torch's GammaBetaBackward runs 16 CTAs for 42 us, 9 times per step).
//
//   forward : z = x + res (fp32), mean/rstd per row, y = (z - mean) * rstd * w + b; z, mean, rstd kept for the backward
//   backward: g = dy * w, xhat = (z - mean) * rstd, dz = rstd * (g - mean_d(g) - xhat * mean_d(g * xhat)),
//             dw = sum_rows dy * xhat, db = sum_rows dy.  Persistent warps walk the rows keeping their column sums in
//             registers, fold them per CTA in shared memory and add them to dw/db with one fp32 reduction per column.
// One warp per row, lanes own float4 packs of columns (d % 128 == 0, d <= 512).
#include "common.cuh"

namespace tamtr {

constexpr int kLnThreads = 256;
constexpr int kLnWarps = kLnThreads / 32;

__device__ __forceinline__ float4 ln_load4(const void *p, int dtype, size_t idx) {
    if (dtype == TAMTR_F32) return *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(p) + idx);
    const uint2 u = *reinterpret_cast<const uint2 *>(reinterpret_cast<const __nv_bfloat16 *>(p) + idx);
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                       __uint_as_float(u.y & 0xffff0000u));
}
__device__ __forceinline__ void ln_store4(void *p, int dtype, size_t idx, float4 v) {
    if (dtype == TAMTR_F32) {
        *reinterpret_cast<float4 *>(reinterpret_cast<float *>(p) + idx) = v;
    } else {
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t *>(&lo);
        u.y = *reinterpret_cast<uint32_t *>(&hi);
        *reinterpret_cast<uint2 *>(reinterpret_cast<__nv_bfloat16 *>(p) + idx) = u;
    }
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int PP>   // float4 packs per lane: d = 128 * PP
__global__ void __launch_bounds__(kLnThreads)
add_layernorm_fwd_kernel(const void *__restrict__ x, int x_dtype, const void *__restrict__ res, int res_dtype,
                         const float *__restrict__ w, const float *__restrict__ b, void *__restrict__ y, int y_dtype,
                         float *__restrict__ z, float *__restrict__ mean, float *__restrict__ rstd, int rows, float eps,
                         const void *__restrict__ pos, int pos_dtype, void *__restrict__ y_lp, void *__restrict__ q_lp) {
    constexpr int d = 128 * PP;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    const size_t base = (size_t)row * d;
    float4 v[PP];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PP; ++i) {
        const size_t idx = base + (size_t)(i * 32 + lane) * 4;
        v[i] = ln_load4(x, x_dtype, idx);
        if (res != nullptr) {
            const float4 r = ln_load4(res, res_dtype, idx);
            v[i].x += r.x; v[i].y += r.y; v[i].z += r.z; v[i].w += r.w;
        }
        if (z != nullptr) *reinterpret_cast<float4 *>(z + idx) = v[i];
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mu = warp_sum(s) * (1.0f / d);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PP; ++i) {
        const float a = v[i].x - mu, c = v[i].y - mu, e = v[i].z - mu, f = v[i].w - mu;
        q += (a * a + c * c) + (e * e + f * f);
    }
    const float rs = rsqrtf(warp_sum(q) * (1.0f / d) + eps);
    if (lane == 0) {
        if (mean != nullptr) mean[row] = mu;
        if (rstd != nullptr) rstd[row] = rs;
    }
#pragma unroll
    for (int i = 0; i < PP; ++i) {
        const int col = (i * 32 + lane) * 4;
        const float4 ww = *reinterpret_cast<const float4 *>(w + col), bb = *reinterpret_cast<const float4 *>(b + col);
        float4 o;
        o.x = fmaf((v[i].x - mu) * rs, ww.x, bb.x);
        o.y = fmaf((v[i].y - mu) * rs, ww.y, bb.y);
        o.z = fmaf((v[i].z - mu) * rs, ww.z, bb.z);
        o.w = fmaf((v[i].w - mu) * rs, ww.w, bb.w);
        ln_store4(y, y_dtype, base + col, o);
        // side outputs for the consumers that compute in bf16 (what autocast's casts of y and of y + pos would produce)
        if (y_lp != nullptr) ln_store4(y_lp, TAMTR_BF16, base + col, o);
        if (q_lp != nullptr) {
            const float4 p = ln_load4(pos, pos_dtype, base + col);
            ln_store4(q_lp, TAMTR_BF16, base + col, make_float4(o.x + p.x, o.y + p.y, o.z + p.z, o.w + p.w));
        }
    }
}

template <int PP>
__global__ void __launch_bounds__(kLnThreads)
add_layernorm_bwd_kernel(const void *__restrict__ dy, int dy_dtype, const float *__restrict__ z,
                         const float *__restrict__ mean, const float *__restrict__ rstd, const float *__restrict__ w,
                         void *__restrict__ dx, int dx_dtype, void *__restrict__ dres, int dres_dtype,
                         float *__restrict__ dwb, int rows, const float *__restrict__ extra,
                         const void *__restrict__ g1, const void *__restrict__ g2) {
    constexpr int d = 128 * PP;
    __shared__ float s_red[kLnWarps][2][d];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 ww[PP], aw[PP], ab[PP];
#pragma unroll
    for (int i = 0; i < PP; ++i) {
        ww[i] = *reinterpret_cast<const float4 *>(w + (i * 32 + lane) * 4);
        aw[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int row = blockIdx.x * kLnWarps + warp; row < rows; row += gridDim.x * kLnWarps) {
        const size_t base = (size_t)row * d;
        const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
        float4 g[PP], xh[PP];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < PP; ++i) {
            const size_t idx = base + (size_t)(i * 32 + lane) * 4;
            float4 gy = dy != nullptr ? ln_load4(dy, dy_dtype, idx) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (g1 != nullptr) {                       // gradients of the bf16 side outputs (y_lp, q_lp): same dy
                const float4 e = ln_load4(g1, TAMTR_BF16, idx);
                gy.x += e.x; gy.y += e.y; gy.z += e.z; gy.w += e.w;
            }
            if (g2 != nullptr) {
                const float4 e = ln_load4(g2, TAMTR_BF16, idx);
                gy.x += e.x; gy.y += e.y; gy.z += e.z; gy.w += e.w;
            }
            const float4 zz = *reinterpret_cast<const float4 *>(z + idx);
            xh[i] = make_float4((zz.x - mu) * rs, (zz.y - mu) * rs, (zz.z - mu) * rs, (zz.w - mu) * rs);
            g[i] = make_float4(gy.x * ww[i].x, gy.y * ww[i].y, gy.z * ww[i].z, gy.w * ww[i].w);
            aw[i].x = fmaf(gy.x, xh[i].x, aw[i].x); aw[i].y = fmaf(gy.y, xh[i].y, aw[i].y);
            aw[i].z = fmaf(gy.z, xh[i].z, aw[i].z); aw[i].w = fmaf(gy.w, xh[i].w, aw[i].w);
            ab[i].x += gy.x; ab[i].y += gy.y; ab[i].z += gy.z; ab[i].w += gy.w;
            s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
            s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
        }
        const float m1 = warp_sum(s1) * (1.0f / d), m2 = warp_sum(s2) * (1.0f / d);
#pragma unroll
        for (int i = 0; i < PP; ++i) {
            const size_t idx = base + (size_t)(i * 32 + lane) * 4;
            float4 o;
            o.x = rs * (g[i].x - m1 - xh[i].x * m2);
            o.y = rs * (g[i].y - m1 - xh[i].y * m2);
            o.z = rs * (g[i].z - m1 - xh[i].z * m2);
            o.w = rs * (g[i].w - m1 - xh[i].w * m2);
            if (extra != nullptr) {                    // gradient reaching z on a second path (a residual connection)
                const float4 e = *reinterpret_cast<const float4 *>(extra + idx);
                o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
            }
            if (dx != nullptr) ln_store4(dx, dx_dtype, idx, o);
            if (dres != nullptr) ln_store4(dres, dres_dtype, idx, o);
        }
    }
    // column sums: warps -> CTA (shared memory) -> one fp32 reduction per column into dw / db (zeroed by the launcher).
    // (A deterministic "last CTA folds the partials" pass was tried first: one SM pulling every partial took longer
    // than the rest of the kernel.)
#pragma unroll
    for (int i = 0; i < PP; ++i) {
        *reinterpret_cast<float4 *>(&s_red[warp][0][(i * 32 + lane) * 4]) = aw[i];
        *reinterpret_cast<float4 *>(&s_red[warp][1][(i * 32 + lane) * 4]) = ab[i];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 2 * d; j += kLnThreads) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < kLnWarps; ++k) s += (&s_red[k][0][0])[j];
        atomicAdd(dwb + j, s);
    }
}

// q_lp = bf16(x + pos), x_lp = bf16(x): the casts autocast puts in front of the self-attention projections
// (transformer.py:544-547: q = k = embed + query_pos, v = embed) as one pass over x.
__global__ void __launch_bounds__(256) pos_cast_kernel(const void *__restrict__ x, int x_dtype, const void *__restrict__ pos,
                                                       int pos_dtype, void *__restrict__ x_lp, void *__restrict__ q_lp,
                                                       size_t n4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = ln_load4(x, x_dtype, i * 4);
        if (x_lp != nullptr) ln_store4(x_lp, TAMTR_BF16, i * 4, v);
        if (q_lp != nullptr) {
            const float4 p = ln_load4(pos, pos_dtype, i * 4);
            ln_store4(q_lp, TAMTR_BF16, i * 4, make_float4(v.x + p.x, v.y + p.y, v.z + p.z, v.w + p.w));
        }
    }
}

// out = a + b + c (each optional but one; fp32 sum): the gradient of a tensor that fed pos_cast and a residual connection
__global__ void __launch_bounds__(256) grad_sum3_kernel(const void *__restrict__ a, int a_dtype, const void *__restrict__ b,
                                                        int b_dtype, const void *__restrict__ c, int c_dtype,
                                                        void *__restrict__ out, int out_dtype, size_t n4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a != nullptr) s = ln_load4(a, a_dtype, i * 4);
        if (b != nullptr) {
            const float4 e = ln_load4(b, b_dtype, i * 4);
            s.x += e.x; s.y += e.y; s.z += e.z; s.w += e.w;
        }
        if (c != nullptr) {
            const float4 e = ln_load4(c, c_dtype, i * 4);
            s.x += e.x; s.y += e.y; s.z += e.z; s.w += e.w;
        }
        ln_store4(out, out_dtype, i * 4, s);
    }
}

static int ln_ctas(int rows) {
    int n_sm = 148;
    n_sm = ::tamtr::sm_count();
    const int need = (rows + kLnWarps - 1) / kLnWarps;
    return need < 2 * n_sm ? need : 2 * n_sm;
}

}  // namespace tamtr

using namespace tamtr;

static int add_layernorm_forward_impl(const void *x, int x_dtype, const void *res, int res_dtype, const float *w,
                                      const float *b, void *y, int y_dtype, float *z, float *mean, float *rstd,
                                      const void *pos, int pos_dtype, void *y_lp, void *q_lp, int rows, int d, float eps,
                                      void *stream) {
    TAMTR_CHECK_ARG(x && w && b && y, TAMTR_E_BADARG, "add_layernorm_forward: null pointer");
    TAMTR_CHECK_ARG(q_lp == nullptr || pos != nullptr, TAMTR_E_BADARG, "add_layernorm_forward: q_lp needs pos");
    TAMTR_CHECK_ARG(rows > 0, TAMTR_E_BADARG, "add_layernorm_forward: rows = %d", rows);
    TAMTR_CHECK_ARG(d % 128 == 0 && d >= 128 && d <= 512, TAMTR_E_UNSUPPORTED,
                    "add_layernorm: d = %d must be a multiple of 128 in [128, 512]", d);
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = (rows + kLnWarps - 1) / kLnWarps;
    {
        KernelTimer timer(K_LN_FWD, st);
#define TAMTR_LN_FWD(PP)                                                                                              \
    case PP:                                                                                                          \
        add_layernorm_fwd_kernel<PP><<<grid, kLnThreads, 0, st>>>(x, x_dtype, res, res_dtype, w, b, y, y_dtype, z,   \
                                                                  mean, rstd, rows, eps, pos, pos_dtype, y_lp, q_lp); \
        break;
        switch (d / 128) {
            TAMTR_LN_FWD(1) TAMTR_LN_FWD(2) TAMTR_LN_FWD(3) TAMTR_LN_FWD(4)
        }
#undef TAMTR_LN_FWD
    }
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_add_layernorm_forward(const void *x, int x_dtype, const void *res, int res_dtype, const float *w,
                                           const float *b, void *y, int y_dtype, float *z, float *mean, float *rstd,
                                           int rows, int d, float eps, void *stream) {
    return add_layernorm_forward_impl(x, x_dtype, res, res_dtype, w, b, y, y_dtype, z, mean, rstd, nullptr, 0, nullptr,
                                      nullptr, rows, d, eps, stream);
}

extern "C" int tamtr_add_layernorm_forward_sides(const void *x, int x_dtype, const void *res, int res_dtype, const float *w,
                                                 const float *b, void *y, int y_dtype, float *z, float *mean, float *rstd,
                                                 const void *pos, int pos_dtype, void *y_bf16, void *q_bf16, int rows,
                                                 int d, float eps, void *stream) {
    return add_layernorm_forward_impl(x, x_dtype, res, res_dtype, w, b, y, y_dtype, z, mean, rstd, pos, pos_dtype, y_bf16,
                                      q_bf16, rows, d, eps, stream);
}

static int add_layernorm_backward_impl(const void *dy, int dy_dtype, const float *z, const float *mean, const float *rstd,
                                       const float *w, const float *extra, void *dx, int dx_dtype, void *dres,
                                       int dres_dtype, float *dwb, int rows, int d, void *stream,
                                       const void *g1 = nullptr, const void *g2 = nullptr) {
    TAMTR_CHECK_ARG((dy || g1 || g2) && z && mean && rstd && w && dwb, TAMTR_E_BADARG,
                    "add_layernorm_backward: null pointer");
    TAMTR_CHECK_ARG(rows > 0, TAMTR_E_BADARG, "add_layernorm_backward: rows = %d", rows);
    TAMTR_CHECK_ARG(d % 128 == 0 && d >= 128 && d <= 512, TAMTR_E_UNSUPPORTED,
                    "add_layernorm: d = %d must be a multiple of 128 in [128, 512]", d);
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = ln_ctas(rows);
    TAMTR_CUDA_OK(cudaMemsetAsync(dwb, 0, (size_t)2 * d * sizeof(float), st));
    {
        KernelTimer timer(K_LN_BWD, st);
#define TAMTR_LN_BWD(PP)                                                                                              \
    case PP:                                                                                                          \
        add_layernorm_bwd_kernel<PP><<<grid, kLnThreads, 0, st>>>(dy, dy_dtype, z, mean, rstd, w, dx, dx_dtype, dres, \
                                                                  dres_dtype, dwb, rows, extra, g1, g2);            \
        break;
        switch (d / 128) {
            TAMTR_LN_BWD(1) TAMTR_LN_BWD(2) TAMTR_LN_BWD(3) TAMTR_LN_BWD(4)
        }
#undef TAMTR_LN_BWD
    }
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_add_layernorm_backward(const void *dy, int dy_dtype, const float *z, const float *mean,
                                            const float *rstd, const float *w, void *dx, int dx_dtype, void *dres,
                                            int dres_dtype, float *dwb, int rows, int d, void *stream) {
    return add_layernorm_backward_impl(dy, dy_dtype, z, mean, rstd, w, nullptr, dx, dx_dtype, dres, dres_dtype, dwb, rows, d,
                                       stream);
}

extern "C" int tamtr_add_layernorm_backward_res(const void *dy, int dy_dtype, const float *z, const float *mean,
                                                const float *rstd, const float *w, const float *extra, void *dx,
                                                int dx_dtype, void *dres, int dres_dtype, float *dwb, int rows, int d,
                                                void *stream) {
    TAMTR_CHECK_ARG(extra != nullptr, TAMTR_E_BADARG, "add_layernorm_backward_res: null pointer");
    return add_layernorm_backward_impl(dy, dy_dtype, z, mean, rstd, w, extra, dx, dx_dtype, dres, dres_dtype, dwb, rows, d,
                                       stream);
}

extern "C" int tamtr_add_layernorm_backward_sides(const void *dy, int dy_dtype, const void *g_y_bf16, const void *g_q_bf16,
                                                  const float *z, const float *mean, const float *rstd, const float *w,
                                                  void *dx, int dx_dtype, void *dres, int dres_dtype, float *dwb, int rows,
                                                  int d, void *stream) {
    return add_layernorm_backward_impl(dy, dy_dtype, z, mean, rstd, w, nullptr, dx, dx_dtype, dres, dres_dtype, dwb, rows, d,
                                       stream, g_y_bf16, g_q_bf16);
}

static int ew_grid(size_t n4) {
    const size_t need = (n4 + 255) / 256;
    const size_t cap = (size_t)::tamtr::sm_count() * 8;
    return (int)(need < cap ? need : cap);
}

extern "C" int tamtr_pos_cast(const void *x, int x_dtype, const void *pos, int pos_dtype, void *x_bf16, void *q_bf16,
                              long n, void *stream) {
    TAMTR_CHECK_ARG(x && (x_bf16 || q_bf16), TAMTR_E_BADARG, "pos_cast: null pointer");
    TAMTR_CHECK_ARG(q_bf16 == nullptr || pos != nullptr, TAMTR_E_BADARG, "pos_cast: q needs pos");
    TAMTR_CHECK_ARG(n > 0 && n % 4 == 0, TAMTR_E_UNSUPPORTED, "pos_cast: n = %ld must be a positive multiple of 4", n);
    pos_cast_kernel<<<ew_grid((size_t)n / 4), 256, 0, (cudaStream_t)stream>>>(x, x_dtype, pos, pos_dtype, x_bf16, q_bf16,
                                                                              (size_t)n / 4);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_grad_sum3(const void *a, int a_dtype, const void *b, int b_dtype, const void *c, int c_dtype, void *out,
                               int out_dtype, long n, void *stream) {
    TAMTR_CHECK_ARG(out && (a || b || c), TAMTR_E_BADARG, "grad_sum3: null pointer");
    TAMTR_CHECK_ARG(n > 0 && n % 4 == 0, TAMTR_E_UNSUPPORTED, "grad_sum3: n = %ld must be a positive multiple of 4", n);
    grad_sum3_kernel<<<ew_grid((size_t)n / 4), 256, 0, (cudaStream_t)stream>>>(a, a_dtype, b, b_dtype, c, c_dtype, out,
                                                                               out_dtype, (size_t)n / 4);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}
