// Depth-wise 3x3 convolution + bias + SiLU of SS2D (ultralytics/nn/extra_modules/VManba/vmamba.py:1026-1027:
// `x = self.act(self.conv2d(x))`, conv2d = Conv2d(d_inner, d_inner, 3, padding=1, groups=d_inner), act = SiLU), forward and
// backward, NCHW.  The op is a 9-tap stencil per channel plane -- 2 bytes in, 2 bytes out per element, HBM-bound -- but the
// library's depth-wise kernels take 0.9 ms forward and 2.7 ms backward on the head's largest level (B=16, 256 x 160 x 160
// bf16: 210 MB each way, 65 us at HBM peak) and the SiLU is two more passes.  Here a CTA stages a 16-row x 128-column
// tile (+ halo) of one plane in shared memory as
// fp32 and every thread slides a 3x3 window down its column.
//   forward : y = silu(conv(x) + b)
//   backward: pre = conv(x) + b is recomputed from the staged x tile (x is needed for the weight gradient anyway),
//             gp = g * silu'(pre);  dx = conv^T(gp);  dw[c, dy, dx] = sum gp * x(shifted);  db[c] = sum gp
//             (fp32 partial sums per thread -> warp shuffles -> one atomicAdd per value and CTA)
#include "common.cuh"

namespace tamtr {

// tile = kDwTH rows x TW columns (TW = 128; 64 / 32 for maps that narrow); 256 threads = (256 / TW) row groups x TW columns
constexpr int kDwTH = 16, kDwThreads = 256;

template <typename T> __device__ __forceinline__ float dw_ld(const T *p);
template <> __device__ __forceinline__ float dw_ld<float>(const float *p) { return __ldg(p); }
template <> __device__ __forceinline__ float dw_ld<__nv_bfloat16>(const __nv_bfloat16 *p) {
    return __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16 *>(p)));
}
template <typename T> __device__ __forceinline__ void dw_st(T *p, float v);
template <> __device__ __forceinline__ void dw_st<float>(float *p, float v) { *p = v; }
template <> __device__ __forceinline__ void dw_st<__nv_bfloat16>(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }

// plane tile [h0 - HALO, h0 + kDwTH + HALO) x [w0 - HALO, w0 + TW + HALO) of `src` -> smem (zeros outside the image);
// one warp per tile row, lanes along the row
template <typename T, int HALO, int TW>
__device__ __forceinline__ void dw_stage(float (*tile)[TW + 2 * HALO + 1], const T *__restrict__ src, int H, int W, int h0,
                                         int w0) {
    constexpr int TWH = TW + 2 * HALO, THH = kDwTH + 2 * HALO;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < THH; r += kDwThreads / 32) {
        const int h = h0 - HALO + r;
        const bool row_ok = h >= 0 && h < H;
        const T *row = src + (size_t)(row_ok ? h : 0) * W;
        for (int c = lane; c < TWH; c += 32) {
            const int w = w0 - HALO + c;
            tile[r][c] = (row_ok && w >= 0 && w < W) ? dw_ld(row + w) : 0.0f;
        }
    }
}

template <typename T, int TW>
__global__ void __launch_bounds__(kDwThreads)
dwconv3x3_silu_fwd_kernel(const T *__restrict__ x, const float *__restrict__ wgt, const float *__restrict__ bias,
                          T *__restrict__ y, int D, int H, int W) {
    constexpr int RPG = kDwTH * TW / kDwThreads;       // rows per thread: 2 / 4 / 8
    __shared__ float xs[kDwTH + 2][TW + 3];
    const int plane = blockIdx.x, c = plane % D;                 // planes on grid.x (no 65 535 limit)
    const int h0 = blockIdx.z * kDwTH, w0 = blockIdx.y * TW;
    const T *src = x + (size_t)plane * H * W;
    dw_stage<T, 1, TW>(xs, src, H, W, h0, w0);
    float k[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) k[i] = __ldg(wgt + c * 9 + i);
    const float b = bias ? __ldg(bias + c) : 0.0f;
    __syncthreads();
    const int col = threadIdx.x % TW, r0 = (threadIdx.x / TW) * RPG;
    if (w0 + col >= W) return;
    T *dst = y + (size_t)plane * H * W + w0 + col;
    float a[3], m[3], n[3];                      // window rows r-1, r, r+1 (tile coordinates: +1 halo)
#pragma unroll
    for (int j = 0; j < 3; ++j) { a[j] = xs[r0][col + j]; m[j] = xs[r0 + 1][col + j]; }
#pragma unroll
    for (int r = 0; r < RPG; ++r) {
#pragma unroll
        for (int j = 0; j < 3; ++j) n[j] = xs[r0 + r + 2][col + j];
        float pre = b;
#pragma unroll
        for (int j = 0; j < 3; ++j) pre = fmaf(k[j], a[j], fmaf(k[3 + j], m[j], fmaf(k[6 + j], n[j], pre)));
        const int h = h0 + r0 + r;
        if (h < H) dw_st(dst + (size_t)h * W, pre / (1.0f + __expf(-pre)));
#pragma unroll
        for (int j = 0; j < 3; ++j) { a[j] = m[j]; m[j] = n[j]; }
    }
}

template <typename T, int TW>
__global__ void __launch_bounds__(kDwThreads)
dwconv3x3_silu_bwd_kernel(const T *__restrict__ g, const T *__restrict__ x, const float *__restrict__ wgt,
                          const float *__restrict__ bias, T *__restrict__ gx, float *__restrict__ gw, float *__restrict__ gb,
                          int D, int H, int W) {
    constexpr int RPG = kDwTH * TW / kDwThreads;
    __shared__ float xs[kDwTH + 4][TW + 5];          // x, halo 2
    __shared__ float gs[kDwTH + 2][TW + 3];          // g, then gp = g * silu'(pre), halo 1
    __shared__ float red[kDwThreads / 32][10];
    const int plane = blockIdx.x, c = plane % D;                 // planes on grid.x (no 65 535 limit)
    const int h0 = blockIdx.z * kDwTH, w0 = blockIdx.y * TW;
    dw_stage<T, 2, TW>(xs, x + (size_t)plane * H * W, H, W, h0, w0);
    dw_stage<T, 1, TW>(gs, g + (size_t)plane * H * W, H, W, h0, w0);
    float k[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) k[i] = __ldg(wgt + c * 9 + i);
    const float b = bias ? __ldg(bias + c) : 0.0f;
    __syncthreads();
    // gp on the tile + halo 1 (outside the image g == 0, hence gp == 0)
    constexpr int TW1 = TW + 2, TH1 = kDwTH + 2;
    for (int r = threadIdx.x >> 5; r < TH1; r += kDwThreads / 32) {          // gs coordinates; xs coordinates are +1
        for (int cc = threadIdx.x & 31; cc < TW1; cc += 32) {
            float pre = b;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) pre = fmaf(k[dy * 3 + dx], xs[r + dy][cc + dx], pre);
            const float s = 1.0f / (1.0f + __expf(-pre));
            gs[r][cc] *= s * (1.0f + pre * (1.0f - s));
        }
    }
    __syncthreads();
    const int col = threadIdx.x % TW, r0 = (threadIdx.x / TW) * RPG;
    float acc[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) acc[i] = 0.0f;
    if (w0 + col < W) {
        T *dst = gx + (size_t)plane * H * W + w0 + col;
#pragma unroll 2
        for (int r = 0; r < RPG; ++r) {
            const int h = h0 + r0 + r;
            if (h >= H) break;
            const int gr = r0 + r + 1, gc = col + 1;          // this position in gs; in xs it is (+2, +2)
            // dx[h, w] = sum_{dy,dx} k[dy][dx] * gp[h - dy + 1][w - dx + 1]
            float d = 0.0f;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) d = fmaf(k[dy * 3 + dx], gs[gr - dy + 1][gc - dx + 1], d);
            dw_st(dst + (size_t)h * W, d);
            const float gp = gs[gr][gc];
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) acc[dy * 3 + dx] = fmaf(gp, xs[gr + dy][gc + dx], acc[dy * 3 + dx]);
            acc[9] += gp;
        }
    }
#pragma unroll
    for (int i = 0; i < 10; ++i) {
#pragma unroll
        for (int mlane = 16; mlane > 0; mlane >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], mlane);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 10; ++i) red[warp][i] = acc[i];
    }
    __syncthreads();
    if (threadIdx.x < 10) {
        float s = 0.0f;
        for (int wv = 0; wv < kDwThreads / 32; ++wv) s += red[wv][threadIdx.x];
        if (threadIdx.x < 9) atomicAdd(gw + c * 9 + threadIdx.x, s);
        else if (gb) atomicAdd(gb + c, s);
    }
}

static int dw_check(const void *a, const void *b, const void *c, int dtype, int Bn, int D, int H, int W) {
    TAMTR_CHECK_ARG(a && b && c, TAMTR_E_BADARG, "dwconv3x3_silu: null pointer");
    TAMTR_CHECK_ARG(Bn > 0 && D > 0 && H > 0 && W > 0, TAMTR_E_BADARG, "dwconv3x3_silu: non-positive size");
    TAMTR_CHECK_ARG(dtype == TAMTR_F32 || dtype == TAMTR_BF16, TAMTR_E_UNSUPPORTED, "dwconv3x3_silu: dtype %d", dtype);
    TAMTR_CHECK_ARG((long)Bn * D <= 2147483647L && (H + kDwTH - 1) / kDwTH <= 65535 && (W + 31) / 32 <= 65535,
                    TAMTR_E_UNSUPPORTED, "dwconv3x3_silu: too many planes or tiles");
    return 0;
}

}  // namespace tamtr

using namespace tamtr;

// 128-column tiles unless the map is narrower.  (Picking the width with the fewest padded columns -- 5 x 32 for 160-wide
// maps instead of 2 x 128 -- was measured SLOWER, 0.71 vs 0.47 ms forward: two rows per thread lose the sliding-window
// reuse and the row segments shrink to 64 bytes.)
static int dw_tile_width(int W) { return W <= 32 ? 32 : W <= 64 ? 64 : 128; }

extern "C" int tamtr_dwconv3x3_silu_forward(const void *x, const float *weight, const float *bias, void *y, int dtype, int Bn,
                                            int D, int H, int W, void *stream) {
    const int rc = dw_check(x, weight, y, dtype, Bn, D, H, W);
    if (rc) return rc;
    const int tw = dw_tile_width(W);
    const dim3 grid(Bn * D, (W + tw - 1) / tw, (H + kDwTH - 1) / kDwTH);
    cudaStream_t st = (cudaStream_t)stream;
#define DW_FWD(T, TWV) dwconv3x3_silu_fwd_kernel<T, TWV><<<grid, kDwThreads, 0, st>>>((const T *)x, weight, bias, (T *)y, D, H, W)
    if (dtype == TAMTR_F32) { if (tw == 32) DW_FWD(float, 32); else if (tw == 64) DW_FWD(float, 64); else DW_FWD(float, 128); }
    else { if (tw == 32) DW_FWD(__nv_bfloat16, 32); else if (tw == 64) DW_FWD(__nv_bfloat16, 64); else DW_FWD(__nv_bfloat16, 128); }
#undef DW_FWD
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_dwconv3x3_silu_backward(const void *grad_y, const void *x, const float *weight, const float *bias,
                                             void *grad_x, float *grad_weight, float *grad_bias, int dtype, int Bn, int D,
                                             int H, int W, void *stream) {
    const int rc = dw_check(grad_y, x, grad_x, dtype, Bn, D, H, W);
    if (rc) return rc;
    TAMTR_CHECK_ARG(weight && grad_weight, TAMTR_E_BADARG, "dwconv3x3_silu_backward: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    TAMTR_CUDA_OK(cudaMemsetAsync(grad_weight, 0, sizeof(float) * 9 * (size_t)D, st));
    if (grad_bias) TAMTR_CUDA_OK(cudaMemsetAsync(grad_bias, 0, sizeof(float) * (size_t)D, st));
    const int tw = dw_tile_width(W);
    const dim3 grid(Bn * D, (W + tw - 1) / tw, (H + kDwTH - 1) / kDwTH);
#define DW_BWD(T, TWV)                                                                                                      \
    dwconv3x3_silu_bwd_kernel<T, TWV><<<grid, kDwThreads, 0, st>>>((const T *)grad_y, (const T *)x, weight, bias, (T *)grad_x, \
                                                                   grad_weight, grad_bias, D, H, W)
    if (dtype == TAMTR_F32) { if (tw == 32) DW_BWD(float, 32); else if (tw == 64) DW_BWD(float, 64); else DW_BWD(float, 128); }
    else { if (tw == 32) DW_BWD(__nv_bfloat16, 32); else if (tw == 64) DW_BWD(__nv_bfloat16, 64); else DW_BWD(__nv_bfloat16, 128); }
#undef DW_BWD
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}
