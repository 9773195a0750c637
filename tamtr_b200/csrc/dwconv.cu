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

// ---- version 2 (maps whose width is a multiple of 4): no shared-memory tile.  A thread owns a strip of 4 columns and walks
// down a band of rows with the 3x3 neighbourhood in registers: per row one vector load of its 4 elements plus the halo
// elements left and right (L1 hits: the neighbouring strips' data), 36 FMAs, one vector store.  A CTA works on ONE channel
// (grid.x) for a slice of the batch (grid.y), its threads striding over (image, band, strip) items -- consecutive threads
// take consecutive strips, so a warp reads and writes whole row segments; the filter sits in registers and the weight /
// bias gradients are reduced once per CTA.  (The tiled version above staged 2-byte scalars into a 16 x 128 fp32 tile and
// re-read 9 shared-memory values per output: 0.45 ms forward / 1.68 ms backward at the head's largest level against
// 65 / 130 us of HBM time; on 160-wide maps its second column tile was three quarters padding.)
template <typename T> struct DwVec;
template <> struct DwVec<float> {
    static __device__ __forceinline__ void ld4(const float *p, float (&v)[4]) {
        const float4 q = __ldg(reinterpret_cast<const float4 *>(p));
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
    static __device__ __forceinline__ void ld2(const float *p, float &a, float &b) {
        const float2 q = __ldg(reinterpret_cast<const float2 *>(p));
        a = q.x; b = q.y;
    }
    static __device__ __forceinline__ void st4(float *p, const float (&v)[4]) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct DwVec<__nv_bfloat16> {
    static __device__ __forceinline__ void ld4(const __nv_bfloat16 *p, float (&v)[4]) {
        const uint2 q = __ldg(reinterpret_cast<const uint2 *>(p));
        v[0] = __uint_as_float(q.x << 16); v[1] = __uint_as_float(q.x & 0xffff0000u);
        v[2] = __uint_as_float(q.y << 16); v[3] = __uint_as_float(q.y & 0xffff0000u);
    }
    static __device__ __forceinline__ void ld2(const __nv_bfloat16 *p, float &a, float &b) {
        const uint32_t q = __ldg(reinterpret_cast<const uint32_t *>(p));
        a = __uint_as_float(q << 16); b = __uint_as_float(q & 0xffff0000u);
    }
    static __device__ __forceinline__ void st4(__nv_bfloat16 *p, const float (&v)[4]) {
        const __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
        *reinterpret_cast<uint2 *>(p) = make_uint2(*reinterpret_cast<const uint32_t *>(&lo), *reinterpret_cast<const uint32_t *>(&hi));
    }
};

// columns c0 - HALO .. c0 + 3 + HALO of row r (zeros outside the image); HALO = 1 or 2, c0 % 4 == 0, W % 4 == 0
template <typename T, int HALO>
__device__ __forceinline__ void dw_row(const T *__restrict__ plane, int r, int H, int W, int c0, float (&v)[4 + 2 * HALO]) {
    if (r < 0 || r >= H) {
#pragma unroll
        for (int j = 0; j < 4 + 2 * HALO; ++j) v[j] = 0.0f;
        return;
    }
    const T *row = plane + (size_t)r * W + c0;
    float m[4];
    DwVec<T>::ld4(row, m);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[HALO + j] = m[j];
    if (HALO == 2) {
        float a = 0.0f, b = 0.0f, c = 0.0f, d = 0.0f;
        if (c0 > 0) DwVec<T>::ld2(row - 2, a, b);
        if (c0 + 4 < W) DwVec<T>::ld2(row + 4, c, d);
        v[0] = a; v[1] = b; v[HALO + 4] = c; v[HALO + 5] = d;
    } else {
        v[0] = c0 > 0 ? dw_ld(row - 1) : 0.0f;
        v[HALO + 4] = c0 + 4 < W ? dw_ld(row + 4) : 0.0f;
    }
}

constexpr int kDw2Threads = 256;

template <typename T>
__global__ void __launch_bounds__(kDw2Threads)
dwconv3x3_silu_fwd2_kernel(const T *__restrict__ x, const float *__restrict__ wgt, const float *__restrict__ bias,
                           T *__restrict__ y, int Bn, int D, int H, int W, int RB, int imgs_per_cta) {
    const int c = blockIdx.x, S = W >> 2, NB = (H + RB - 1) / RB;
    float k[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) k[i] = __ldg(wgt + c * 9 + i);
    const float b = bias ? __ldg(bias + c) : 0.0f;
    const int items = imgs_per_cta * NB * S;
    for (int i = threadIdx.x; i < items; i += kDw2Threads) {
        const int strip = i % S, band = (i / S) % NB, img = blockIdx.y * imgs_per_cta + i / (S * NB);
        if (img >= Bn) break;
        const size_t po = ((size_t)img * D + c) * H * W;
        const T *plane = x + po;
        T *out = y + po;
        const int c0 = strip << 2, r0 = band * RB, r1 = min(H, r0 + RB);
        float a[6], m[6], n[6];
        dw_row<T, 1>(plane, r0 - 1, H, W, c0, a);
        dw_row<T, 1>(plane, r0, H, W, c0, m);
        for (int r = r0; r < r1; ++r) {
            dw_row<T, 1>(plane, r + 1, H, W, c0, n);
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float pre = b;
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) pre = fmaf(k[dx], a[j + dx], fmaf(k[3 + dx], m[j + dx], fmaf(k[6 + dx], n[j + dx], pre)));
                o[j] = __fdividef(pre, 1.0f + __expf(-pre));
            }
            DwVec<T>::st4(out + (size_t)r * W + c0, o);
#pragma unroll
            for (int j = 0; j < 6; ++j) { a[j] = m[j]; m[j] = n[j]; }
        }
    }
}

// backward: gp = g * silu'(conv(x) + b) is formed row by row on the strip + 1 halo column each side (x window: 3 rows x 8
// columns); dx of row q - 1 needs gp rows q - 2 .. q; dw / db use the strip's own 4 columns of gp row q.
template <typename T>
__global__ void __launch_bounds__(kDw2Threads)
dwconv3x3_silu_bwd2_kernel(const T *__restrict__ g, const T *__restrict__ x, const float *__restrict__ wgt,
                           const float *__restrict__ bias, T *__restrict__ gx, float *__restrict__ gw, float *__restrict__ gb,
                           int Bn, int D, int H, int W, int RB, int imgs_per_cta) {
    __shared__ float red[kDw2Threads / 32][10];
    const int c = blockIdx.x, S = W >> 2, NB = (H + RB - 1) / RB;
    float k[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) k[i] = __ldg(wgt + c * 9 + i);
    const float b = bias ? __ldg(bias + c) : 0.0f;
    float acc[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) acc[i] = 0.0f;
    const int items = imgs_per_cta * NB * S;
    for (int i = threadIdx.x; i < items; i += kDw2Threads) {
        const int strip = i % S, band = (i / S) % NB, img = blockIdx.y * imgs_per_cta + i / (S * NB);
        if (img >= Bn) break;
        const size_t po = ((size_t)img * D + c) * H * W;
        const T *xp = x + po, *gpl = g + po;
        T *out = gx + po;
        const int c0 = strip << 2, r0 = band * RB, r1 = min(H, r0 + RB);
        float xa[8], xm[8], xn[8];                     // x rows q - 1, q, q + 1; columns c0 - 2 .. c0 + 5
        float pa[6], pm[6], pn[6];                     // gp rows q - 2, q - 1, q; columns c0 - 1 .. c0 + 4
#pragma unroll
        for (int j = 0; j < 6; ++j) pa[j] = pm[j] = 0.0f;
        dw_row<T, 2>(xp, r0 - 2, H, W, c0, xa);
        dw_row<T, 2>(xp, r0 - 1, H, W, c0, xm);
        for (int q = r0 - 1; q <= r1; ++q) {           // gp row q; dx row q - 1
            dw_row<T, 2>(xp, q + 1, H, W, c0, xn);
            float gq[6];
            dw_row<T, 1>(gpl, q, H, W, c0, gq);        // zeros outside the image: gp = 0 there
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                float pre = b;
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) pre = fmaf(k[dx], xa[j + dx], fmaf(k[3 + dx], xm[j + dx], fmaf(k[6 + dx], xn[j + dx], pre)));
                const float s = __fdividef(1.0f, 1.0f + __expf(-pre));
                pn[j] = gq[j] * s * fmaf(pre, 1.0f - s, 1.0f);
            }
            if (q >= r0 && q < r1) {                   // weight / bias gradient: the band's own rows and columns
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float gp = pn[j + 1];
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        acc[dx] = fmaf(gp, xa[j + 1 + dx], acc[dx]);
                        acc[3 + dx] = fmaf(gp, xm[j + 1 + dx], acc[3 + dx]);
                        acc[6 + dx] = fmaf(gp, xn[j + 1 + dx], acc[6 + dx]);
                    }
                    acc[9] += gp;
                }
            }
            if (q > r0) {                              // dx[q-1][c] = sum k[dy][dx] * gp[q - 1 - dy + 1][c - dx + 1]
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float d = 0.0f;
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx)
                        d = fmaf(k[dx], pn[j + 2 - dx], fmaf(k[3 + dx], pm[j + 2 - dx], fmaf(k[6 + dx], pa[j + 2 - dx], d)));
                    o[j] = d;
                }
                DwVec<T>::st4(out + (size_t)(q - 1) * W + c0, o);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) { xa[j] = xm[j]; xm[j] = xn[j]; }
#pragma unroll
            for (int j = 0; j < 6; ++j) { pa[j] = pm[j]; pm[j] = pn[j]; }
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
        for (int wv = 0; wv < kDw2Threads / 32; ++wv) s += red[wv][threadIdx.x];
        if (threadIdx.x < 9) atomicAdd(gw + c * 9 + threadIdx.x, s);
        else if (gb) atomicAdd(gb + c, s);
    }
}

// rows per band and images per CTA of version 2: enough CTAs for ~4 waves, bands of at most 32 rows
static void dw2_geometry(int Bn, int D, int H, int W, int &RB, int &imgs_per_cta, dim3 &grid) {
    RB = H < 32 ? H : 32;
    int split = 1;
    while (split < Bn && (long)D * split < 4L * ::tamtr::sm_count()) split *= 2;
    imgs_per_cta = (Bn + split - 1) / split;
    grid = dim3(D, (Bn + imgs_per_cta - 1) / imgs_per_cta);
}
static bool dw2_ok(const void *a, const void *b, const void *c, int D, int W) {
    return W % 4 == 0 && D <= 65535 * 32 && ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)c)) & 15) == 0;
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
    cudaStream_t st = (cudaStream_t)stream;
    if (dw2_ok(x, y, y, D, W)) {
        int RB, ipc;
        dim3 g2;
        dw2_geometry(Bn, D, H, W, RB, ipc, g2);
        if (dtype == TAMTR_F32)
            dwconv3x3_silu_fwd2_kernel<float><<<g2, kDw2Threads, 0, st>>>((const float *)x, weight, bias, (float *)y, Bn, D, H, W, RB, ipc);
        else
            dwconv3x3_silu_fwd2_kernel<__nv_bfloat16><<<g2, kDw2Threads, 0, st>>>((const __nv_bfloat16 *)x, weight, bias,
                                                                                (__nv_bfloat16 *)y, Bn, D, H, W, RB, ipc);
        count_launch();
        TAMTR_CUDA_OK(cudaGetLastError());
        return 0;
    }
    const int tw = dw_tile_width(W);
    const dim3 grid(Bn * D, (W + tw - 1) / tw, (H + kDwTH - 1) / kDwTH);
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
    if (dw2_ok(grad_y, x, grad_x, D, W)) {
        int RB, ipc;
        dim3 g2;
        dw2_geometry(Bn, D, H, W, RB, ipc, g2);
        if (dtype == TAMTR_F32)
            dwconv3x3_silu_bwd2_kernel<float><<<g2, kDw2Threads, 0, st>>>((const float *)grad_y, (const float *)x, weight, bias,
                                                                        (float *)grad_x, grad_weight, grad_bias, Bn, D, H, W, RB, ipc);
        else
            dwconv3x3_silu_bwd2_kernel<__nv_bfloat16><<<g2, kDw2Threads, 0, st>>>(
                (const __nv_bfloat16 *)grad_y, (const __nv_bfloat16 *)x, weight, bias, (__nv_bfloat16 *)grad_x, grad_weight,
                grad_bias, Bn, D, H, W, RB, ipc);
        count_launch();
        TAMTR_CUDA_OK(cudaGetLastError());
        return 0;
    }
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
