// The four scan orders of SS2D (ultralytics/nn/extra_modules/VManba/csms6s.py:4-47): CrossScan writes a [b, d, h, w] map
// as four sequences -- row-major, column-major and both reversed -- and CrossMerge adds four such sequences back into one
// map; each is the other's backward.  The reference builds them from flatten / transpose / flip / cat (and add) library
// calls, 5-8 passes over [b, 4, d, h*w]; here each direction of the data flow is ONE pass: a 32x32 tile of one channel
// goes through shared memory so that both the row-major and the column-major sequences are read / written as full
// 128-byte segments (the reversed ones too: 32 consecutive addresses in descending lane order).
#include "common.cuh"

namespace tamtr {

template <typename T> __device__ __forceinline__ float cs_load(const T *p);
template <> __device__ __forceinline__ float cs_load<float>(const float *p) { return __ldg(p); }
template <> __device__ __forceinline__ float cs_load<__nv_bfloat16>(const __nv_bfloat16 *p) {
    return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void cs_store(T *p, float v);
template <> __device__ __forceinline__ void cs_store<float>(float *p, float v) { *p = v; }
template <> __device__ __forceinline__ void cs_store<__nv_bfloat16>(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }

// x [BD, H, W] -> xs [b, 4, D, L] (BD = b*D rows; row r = b*D + d lives at xs[((b*4 + k)*D + d)*L])
template <typename T>
__global__ void __launch_bounds__(256) cross_scan_kernel(const T *__restrict__ x, T *__restrict__ xs, int D, int H, int W) {
    __shared__ float tile[32][33];
    const int r = blockIdx.x, b = r / D, d = r - b * D;     // rows on grid.x (no 65 535 limit)
    const int h0 = blockIdx.z * 32, w0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const size_t L = (size_t)H * W;
    const T *src = x + (size_t)r * L;
    T *o0 = xs + (((size_t)b * 4 + 0) * D + d) * L, *o1 = o0 + (size_t)D * L, *o2 = o1 + (size_t)D * L, *o3 = o2 + (size_t)D * L;
    for (int i = ty; i < 32; i += 8) {
        const int h = h0 + i, w = w0 + tx;
        if (h < H && w < W) {
            const float v = cs_load(src + (size_t)h * W + w);
            tile[i][tx] = v;
            const size_t p = (size_t)h * W + w;
            cs_store(o0 + p, v);
            cs_store(o2 + (L - 1 - p), v);
        }
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {            // column-major: position w*H + h, h fastest
        const int w = w0 + i, h = h0 + tx;
        if (h < H && w < W) {
            const float v = tile[tx][i];
            const size_t p = (size_t)w * H + h;
            cs_store(o1 + p, v);
            cs_store(o3 + (L - 1 - p), v);
        }
    }
}

// ys [b, 4, D, L] -> y [BD, H, W]: y = ys0[p] + ys2[L-1-p] + (ys1[q] + ys3[L-1-q]), p = h*W + w, q = w*H + h
// (the reference's association: (ys0 + flip(ys2)) + transpose(ys1 + flip(ys3)), csms6s.py:31-33)
template <typename T>
__global__ void __launch_bounds__(256) cross_merge_kernel(const T *__restrict__ ys, T *__restrict__ y, int D, int H, int W) {
    __shared__ float tile[32][33];
    const int r = blockIdx.x, b = r / D, d = r - b * D;     // rows on grid.x (no 65 535 limit)
    const int h0 = blockIdx.z * 32, w0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const size_t L = (size_t)H * W;
    const T *i0 = ys + (((size_t)b * 4 + 0) * D + d) * L, *i1 = i0 + (size_t)D * L, *i2 = i1 + (size_t)D * L, *i3 = i2 + (size_t)D * L;
    for (int i = ty; i < 32; i += 8) {
        const int w = w0 + i, h = h0 + tx;
        if (h < H && w < W) {
            const size_t q = (size_t)w * H + h;
            float s = cs_load(i1 + q) + cs_load(i3 + (L - 1 - q));
            if (sizeof(T) == 2) s = __bfloat162float(__float2bfloat16_rn(s));     // the reference adds in the tensor dtype
            tile[tx][i] = s;
        }
    }
    __syncthreads();
    T *dst = y + (size_t)r * L;
    for (int i = ty; i < 32; i += 8) {
        const int h = h0 + i, w = w0 + tx;
        if (h < H && w < W) {
            const size_t p = (size_t)h * W + w;
            float s = cs_load(i0 + p) + cs_load(i2 + (L - 1 - p));
            if (sizeof(T) == 2) s = __bfloat162float(__float2bfloat16_rn(s));
            cs_store(dst + p, s + tile[i][tx]);
        }
    }
}

}  // namespace tamtr

using namespace tamtr;

static int cs_check(const void *a, const void *b, int dtype, int Bn, int D, int H, int W) {
    TAMTR_CHECK_ARG(a && b, TAMTR_E_BADARG, "cross_scan/merge: null pointer");
    TAMTR_CHECK_ARG(Bn > 0 && D > 0 && H > 0 && W > 0, TAMTR_E_BADARG, "cross_scan/merge: non-positive size");
    TAMTR_CHECK_ARG(dtype == TAMTR_F32 || dtype == TAMTR_BF16, TAMTR_E_UNSUPPORTED, "cross_scan/merge: dtype %d", dtype);
    TAMTR_CHECK_ARG((long)Bn * D <= 2147483647L && (H + 31) / 32 <= 65535 && (W + 31) / 32 <= 65535, TAMTR_E_UNSUPPORTED,
                    "cross_scan/merge: too large");
    return 0;
}

extern "C" int tamtr_cross_scan(const void *x, void *xs, int dtype, int Bn, int D, int H, int W, void *stream) {
    const int rc = cs_check(x, xs, dtype, Bn, D, H, W);
    if (rc) return rc;
    const dim3 grid(Bn * D, (W + 31) / 32, (H + 31) / 32);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == TAMTR_F32) cross_scan_kernel<float><<<grid, 256, 0, st>>>((const float *)x, (float *)xs, D, H, W);
    else cross_scan_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x, (__nv_bfloat16 *)xs, D, H, W);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_cross_merge(const void *ys, void *y, int dtype, int Bn, int D, int H, int W, void *stream) {
    const int rc = cs_check(ys, y, dtype, Bn, D, H, W);
    if (rc) return rc;
    const dim3 grid(Bn * D, (W + 31) / 32, (H + 31) / 32);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == TAMTR_F32) cross_merge_kernel<float><<<grid, 256, 0, st>>>((const float *)ys, (float *)y, D, H, W);
    else cross_merge_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)ys, (__nv_bfloat16 *)y, D, H, W);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}
