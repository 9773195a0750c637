// Multi-scale deformable attention sampler, forward + backward, hand-written for sm_100a.
//
// Replaces /root/reference ultralytics/nn/modules/utils.py:42-89 (multi_scale_deformable_attn_pytorch: per-level
// F.grid_sample on a transposed NCHW view, a materialised [B*H, Dh, Lq, L*P] tensor, multiply, reduce, transpose)
// with ONE kernel per direction that reads `value` in the head-major layout it already has ([B, Lv, H, Dh],
// transformer.py:276) and never materialises the sampled tensor.
//
// Mapping: one warp per (batch, query, head).  Phase 1: the warp's lanes turn the L*P sampling locations into
// 4*L*P (corner offset, weight) "taps" -- the index math is done ONCE per tap, not once per channel lane -- and
// stage them in shared memory.  Phase 2: every lane issues 16-byte loads; a corner's Dh channels are covered by
// LPC = Dh*sizeof(T)/16 adjacent lanes, so one warp-wide load instruction gathers 32/LPC corners as fully
// coalesced 16*LPC-byte segments.  All of a warp's loads are issued before the first FMA (memory-level
// parallelism), accumulation is fp32, the cross-corner reduction is warp shuffles.  HBM-bound: see DESIGN.md.
#include "common.cuh"

namespace tamtr {

// ------------------------------------------------------------------------------------------- index math (contract)
// utils.py:58 computes g = 2*loc - 1 (two separately rounded fp32 ops); ATen then evaluates ((g+1)*size-1)/2
// (torch/include/ATen/native/GridSampler.h:34) which executes FMA-contracted.  Intrinsics pin the rounding so that
// neither nvcc's -fmad nor algebraic simplification can change a bit.
__device__ __forceinline__ float unnormalize(float loc, int size) {
    const float g = __fadd_rn(__fmul_rn(2.0f, loc), -1.0f);
    const float gp = __fadd_rn(g, 1.0f);
    return __fmul_rn(__fmaf_rn(gp, (float)size, -1.0f), 0.5f);
}

struct Tap {
    float ix, iy, fx, fy;  // unnormalised coordinate and its floor
    int x0, y0;            // north-west corner (clamped to [-2, size+1]; NaN -> -2, i.e. fully out of bounds)
};

__device__ __forceinline__ Tap make_tap(float lx, float ly, int Hl, int Wl) {
    Tap t;
    t.ix = unnormalize(lx, Wl);
    t.iy = unnormalize(ly, Hl);
    t.fx = floorf(t.ix);
    t.fy = floorf(t.iy);
    t.x0 = __float2int_rd(fminf(fmaxf(t.fx, -2.0f), (float)Wl + 1.0f));
    t.y0 = __float2int_rd(fminf(fmaxf(t.fy, -2.0f), (float)Hl + 1.0f));
    return t;
}

// corner k: 0 = nw, 1 = ne, 2 = sw, 3 = se.  Weights are formed exactly as ATen does: (x1-ix)*(y1-iy) etc.
__device__ __forceinline__ float corner_weight(const Tap &t, int k) {
    const float wx = (k & 1) ? __fsub_rn(t.ix, t.fx) : __fsub_rn(__fadd_rn(t.fx, 1.0f), t.ix);
    const float wy = (k & 2) ? __fsub_rn(t.iy, t.fy) : __fsub_rn(__fadd_rn(t.fy, 1.0f), t.iy);
    return __fmul_rn(wx, wy);
}

__device__ __forceinline__ bool corner_inb(const Tap &t, int k, int Hl, int Wl, int &x, int &y) {
    x = t.x0 + (k & 1);
    y = t.y0 + (k >> 1);
    return (x >= 0) & (x < Wl) & (y >= 0) & (y < Hl);
}

// ------------------------------------------------------------------------------------------- vector helpers
template <typename T> struct Vec;
template <> struct Vec<float> {
    static constexpr int N = 4;
    __device__ static __forceinline__ void load(const float *p, float (&f)[4]) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    }
    __device__ static __forceinline__ void unpack(const uint4 &u, float (&f)[4]) {
        f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y);
        f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
    }
    __device__ static __forceinline__ void store(float *p, const float (&f)[4]) {
        *reinterpret_cast<float4 *>(p) = make_float4(f[0], f[1], f[2], f[3]);
    }
    __device__ static __forceinline__ void red_add(float *p, const float (&f)[4]) {
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3])
                     : "memory");
    }
};
template <> struct Vec<__nv_bfloat16> {
    static constexpr int N = 8;
    __device__ static __forceinline__ void unpack(const uint4 &u, float (&f)[8]) {
        // bf16 -> fp32 is a 16-bit shift: low half <<16, high half masked
        f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
        f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
        f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
        f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
    }
    __device__ static __forceinline__ void load(const __nv_bfloat16 *p, float (&f)[8]) {
        const uint4 u = __ldg(reinterpret_cast<const uint4 *>(p));
        unpack(u, f);
    }
    __device__ static __forceinline__ uint32_t pack2(float lo, float hi) {
        uint32_t r;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
        return r;
    }
    __device__ static __forceinline__ void store(__nv_bfloat16 *p, const float (&f)[8]) {
        *reinterpret_cast<uint4 *>(p) =
            make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
    }
    __device__ static __forceinline__ void red_add(__nv_bfloat16 *p, const float (&f)[8]) {
        asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(pack2(f[0], f[1])),
                     "r"(pack2(f[2], f[3])), "r"(pack2(f[4], f[5])), "r"(pack2(f[6], f[7]))
                     : "memory");
    }
};

constexpr int kWarpsPerCta = 8;
constexpr int kMaxTaps = 4 * kMaxSamples;

// Phase 1 (shared by fwd and bwd): lane s < L*P does the index math of sample s ONCE and writes its four
// (corner offset, weight) taps to this warp's smem slice as two 16-byte stores.
// offset = element offset of the corner's Dh-vector inside image b's value slab, or -1 when out of bounds.
struct SampleIn {   // one lane's sampling location + attention weight (lanes >= L*P hold zeros)
    float x, y, a;
};

__device__ __forceinline__ SampleIn fetch_sample(const float *__restrict__ loc, const float *__restrict__ attn,
                                                 size_t qh, int S, int lane) {
    SampleIn in = {0.f, 0.f, 0.f};
    if (lane < S) {
        const float2 xy = __ldg(reinterpret_cast<const float2 *>(loc) + qh * S + lane);
        in.x = xy.x;
        in.y = xy.y;
        in.a = __ldg(attn + qh * S + lane);
    }
    return in;
}

template <int DH>
__device__ __forceinline__ void stage_taps(int2 *taps, const SampleIn in, const Levels &lv, int S, int h, int rowstride,
                                           int lane) {
    if (lane < S) {
        const int l = lv.level_of[lane];
        const float2 xy = make_float2(in.x, in.y);
        const float a = in.a;
        const int Hl = lv.h[l], Wl = lv.w[l];
        const Tap t = make_tap(xy.x, xy.y, Hl, Wl);
        const int o_nw = (lv.start[l] + t.y0 * Wl + t.x0) * rowstride + h * DH;
        const bool vx0 = (t.x0 >= 0) & (t.x0 < Wl), vx1 = (t.x0 >= -1) & (t.x0 < Wl - 1);
        const bool vy0 = (t.y0 >= 0) & (t.y0 < Hl), vy1 = (t.y0 >= -1) & (t.y0 < Hl - 1);
        int4 lo, hi;  // {off_nw, w_nw, off_ne, w_ne}, {off_sw, w_sw, off_se, w_se}
        lo.x = (vx0 & vy0) ? o_nw : -1;
        lo.y = (vx0 & vy0) ? __float_as_int(__fmul_rn(a, corner_weight(t, 0))) : 0;
        lo.z = (vx1 & vy0) ? o_nw + rowstride : -1;
        lo.w = (vx1 & vy0) ? __float_as_int(__fmul_rn(a, corner_weight(t, 1))) : 0;
        hi.x = (vx0 & vy1) ? o_nw + Wl * rowstride : -1;
        hi.y = (vx0 & vy1) ? __float_as_int(__fmul_rn(a, corner_weight(t, 2))) : 0;
        hi.z = (vx1 & vy1) ? o_nw + (Wl + 1) * rowstride : -1;
        hi.w = (vx1 & vy1) ? __float_as_int(__fmul_rn(a, corner_weight(t, 3))) : 0;
        int4 *dst = reinterpret_cast<int4 *>(taps + 4 * lane);
        dst[0] = lo;
        dst[1] = hi;
    }
    __syncwarp();
}

// 16-byte gather of one corner; predicated off (registers stay zero) for out-of-bounds corners.  Written as
// volatile asm so that the CHUNK loads of a warp are issued back to back (nvcc otherwise interleaves each load with
// the FMAs of the previous one and leaves ~2 loads in flight per lane).
__device__ __forceinline__ uint4 gather16(const void *base, int off_elems, int elem_bytes) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    const char *p = reinterpret_cast<const char *>(base) + (long)off_elems * elem_bytes;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %5, 0;\n\t"
        "@p ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];\n\t}"
        : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w)
        : "l"(p), "r"(off_elems));
    return v;
}

// ------------------------------------------------------------------------------------------- forward
// T: element type of value/out.  LPC: lanes per corner (Dh = LPC * 16/sizeof(T)).  NS: L*P when known at compile
// time (12 for the 3-level, 4-point TAM-TR / RT-DETR heads), 0 = runtime.
template <typename T, int LPC, int NS>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 3)
msda_fwd_kernel(const T *__restrict__ value, const float *__restrict__ loc, const float *__restrict__ attn,
                T *__restrict__ out, const Levels lv, int Lq, int H, int Lv, int total, int tok_stride) {
    constexpr int VEC = Vec<T>::N;
    constexpr int DH = LPC * VEC;
    constexpr int CPL = 32 / LPC;  // corners gathered per warp-wide load
    __shared__ int2 s_taps[kWarpsPerCta][kMaxTaps];

    const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int qh = blockIdx.x * kWarpsPerCta + wic;
    if (qh >= total) return;  // warp-uniform; no block-wide barrier below
    const int S = NS > 0 ? NS : lv.S;
    const int stride = gridDim.x * kWarpsPerCta;   // persistent warps: item qh, qh + stride, ...
    const int cs = lane / LPC, cg = lane % LPC;
    const int npairs = 4 * S;
    int2 *taps = s_taps[wic];
    constexpr int ITERS = NS > 0 ? (4 * NS + CPL - 1) / CPL : 0;
    constexpr int CHUNK = NS > 0 ? (ITERS < 12 ? ITERS : 12) : 4;

    SampleIn cur = fetch_sample(loc, attn, (size_t)qh, S, lane);
    for (; qh < total; qh += stride) {
        // software pipeline: the next item's locations/weights are requested before this item's gather is issued,
        // so a warp pays ONE exposed DRAM round trip per item instead of two dependent ones
        SampleIn nxt = {0.f, 0.f, 0.f};
        if (qh + stride < total) nxt = fetch_sample(loc, attn, (size_t)(qh + stride), S, lane);
        const int h = qh % H;
        const int b = qh / (Lq * H);
        stage_taps<DH>(taps, cur, lv, S, h, tok_stride, lane);
        const T *vbase = value + (size_t)b * Lv * tok_stride + cg * VEC;
        float acc[VEC];
#pragma unroll
        for (int c = 0; c < VEC; ++c) acc[c] = 0.0f;
        for (int base = 0; base < npairs; base += CPL * CHUNK) {
            uint4 v[CHUNK];
            float w[CHUNK];
#pragma unroll
            for (int i = 0; i < CHUNK; ++i) {  // issue every load of the chunk before the first use
                const int pair = base + i * CPL + cs;
                int2 t = make_int2(-1, 0);
                if (pair < npairs) t = taps[pair];
                w[i] = __int_as_float(t.y);
                v[i] = gather16(vbase, t.x, (int)sizeof(T));
            }
#pragma unroll
            for (int i = 0; i < CHUNK; ++i) {
                float f[VEC];
                Vec<T>::unpack(v[i], f);
#pragma unroll
                for (int c = 0; c < VEC; ++c) acc[c] = fmaf(w[i], f[c], acc[c]);
            }
        }
#pragma unroll
        for (int m = LPC; m < 32; m <<= 1) {
#pragma unroll
            for (int c = 0; c < VEC; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], m);
        }
        if (cs == 0) Vec<T>::store(out + (size_t)qh * DH + cg * VEC, acc);
        __syncwarp();   // taps are rewritten by the next item
        cur = nxt;
    }
}

// ------------------------------------------------------------------------------------------- backward
// Same mapping.  Per tap: d_k = <grad_out, v_k> (partial per lane, butterfly over the LPC lanes of the corner)
// and grad_value[corner] += (A*w_k) * grad_out as one 16-byte vector reduction (REDG.F32x4 / REDG.BF16x8).
// Phase 3: lanes < L*P turn the four dots of their sample into grad_attn and grad_loc.
template <typename T, int LPC, int NS>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 2)
msda_bwd_kernel(const T *__restrict__ grad_out, const T *__restrict__ value, const float *__restrict__ loc,
                const float *__restrict__ attn, T *__restrict__ grad_value, float *__restrict__ grad_loc,
                float *__restrict__ grad_attn, float *__restrict__ tap_weight_sum, const Levels lv, int Lq, int H,
                int Lv, int total, int tok_stride) {
    constexpr int VEC = Vec<T>::N;
    constexpr int DH = LPC * VEC;
    constexpr int CPL = 32 / LPC;
    __shared__ int2 s_taps[kWarpsPerCta][kMaxTaps];
    __shared__ float s_dots[kWarpsPerCta][kMaxTaps];

    const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int qh = blockIdx.x * kWarpsPerCta + wic;
    if (qh >= total) return;
    const int S = NS > 0 ? NS : lv.S;
    const int stride = gridDim.x * kWarpsPerCta;
    int2 *taps = s_taps[wic];
    float *dots = s_dots[wic];
    const int cs = lane / LPC, cg = lane % LPC;
    const int npairs = 4 * S;
    constexpr int ITERS = NS > 0 ? (4 * NS + CPL - 1) / CPL : 0;
    constexpr int CHUNK = NS > 0 ? (ITERS < 12 ? ITERS : 12) : 4;

    SampleIn cur = fetch_sample(loc, attn, (size_t)qh, S, lane);
    uint4 graw = __ldg(reinterpret_cast<const uint4 *>(grad_out + (size_t)qh * DH + cg * VEC));
    for (; qh < total; qh += stride) {
        // software pipeline: next item's locations / weights / grad_out row are in flight during this item's gather
        SampleIn nxt = {0.f, 0.f, 0.f};
        uint4 gnext = make_uint4(0u, 0u, 0u, 0u);
        if (qh + stride < total) {
            nxt = fetch_sample(loc, attn, (size_t)(qh + stride), S, lane);
            gnext = __ldg(reinterpret_cast<const uint4 *>(grad_out + (size_t)(qh + stride) * DH + cg * VEC));
        }
        const int h = qh % H;
        const int b = qh / (Lq * H);
        stage_taps<DH>(taps, cur, lv, S, h, tok_stride, lane);
        const T *vbase = value + (size_t)b * Lv * tok_stride + cg * VEC;
        T *gvbase = grad_value + (size_t)b * Lv * tok_stride + cg * VEC;   // same token stride as value
        float g[VEC];
        Vec<T>::unpack(graw, g);
        for (int base = 0; base < npairs; base += CPL * CHUNK) {
            uint4 v[CHUNK];
            int off[CHUNK];
            float w[CHUNK];
#pragma unroll
            for (int i = 0; i < CHUNK; ++i) {
                const int pair = base + i * CPL + cs;
                int2 t = make_int2(-1, 0);
                if (pair < npairs) t = taps[pair];
                off[i] = t.x;
                w[i] = __int_as_float(t.y);
                v[i] = gather16(vbase, t.x, (int)sizeof(T));
            }
#pragma unroll
            for (int i = 0; i < CHUNK; ++i) {
                if (off[i] >= 0) {
                    float gv[VEC];
#pragma unroll
                    for (int c = 0; c < VEC; ++c) gv[c] = w[i] * g[c];
                    Vec<T>::red_add(gvbase + off[i], gv);
                }
                float f[VEC];
                Vec<T>::unpack(v[i], f);
                float d = 0.0f;
#pragma unroll
                for (int c = 0; c < VEC; ++c) d = fmaf(g[c], f[c], d);
#pragma unroll
                for (int m = 1; m < LPC; m <<= 1) d += __shfl_xor_sync(0xffffffffu, d, m);
                const int pair = base + i * CPL + cs;
                if (cg == 0 && pair < npairs) dots[pair] = d;  // out-of-bounds corners: v == 0 -> d == 0
            }
        }
        __syncwarp();

        if (lane < S) {      // lane s still holds sample s (L*P <= 32)
            const int l = lv.level_of[lane];
            const int Hl = lv.h[l], Wl = lv.w[l];
            const size_t si = (size_t)qh * S + lane;
            const Tap t = make_tap(cur.x, cur.y, Hl, Wl);
            const float4 d = *reinterpret_cast<const float4 *>(dots + 4 * lane);
            const float tx = t.ix - t.fx, ty = t.iy - t.fy;
            const float ux = (t.fx + 1.0f) - t.ix, uy = (t.fy + 1.0f) - t.iy;
            // grid_sampler_2d_backward: gix = -nw*(y1-iy) + ne*(y1-iy) - sw*(iy-y0) + se*(iy-y0), giy analogous
            const float ga = (ux * uy) * d.x + (tx * uy) * d.y + (ux * ty) * d.z + (tx * ty) * d.w;
            const float gix = cur.a * (uy * (d.y - d.x) + ty * (d.w - d.z));
            const float giy = cur.a * (ux * (d.z - d.x) + tx * (d.w - d.y));
            // d ix / d loc_x = (W_l / 2) * 2   (GridSampler.h:51 times d(2*loc-1)/d loc)
            reinterpret_cast<float2 *>(grad_loc)[si] = make_float2(gix * (float)Wl, giy * (float)Hl);
            grad_attn[si] = ga;
        }
        if (tap_weight_sum) {
            // sum of the in-bounds tap weights A*w_k of this (query, head): the column sums of grad_value -- i.e. the
            // value_proj bias gradient -- follow from it as sum_q tap_weight_sum[q,h] * grad_out[q,h,:] without ever
            // reading the dense [B, Lv, d] gradient back (ops.py::_ValueProjFn)
            float ws = 0.f;
            if (lane < S) {
                const int4 lo = *reinterpret_cast<const int4 *>(taps + 4 * lane);
                const int4 hi = *reinterpret_cast<const int4 *>(taps + 4 * lane + 2);
                ws = __int_as_float(lo.y) + __int_as_float(lo.w) + __int_as_float(hi.y) + __int_as_float(hi.w);
            }
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) ws += __shfl_xor_sync(0xffffffffu, ws, m);
            if (lane == 0) tap_weight_sum[qh] = ws;
        }
        __syncwarp();   // taps / dots are rewritten by the next item
        cur = nxt;
        graw = gnext;
    }
}

// ------------------------------------------------------------------------------------------- corners (parity export)
__global__ void msda_corners_kernel(const float *__restrict__ loc, int32_t *__restrict__ x0, int32_t *__restrict__ y0,
                                    uint8_t *__restrict__ inb, const Levels lv, long n_samples) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_samples) return;
    const int l = lv.level_of[(int)(i % lv.S)];
    const int Hl = lv.h[l], Wl = lv.w[l];
    const float2 xy = reinterpret_cast<const float2 *>(loc)[i];
    const Tap t = make_tap(xy.x, xy.y, Hl, Wl);
    x0[i] = t.x0;
    y0[i] = t.y0;
    uchar4 f;
    int x, y;
    f.x = corner_inb(t, 0, Hl, Wl, x, y);
    f.y = corner_inb(t, 1, Hl, Wl, x, y);
    f.z = corner_inb(t, 2, Hl, Wl, x, y);
    f.w = corner_inb(t, 3, Hl, Wl, x, y);
    reinterpret_cast<uchar4 *>(inb)[i] = f;
}

// ------------------------------------------------------------------------------------------- dispatch
// Persistent launch: one wave of `ctas_per_sm` CTAs per SM (or fewer when there is less work); every warp then walks
// items qh, qh + grid*8, ... so that the prefetch of the next item overlaps the gather of the current one.
static int persistent_grid(long total, int ctas_per_sm) {
    const int n_sm = sm_count();   // of the CURRENT device (a cached count of device 0 mis-sizes the grid elsewhere)
    const long need = (total + kWarpsPerCta - 1) / kWarpsPerCta;
    const long wave = (long)n_sm * ctas_per_sm;
    return (int)(need < wave ? need : wave);
}

template <typename T, int LPC>
static int launch_fwd(const void *value, const float *loc, const float *attn, void *out, const Levels &lv, int B,
                      int Lq, int H, int Lv, int tok_stride, cudaStream_t st) {
    const long total = (long)B * Lq * H;
    const int grid = persistent_grid(total, 3);
    KernelTimer timer(K_MSDA_FWD, st);
    if (lv.S == 12)
        msda_fwd_kernel<T, LPC, 12><<<grid, kWarpsPerCta * 32, 0, st>>>((const T *)value, loc, attn, (T *)out, lv, Lq,
                                                                         H, Lv, (int)total, tok_stride);
    else
        msda_fwd_kernel<T, LPC, 0><<<grid, kWarpsPerCta * 32, 0, st>>>((const T *)value, loc, attn, (T *)out, lv, Lq,
                                                                        H, Lv, (int)total, tok_stride);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename T, int LPC>
static int launch_bwd(const void *grad_out, const void *value, const float *loc, const float *attn, void *grad_value,
                      float *grad_loc, float *grad_attn, float *tap_weight_sum, const Levels &lv, int B, int Lq, int H,
                      int Lv, int tok_stride, int zero_grad_value, cudaStream_t st) {
    const long total = (long)B * Lq * H;
    const int grid = persistent_grid(total, 2);
    if (zero_grad_value) {
        TAMTR_CUDA_OK(cudaMemsetAsync(grad_value, 0, (size_t)B * Lv * tok_stride * sizeof(T), st));
        count_launch();
    }
    KernelTimer timer(K_MSDA_BWD, st);
    if (lv.S == 12)
        msda_bwd_kernel<T, LPC, 12><<<grid, kWarpsPerCta * 32, 0, st>>>((const T *)grad_out, (const T *)value, loc,
                                                                         attn, (T *)grad_value, grad_loc, grad_attn,
                                                                         tap_weight_sum, lv, Lq, H, Lv, (int)total,
                                                                         tok_stride);
    else
        msda_bwd_kernel<T, LPC, 0><<<grid, kWarpsPerCta * 32, 0, st>>>((const T *)grad_out, (const T *)value, loc,
                                                                        attn, (T *)grad_value, grad_loc, grad_attn,
                                                                        tap_weight_sum, lv, Lq, H, Lv, (int)total,
                                                                        tok_stride);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

static int check_common(int dtype, int B, int Lv, int H, int Dh, int Lq, int L, int P, const int32_t *shapes,
                        int &tok_stride, Levels &lv, int &lpc, const int32_t *points = nullptr) {
    TAMTR_CHECK_ARG(dtype == TAMTR_F32 || dtype == TAMTR_BF16, TAMTR_E_UNSUPPORTED, "msda: dtype %d not supported",
                    dtype);
    TAMTR_CHECK_ARG(B > 0 && Lv > 0 && H > 0 && Dh > 0 && Lq > 0 && shapes, TAMTR_E_BADARG,
                    "msda: non-positive size or null level_shapes");
    const int bytes = Dh * (dtype == TAMTR_F32 ? 4 : 2);
    TAMTR_CHECK_ARG(bytes == 32 || bytes == 64 || bytes == 128 || bytes == 256, TAMTR_E_UNSUPPORTED,
                    "msda: head_dim %d (%d bytes) unsupported; need Dh*sizeof in {32,64,128,256}", Dh, bytes);
    lpc = bytes / 16;
    const int rc = fill_levels(lv, L, P, shapes, Lv, points);
    TAMTR_CHECK_ARG(rc == 0, rc, "msda: bad levels (L=%d P=%d, need L<=%d, samples<=%d, sum(H_l*W_l)==Lv=%d)", L, P,
                    kMaxLevels, kMaxSamples, Lv);
    if (tok_stride <= 0) tok_stride = H * Dh;
    TAMTR_CHECK_ARG(tok_stride >= H * Dh && (tok_stride * (dtype == TAMTR_F32 ? 4 : 2)) % 16 == 0, TAMTR_E_BADARG,
                    "msda: token stride %d must be >= H*Dh = %d and 16-byte aligned", tok_stride, H * Dh);
    TAMTR_CHECK_ARG((long)Lv * tok_stride < (1L << 31) && (long)B * Lq * H < (1L << 31), TAMTR_E_UNSUPPORTED,
                    "msda: per-image value slab or query count exceeds int32 indexing");
    return 0;
}

}  // namespace tamtr

using namespace tamtr;

static int msda_forward_impl(const void *value, const float *loc, const float *attn, void *out, int dtype, int B, int Lv,
                             int H, int Dh, int Lq, int L, int P, const int32_t *points_host,
                             const int32_t *level_shapes_host, int value_token_stride, void *stream) {
    TAMTR_CHECK_ARG(value && loc && attn && out, TAMTR_E_BADARG, "msda_forward: null pointer");
    Levels lv;
    int lpc = 0;
    int ts = value_token_stride;
    const int rc = check_common(dtype, B, Lv, H, Dh, Lq, L, P, level_shapes_host, ts, lv, lpc, points_host);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
#define FWD(T, N) return launch_fwd<T, N>(value, loc, attn, out, lv, B, Lq, H, Lv, ts, st)
    if (dtype == TAMTR_F32) {
        switch (lpc) { case 2: FWD(float, 2); case 4: FWD(float, 4); case 8: FWD(float, 8); case 16: FWD(float, 16); }
    } else {
        switch (lpc) {
            case 2: FWD(__nv_bfloat16, 2); case 4: FWD(__nv_bfloat16, 4);
            case 8: FWD(__nv_bfloat16, 8); case 16: FWD(__nv_bfloat16, 16);
        }
    }
#undef FWD
    return TAMTR_E_UNSUPPORTED;
}

extern "C" int tamtr_msda_forward(const void *value, const float *loc, const float *attn, void *out, int dtype, int B,
                                  int Lv, int H, int Dh, int Lq, int L, int P, const int32_t *level_shapes_host,
                                  int value_token_stride, void *stream) {
    return msda_forward_impl(value, loc, attn, out, dtype, B, Lv, H, Dh, Lq, L, P, nullptr, level_shapes_host,
                             value_token_stride, stream);
}

extern "C" int tamtr_msda_forward_ragged(const void *value, const float *loc, const float *attn, void *out, int dtype,
                                         int B, int Lv, int H, int Dh, int Lq, int L, const int32_t *points_host,
                                         const int32_t *level_shapes_host, int value_token_stride, void *stream) {
    TAMTR_CHECK_ARG(points_host, TAMTR_E_BADARG, "msda_forward_ragged: null points_host");
    return msda_forward_impl(value, loc, attn, out, dtype, B, Lv, H, Dh, Lq, L, 0, points_host, level_shapes_host,
                             value_token_stride, stream);
}

static int msda_backward_impl(const void *grad_out, const void *value, const float *loc, const float *attn,
                              void *grad_value, float *grad_loc, float *grad_attn, int dtype, int B, int Lv, int H, int Dh,
                              int Lq, int L, int P, const int32_t *points_host, const int32_t *level_shapes_host,
                              int value_token_stride, int zero_grad_value, float *tap_weight_sum, void *stream) {
    TAMTR_CHECK_ARG(grad_out && value && loc && attn && grad_value && grad_loc && grad_attn, TAMTR_E_BADARG,
                    "msda_backward: null pointer");
    Levels lv;
    int lpc = 0;
    int ts = value_token_stride;
    const int rc = check_common(dtype, B, Lv, H, Dh, Lq, L, P, level_shapes_host, ts, lv, lpc, points_host);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
#define BWD(T, N)                                                                                                  \
    return launch_bwd<T, N>(grad_out, value, loc, attn, grad_value, grad_loc, grad_attn, tap_weight_sum, lv, B, Lq, \
                            H, Lv, ts, zero_grad_value, st)
    if (dtype == TAMTR_F32) {
        switch (lpc) { case 2: BWD(float, 2); case 4: BWD(float, 4); case 8: BWD(float, 8); case 16: BWD(float, 16); }
    } else {
        switch (lpc) {
            case 2: BWD(__nv_bfloat16, 2); case 4: BWD(__nv_bfloat16, 4);
            case 8: BWD(__nv_bfloat16, 8); case 16: BWD(__nv_bfloat16, 16);
        }
    }
#undef BWD
    return TAMTR_E_UNSUPPORTED;
}

extern "C" int tamtr_msda_backward(const void *grad_out, const void *value, const float *loc, const float *attn,
                                   void *grad_value, float *grad_loc, float *grad_attn, int dtype, int B, int Lv, int H,
                                   int Dh, int Lq, int L, int P, const int32_t *level_shapes_host,
                                   int value_token_stride, int zero_grad_value, float *tap_weight_sum, void *stream) {
    return msda_backward_impl(grad_out, value, loc, attn, grad_value, grad_loc, grad_attn, dtype, B, Lv, H, Dh, Lq, L, P,
                              nullptr, level_shapes_host, value_token_stride, zero_grad_value, tap_weight_sum, stream);
}

extern "C" int tamtr_msda_backward_ragged(const void *grad_out, const void *value, const float *loc, const float *attn,
                                          void *grad_value, float *grad_loc, float *grad_attn, int dtype, int B, int Lv,
                                          int H, int Dh, int Lq, int L, const int32_t *points_host,
                                          const int32_t *level_shapes_host, int value_token_stride, int zero_grad_value,
                                          float *tap_weight_sum, void *stream) {
    TAMTR_CHECK_ARG(points_host, TAMTR_E_BADARG, "msda_backward_ragged: null points_host");
    return msda_backward_impl(grad_out, value, loc, attn, grad_value, grad_loc, grad_attn, dtype, B, Lv, H, Dh, Lq, L, 0,
                              points_host, level_shapes_host, value_token_stride, zero_grad_value, tap_weight_sum, stream);
}

static int msda_corners_impl(const float *loc, int32_t *x0, int32_t *y0, uint8_t *inb, int B, int Lq, int H, int L, int P,
                             const int32_t *points_host, const int32_t *level_shapes_host, void *stream) {
    TAMTR_CHECK_ARG(loc && x0 && y0 && inb && level_shapes_host, TAMTR_E_BADARG, "msda_corners: null pointer");
    TAMTR_CHECK_ARG(B > 0 && Lq > 0 && H > 0, TAMTR_E_BADARG, "msda_corners: non-positive size");
    Levels lv;
    const int rc = fill_levels(lv, L, P, level_shapes_host, -1, points_host);
    TAMTR_CHECK_ARG(rc == 0, rc, "msda_corners: bad levels");
    const long n = (long)B * Lq * H * lv.S;
    msda_corners_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(loc, x0, y0, inb, lv, n);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_msda_corners(const float *loc, int32_t *x0, int32_t *y0, uint8_t *inb, int B, int Lq, int H,
                                  int L, int P, const int32_t *level_shapes_host, void *stream) {
    return msda_corners_impl(loc, x0, y0, inb, B, Lq, H, L, P, nullptr, level_shapes_host, stream);
}

extern "C" int tamtr_msda_corners_ragged(const float *loc, int32_t *x0, int32_t *y0, uint8_t *inb, int B, int Lq, int H,
                                         int L, const int32_t *points_host, const int32_t *level_shapes_host,
                                         void *stream) {
    TAMTR_CHECK_ARG(points_host, TAMTR_E_BADARG, "msda_corners_ragged: null points_host");
    return msda_corners_impl(loc, x0, y0, inb, B, Lq, H, L, 0, points_host, level_shapes_host, stream);
}
