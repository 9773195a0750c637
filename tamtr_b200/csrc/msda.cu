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
#include <stdlib.h>

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
// A lane owns VEC consecutive channels of one corner (one 16-byte load).  Blackwell-specific arithmetic:
//   * fma.rn.f32x2 / mul.rn.f32x2 (FFMA2 / FMUL2): two IEEE fp32 operations per issue slot -- the accumulate of the
//     forward and the w*grad_out of the backward run on channel PAIRS (bit-identical to scalar FFMA / FMUL);
//   * fma.rn.f32.bf16 (FHFMA.BF16): bf16 x bf16 products are exact in fp32, so the backward's <grad_out, v> dots read the
//     16-byte packs as they arrive -- no bf16 -> fp32 unpack at all.
template <typename T> struct Vec;
template <> struct Vec<float> {
    static constexpr int N = 4;
    __device__ static __forceinline__ void unpack(const uint4 &u, float (&f)[4]) {
        f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y);
        f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
    }
    // acc[j] += w * v[2j, 2j+1]
    __device__ static __forceinline__ void fma_acc(float2 (&acc)[2], const uint4 &u, const float2 ww) {
        acc[0] = __ffma2_rn(make_float2(__uint_as_float(u.x), __uint_as_float(u.y)), ww, acc[0]);
        acc[1] = __ffma2_rn(make_float2(__uint_as_float(u.z), __uint_as_float(u.w)), ww, acc[1]);
    }
    // <g, v> over this lane's channels (g: raw pack of grad_out, gf: the same unpacked)
    __device__ static __forceinline__ float dot(const uint4 &, const float (&gf)[4], const uint4 &v) {
        float d = __uint_as_float(v.x) * gf[0];
        d = fmaf(__uint_as_float(v.y), gf[1], d);
        d = fmaf(__uint_as_float(v.z), gf[2], d);
        return fmaf(__uint_as_float(v.w), gf[3], d);
    }
    __device__ static __forceinline__ void store1(float *p, float a) { *p = a; }
    __device__ static __forceinline__ void store2(float *p, float a, float b) {
        *reinterpret_cast<float2 *>(p) = make_float2(a, b);
    }
    __device__ static __forceinline__ void store4(float *p, float a, float b, float c, float d) {
        *reinterpret_cast<float4 *>(p) = make_float4(a, b, c, d);
    }
};
template <> struct Vec<__nv_bfloat16> {
    static constexpr int N = 8;
    __device__ static __forceinline__ float lo(uint32_t u) { return __uint_as_float(u << 16); }         // bf16 -> fp32 is
    __device__ static __forceinline__ float hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }  // a 16-bit shift
    __device__ static __forceinline__ void unpack(const uint4 &u, float (&f)[8]) {
        f[0] = lo(u.x); f[1] = hi(u.x); f[2] = lo(u.y); f[3] = hi(u.y);
        f[4] = lo(u.z); f[5] = hi(u.z); f[6] = lo(u.w); f[7] = hi(u.w);
    }
    __device__ static __forceinline__ void fma_acc(float2 (&acc)[4], const uint4 &u, const float2 ww) {
        acc[0] = __ffma2_rn(make_float2(lo(u.x), hi(u.x)), ww, acc[0]);
        acc[1] = __ffma2_rn(make_float2(lo(u.y), hi(u.y)), ww, acc[1]);
        acc[2] = __ffma2_rn(make_float2(lo(u.z), hi(u.z)), ww, acc[2]);
        acc[3] = __ffma2_rn(make_float2(lo(u.w), hi(u.w)), ww, acc[3]);
    }
    // d += a.lo*b.lo ; d += a.hi*b.hi with exact bf16 products and fp32 accumulation (two FHFMA.BF16, no unpack)
    __device__ static __forceinline__ float fhfma2(uint32_t a, uint32_t b, float d) {
        asm("{\n\t.reg .b16 al, ah, bl, bh;\n\t"
            "mov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\t"
            "fma.rn.f32.bf16 %0, al, bl, %0;\n\tfma.rn.f32.bf16 %0, ah, bh, %0;\n\t}"
            : "+f"(d) : "r"(a), "r"(b));
        return d;
    }
    __device__ static __forceinline__ float dot(const uint4 &g, const float (&)[8], const uint4 &v) {
        float d = fhfma2(g.x, v.x, 0.0f);
        d = fhfma2(g.y, v.y, d);
        d = fhfma2(g.z, v.z, d);
        return fhfma2(g.w, v.w, d);
    }
    __device__ static __forceinline__ uint32_t pack2(float lo_, float hi_) {
        uint32_t r;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_), "f"(lo_));
        return r;
    }
    __device__ static __forceinline__ void store1(__nv_bfloat16 *p, float a) { *p = __float2bfloat16_rn(a); }
    __device__ static __forceinline__ void store2(__nv_bfloat16 *p, float a, float b) {
        *reinterpret_cast<uint32_t *>(p) = pack2(a, b);
    }
    __device__ static __forceinline__ void store4(__nv_bfloat16 *p, float a, float b, float c, float d) {
        *reinterpret_cast<uint2 *>(p) = make_uint2(pack2(a, b), pack2(c, d));
    }
};

// grad_value[corner] += w * grad_out as 16-byte vector reductions.  GT = element type of grad_value: the value dtype, or
// fp32 for bf16 values (an fp32 gradient arena: every add then rounds at 2^-24 instead of 2^-9 of the running sum).
template <typename GT, int VEC> struct Red;
template <> struct Red<float, 4> {
    __device__ static __forceinline__ void add(float *p, const float2 (&g)[2], const float2 ww) {
        const float2 a = __fmul2_rn(g[0], ww), b = __fmul2_rn(g[1], ww);
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y)
                     : "memory");
    }
};
template <> struct Red<float, 8> {
    __device__ static __forceinline__ void add(float *p, const float2 (&g)[4], const float2 ww) {
        const float2 a = __fmul2_rn(g[0], ww), b = __fmul2_rn(g[1], ww);
        const float2 c = __fmul2_rn(g[2], ww), d = __fmul2_rn(g[3], ww);
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y)
                     : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p + 4), "f"(c.x), "f"(c.y), "f"(d.x), "f"(d.y)
                     : "memory");
    }
};
template <> struct Red<__nv_bfloat16, 8> {
    __device__ static __forceinline__ void add(__nv_bfloat16 *p, const float2 (&g)[4], const float2 ww) {
        const float2 a = __fmul2_rn(g[0], ww), b = __fmul2_rn(g[1], ww);
        const float2 c = __fmul2_rn(g[2], ww), d = __fmul2_rn(g[3], ww);
        using V = Vec<__nv_bfloat16>;
        asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(V::pack2(a.x, a.y)),
                     "r"(V::pack2(b.x, b.y)), "r"(V::pack2(c.x, c.y)), "r"(V::pack2(d.x, d.y))
                     : "memory");
    }
};

constexpr int kWarpsPerCta = 8;
constexpr int kMaxTaps = 4 * kMaxSamples;

// ------------------------------------------------------------------------------------------- item walk
// Persistent warps: warp w of the grid walks items qh = w, w + stride, ...  (qh = (b*Lq + q)*H + h).  The image index b
// and the head h are carried along instead of two integer divisions per item.
struct Walk {
    int qh, rem, h;             // rem = qh - b*Lq*H
    long img;                   // b * slab: offset of image b's value slab (in the caller's unit)
    int stride, stride_h;       // stride % H
    __device__ __forceinline__ void init(int first, int stride_, int LqH, int H, long slab) {
        qh = first;
        stride = stride_;
        const int b = first / LqH;
        rem = first - b * LqH;
        img = b * slab;
        h = first % H;
        stride_h = stride_ % H;
    }
    __device__ __forceinline__ void next(int LqH, int H, long slab) {
        qh += stride;
        rem += stride;
        if (rem >= LqH) {       // warp-uniform; a division only when the walk crosses an image boundary
            const int k = rem / LqH;
            img += k * slab;
            rem -= k * LqH;
        }
        h += stride_h;
        if (h >= H) h -= H;
    }
};

// ------------------------------------------------------------------------------------------- phase 1: taps
// Lane (s, ky) turns ROW ky (0 = north, 1 = south) of sample s into two (corner offset, A*w_k) taps and stores them as
// two 16-byte smem entries {offset | -1, -, w, w} (the weight twice: the FFMA2 / FMUL2 operand pair).  With 2*S <= 32
// the two rows of a sample sit on lanes s and S + s and one pass stages everything; otherwise lane s does both rows.
// The level geometry of a lane's sample never changes, so it is read from the constant bank once per kernel.
struct StageLane {
    int s, ky, nrows;           // nrows = 0: this lane stages nothing (it computes on sample 0 and stores nothing)
    int Hl, Wl, base;           // base = first token of the level * token stride
    int4 *slot;                 // where this lane's two taps go
    __device__ __forceinline__ void init(const Levels &lv, int S, int tok_stride, int lane, int4 *taps) {
        const bool split = 2 * S <= 32;
        nrows = (lane < (split ? 2 * S : S)) ? (split ? 1 : 2) : 0;
        s = nrows ? ((split && lane >= S) ? lane - S : lane) : 0;
        ky = (nrows && split && lane >= S) ? 1 : 0;
        const int l = lv.level_of[s];
        Hl = lv.h[l];
        Wl = lv.w[l];
        base = lv.start[l] * tok_stride;
        slot = taps + 4 * s + 2 * ky;
    }
};

struct SampleIn {   // one lane's sampling location + attention weight (idle lanes hold zeros)
    float x, y, a;
};

__device__ __forceinline__ SampleIn fetch_sample(const float *__restrict__ loc, const float *__restrict__ attn,
                                                 size_t qh, int S, const StageLane &sl) {
    SampleIn in;            // every lane loads (idle lanes: sample 0): no divergent branch around the two loads
    const float2 xy = __ldg(reinterpret_cast<const float2 *>(loc) + qh * S + sl.s);
    in.x = xy.x;
    in.y = xy.y;
    in.a = __ldg(attn + qh * S + sl.s);
    return in;
}

// One row of a sample -> 2 taps.  Weights exactly as ATen forms them, (x1-ix)*(y1-iy) etc., then times A.
// Returns the sum of the two (in-bounds) tap weights; `all_in` is cleared when a corner is out of bounds.
__device__ __forceinline__ float stage_row(int4 *dst, bool store, const Tap &t, float a, int ky, int Hl, int Wl, int base,
                                           int hoff, int rowstride, bool &all_in) {
    const int y = t.y0 + ky;
    const bool vy = (unsigned)y < (unsigned)Hl;
    const bool v0 = vy & ((unsigned)t.x0 < (unsigned)Wl);
    const bool v1 = vy & ((unsigned)(t.x0 + 1) < (unsigned)Wl);
    const float wxw = __fsub_rn(__fadd_rn(t.fx, 1.0f), t.ix);
    const float wxe = __fsub_rn(t.ix, t.fx);
    const float wy = ky ? __fsub_rn(t.iy, t.fy) : __fsub_rn(__fadd_rn(t.fy, 1.0f), t.iy);
    const float w0 = v0 ? __fmul_rn(a, __fmul_rn(wxw, wy)) : 0.0f;
    const float w1 = v1 ? __fmul_rn(a, __fmul_rn(wxe, wy)) : 0.0f;
    const int o0 = (y * Wl + t.x0) * rowstride + base + hoff;
    if (store) {
        dst[0] = make_int4(v0 ? o0 : -1, 0, __float_as_int(w0), __float_as_int(w0));
        dst[1] = make_int4(v1 ? o0 + rowstride : -1, 0, __float_as_int(w1), __float_as_int(w1));
    }
    all_in = all_in & (v0 & v1 | !store);
    return store ? w0 + w1 : 0.0f;
}

// Stages all taps of the warp's item; returns (warp-uniform) whether every corner is in bounds.  `t` is the lane's
// sample geometry (valid on lanes < S: the backward's phase 3 reuses it), `wsum` the lane's share of sum(A*w_k).
__device__ __forceinline__ bool stage_taps(const SampleIn &in, const StageLane &sl, bool two_rows, int hoff,
                                           int rowstride, Tap &t, float &wsum) {
    bool all_in = true;
    t = make_tap(in.x, in.y, sl.Hl, sl.Wl);
    wsum = stage_row(sl.slot, sl.nrows != 0, t, in.a, sl.ky, sl.Hl, sl.Wl, sl.base, hoff, rowstride, all_in);
    if (two_rows)       // warp-uniform: more than 16 samples, lane s stages both rows of sample s
        wsum += stage_row(sl.slot + 2, sl.nrows != 0, t, in.a, 1, sl.Hl, sl.Wl, sl.base, hoff, rowstride, all_in);
    __syncwarp();                               // the taps are read by other lanes
    return __all_sync(0xffffffffu, all_in);
}

// 16-byte gather of one corner.  Written as volatile asm so that the loads of a chunk keep their program order in
// front of the arithmetic (nvcc otherwise interleaves each load with the FMAs of the previous one).  The predicated
// form (registers stay zero for an out-of-bounds corner) costs a zero-fill + a select per register; items whose taps
// are ALL in bounds -- nearly all of them -- take the plain form.
__device__ __forceinline__ uint4 gather16(const char *base, int off_elems, int elem_bytes) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(base + (long)off_elems * elem_bytes));
    return v;
}
__device__ __forceinline__ uint4 gather16_pred(const char *base, int off_elems, int elem_bytes) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %5, 0;\n\t"
        "@p ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];\n\t}"
        : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w)
        : "l"(base + (long)off_elems * elem_bytes), "r"(off_elems));
    return v;
}

// Sum over the lanes that differ in the bits `m_first, 2*m_first, ..., < m_end` of the lane id, of n values per lane, as a
// TRANSPOSING reduction: at every stage a lane keeps one half of its values and sends the other half, so the shuffle and
// add counts halve from stage to stage (n/2 + n/4 + ... instead of n per stage); once n is odd the remaining stages are
// plain butterflies.  Afterwards the lane holds the totals of elements [first, first + count) and lanes with
// (id & dup) != 0 hold duplicates.  The pairing of the additions is that of the butterfly, so the sums are bit-identical.
template <int N, int M_FIRST, int M_END>
__device__ __forceinline__ void transpose_reduce(float (&a)[N], int lane, int &first, int &dup) {
    first = 0;
    dup = 0;
    int n = N;      // compile-time after unrolling
#pragma unroll
    for (int m = M_FIRST; m < M_END; m <<= 1) {
        if (n % 2 == 0) {
            const int half = n / 2;
            const bool upper = (lane & m) != 0;
#pragma unroll
            for (int j = 0; j < N / 2; ++j) {
                if (j < half) {
                    const float send = upper ? a[j] : a[j + half];
                    const float keep = upper ? a[j + half] : a[j];
                    a[j] = keep + __shfl_xor_sync(0xffffffffu, send, m);
                }
            }
            first += upper ? half : 0;
            n = half;
        } else {
#pragma unroll
            for (int j = 0; j < N; ++j)
                if (j < n) a[j] += __shfl_xor_sync(0xffffffffu, a[j], m);
            dup |= m;
        }
    }
}
template <int N, int M_FIRST, int M_END> constexpr int transpose_reduce_count() {
    int n = N;
    for (int m = M_FIRST; m < M_END; m <<= 1)
        if (n % 2 == 0) n /= 2;
    return n;
}

// ------------------------------------------------------------------------------------------- forward
// T: element type of value/out.  LPC: lanes per corner (Dh = LPC * 16/sizeof(T)).  NS: L*P when known at compile
// time (12 for the 3-level, 4-point TAM-TR / RT-DETR heads), 0 = runtime.
template <typename T, int LPC, int NS, int MAXCHUNK = 12, int MINB = 3>
__global__ void __launch_bounds__(kWarpsPerCta * 32, MINB)
msda_fwd_kernel(const T *__restrict__ value, const float *__restrict__ loc, const float *__restrict__ attn,
                T *__restrict__ out, const Levels lv, int Lq, int H, int Lv, int total, int tok_stride) {
    constexpr int VEC = Vec<T>::N;
    constexpr int DH = LPC * VEC;
    constexpr int CPL = 32 / LPC;  // corners gathered per warp-wide load
    __shared__ int4 s_taps[kWarpsPerCta][kMaxTaps];

    const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int first = blockIdx.x * kWarpsPerCta + wic;
    if (first >= total) return;  // warp-uniform; no block-wide barrier below
    const int S = NS > 0 ? NS : lv.S;
    const int LqH = Lq * H;
    const int cs = lane / LPC, cg = lane % LPC;
    const int npairs = 4 * S;
    int4 *taps = s_taps[wic];
    constexpr int ITERS = NS > 0 ? (4 * NS + CPL - 1) / CPL : 0;
    constexpr int CHUNK = NS > 0 ? (ITERS < MAXCHUNK ? ITERS : MAXCHUNK) : 4;
    constexpr bool EXACT = NS > 0 && (4 * NS) % (CPL * CHUNK) == 0;   // no partial chunk: plain loads possible
    const long slab = (long)Lv * tok_stride * (long)sizeof(T);        // bytes per image
    const bool two_rows = 2 * S > 32;
    Walk it;
    it.init(first, gridDim.x * kWarpsPerCta, LqH, H, slab);
    StageLane sl;
    sl.init(lv, S, tok_stride, lane, taps);
    // this lane's 16 bytes of token 0 of image 0, as ONE opaque 64-bit register value: every gather address is then a
    // single IMAD.WIDE (offset * sizeof(T) + base) instead of a 32-bit partial sum + a 64-bit add of the uniform pointer
    const char *lane_base = reinterpret_cast<const char *>(value) + cg * 16;
    asm volatile("" : "+l"(lane_base));

    SampleIn cur = fetch_sample(loc, attn, (size_t)it.qh, S, sl);
    while (true) {
        // software pipeline: the next item's locations/weights are requested before this item's gather is issued,
        // so a warp pays ONE exposed DRAM round trip per item instead of two dependent ones
        const bool more = it.qh + it.stride < total;
        const SampleIn nxt = fetch_sample(loc, attn, (size_t)(more ? it.qh + it.stride : it.qh), S, sl);
        Tap t;
        float wsum;
        const bool all_in = stage_taps(cur, sl, two_rows, it.h * DH, tok_stride, t, wsum);
        const char *vbase = lane_base + it.img;
        asm volatile("" : "+l"(vbase));
        float2 acc[VEC / 2];
#pragma unroll
        for (int c = 0; c < VEC / 2; ++c) acc[c] = make_float2(0.0f, 0.0f);
        // the loads of a chunk are all issued before the first use; offsets and weights are re-read from shared memory
        // where they are needed instead of staying live across the gather (register budget: 3 CTAs per SM)
        for (int base = 0; base < npairs; base += CPL * CHUNK) {
            uint4 v[CHUNK];
            const int4 *tp = taps + base + cs;
            if (EXACT && all_in) {
#pragma unroll
                for (int i = 0; i < CHUNK; ++i) v[i] = gather16(vbase, tp[i * CPL].x, (int)sizeof(T));
            } else {
#pragma unroll
                for (int i = 0; i < CHUNK; ++i) {
                    const bool live = EXACT || base + i * CPL + cs < npairs;
                    v[i] = gather16_pred(vbase, live ? tp[i * CPL].x : -1, (int)sizeof(T));
                }
            }
#pragma unroll
            for (int i = 0; i < CHUNK; ++i) {
                // past the last tap (partial chunk): v == 0, any finite weight will do
                const bool live = EXACT || base + i * CPL + cs < npairs;
                const float2 w = live ? *reinterpret_cast<const float2 *>(&tp[i * CPL].z) : make_float2(0.f, 0.f);
                Vec<T>::fma_acc(acc, v[i], w);
            }
        }
        // sum over the CPL corner slots; the lane ends up with KEEP of its VEC channels
        float a[VEC];
#pragma unroll
        for (int c = 0; c < VEC / 2; ++c) { a[2 * c] = acc[c].x; a[2 * c + 1] = acc[c].y; }
        int ch0, dup;
        transpose_reduce<VEC, LPC, 32>(a, lane, ch0, dup);
        constexpr int KEEP = transpose_reduce_count<VEC, LPC, 32>();
        if ((lane & dup) == 0) {
            T *o = out + (size_t)it.qh * DH + cg * VEC + ch0;
            if (KEEP == 1) Vec<T>::store1(o, a[0]);
            else if (KEEP == 2) Vec<T>::store2(o, a[0], a[1]);
            else Vec<T>::store4(o, a[0], a[1], a[2], a[3]);
        }
        if (!more) break;
        __syncwarp();   // taps are rewritten by the next item
        cur = nxt;
        it.next(LqH, H, slab);
    }
}

// ------------------------------------------------------------------------------------------- backward
// Same mapping.  Per tap: d_k = <grad_out, v_k> (partial per lane; the LPC lanes of the corner are summed for all taps of
// a chunk at once by a transposing reduction) and grad_value[corner] += (A*w_k) * grad_out as 16-byte vector reductions
// (REDG.F32x4 / REDG.BF16x8).  Phase 3: lanes < L*P turn the four dots of their sample into grad_attn and grad_loc.
// GT: element type of grad_value (T, or float for an fp32 gradient buffer under bf16 values).
template <typename T, typename GT, int LPC, int NS, int MAXCHUNK = 12, int MINB = 2>
__global__ void __launch_bounds__(kWarpsPerCta * 32, MINB)
msda_bwd_kernel(const T *__restrict__ grad_out, const T *__restrict__ value, const float *__restrict__ loc,
                const float *__restrict__ attn, GT *__restrict__ grad_value, float *__restrict__ grad_loc,
                float *__restrict__ grad_attn, float *__restrict__ tap_weight_sum, const Levels lv, int Lq, int H,
                int Lv, int total, int tok_stride) {
    constexpr int VEC = Vec<T>::N;
    constexpr int DH = LPC * VEC;
    constexpr int CPL = 32 / LPC;
    __shared__ int4 s_taps[kWarpsPerCta][kMaxTaps];
    __shared__ float s_dots[kWarpsPerCta][kMaxTaps];

    const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int first = blockIdx.x * kWarpsPerCta + wic;
    if (first >= total) return;
    const int S = NS > 0 ? NS : lv.S;
    const int LqH = Lq * H;
    int4 *taps = s_taps[wic];
    float *dots = s_dots[wic];
    const int cs = lane / LPC, cg = lane % LPC;
    const int npairs = 4 * S;
    constexpr int ITERS = NS > 0 ? (4 * NS + CPL - 1) / CPL : 0;
    constexpr int CHUNK = NS > 0 ? (ITERS < MAXCHUNK ? ITERS : MAXCHUNK) : 4;
    constexpr bool EXACT = NS > 0 && (4 * NS) % (CPL * CHUNK) == 0;
    const long slab = (long)Lv * tok_stride;                                     // elements per image
    const bool two_rows = 2 * S > 32;
    Walk it;
    it.init(first, gridDim.x * kWarpsPerCta, LqH, H, slab);
    StageLane sl;
    sl.init(lv, S, tok_stride, lane, taps);
    const char *lane_base = reinterpret_cast<const char *>(value) + cg * 16;     // opaque 64-bit bases: see the forward
    GT *lane_gbase = grad_value + cg * VEC;
    asm volatile("" : "+l"(lane_base), "+l"(lane_gbase));

    SampleIn cur = fetch_sample(loc, attn, (size_t)it.qh, S, sl);
    uint4 graw = __ldg(reinterpret_cast<const uint4 *>(grad_out + (size_t)it.qh * DH + cg * VEC));
    while (true) {
        // software pipeline: next item's locations / weights / grad_out row are in flight during this item's gather
        const bool more = it.qh + it.stride < total;
        const size_t qn = (size_t)(more ? it.qh + it.stride : it.qh);
        const SampleIn nxt = fetch_sample(loc, attn, qn, S, sl);
        const uint4 gnext = __ldg(reinterpret_cast<const uint4 *>(grad_out + qn * DH + cg * VEC));
        Tap t;
        float wsum;
        const bool all_in = stage_taps(cur, sl, two_rows, it.h * DH, tok_stride, t, wsum);
        const char *vbase = lane_base + it.img * (long)sizeof(T);
        GT *gvbase = lane_gbase + it.img;                           // same token stride (in elements) as value
        asm volatile("" : "+l"(vbase), "+l"(gvbase));
        float gf[VEC];
        Vec<T>::unpack(graw, gf);
        float2 g2[VEC / 2];
#pragma unroll
        for (int c = 0; c < VEC / 2; ++c) g2[c] = make_float2(gf[2 * c], gf[2 * c + 1]);
        for (int base = 0; base < npairs; base += CPL * CHUNK) {
            uint4 v[CHUNK];
            float d[CHUNK];
            const int4 *tp = taps + base + cs;
            if (EXACT && all_in) {
#pragma unroll
                for (int i = 0; i < CHUNK; ++i) v[i] = gather16(vbase, tp[i * CPL].x, (int)sizeof(T));
#pragma unroll
                for (int i = 0; i < CHUNK; ++i) {
                    const int4 e = tp[i * CPL];
                    Red<GT, VEC>::add(gvbase + e.x, g2, make_float2(__int_as_float(e.z), __int_as_float(e.w)));
                    d[i] = Vec<T>::dot(graw, gf, v[i]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < CHUNK; ++i) {
                    const bool live = EXACT || base + i * CPL + cs < npairs;
                    v[i] = gather16_pred(vbase, live ? tp[i * CPL].x : -1, (int)sizeof(T));
                }
#pragma unroll
                for (int i = 0; i < CHUNK; ++i) {
                    const bool live = EXACT || base + i * CPL + cs < npairs;
                    const int4 e = live ? tp[i * CPL] : make_int4(-1, 0, 0, 0);
                    if (e.x >= 0)
                        Red<GT, VEC>::add(gvbase + e.x, g2, make_float2(__int_as_float(e.z), __int_as_float(e.w)));
                    d[i] = Vec<T>::dot(graw, gf, v[i]);     // out-of-bounds corners: v == 0 -> d == 0
                }
            }
            int i0, dup;
            transpose_reduce<CHUNK, 1, LPC>(d, lane, i0, dup);
            constexpr int KEEP = transpose_reduce_count<CHUNK, 1, LPC>();
            if ((lane & dup) == 0) {
#pragma unroll
                for (int j = 0; j < KEEP; ++j) {
                    const int pair = base + (i0 + j) * CPL + cs;
                    if (EXACT || pair < npairs) dots[pair] = d[j];
                }
            }
        }
        __syncwarp();

        if (lane < S) {      // lane s staged (the north row of) sample s: its geometry is still in registers
            const size_t si = (size_t)it.qh * S + lane;
            const float4 d = *reinterpret_cast<const float4 *>(dots + 4 * lane);
            const float tx = t.ix - t.fx, ty = t.iy - t.fy;
            const float ux = (t.fx + 1.0f) - t.ix, uy = (t.fy + 1.0f) - t.iy;
            // grid_sampler_2d_backward: gix = -nw*(y1-iy) + ne*(y1-iy) - sw*(iy-y0) + se*(iy-y0), giy analogous
            const float ga = (ux * uy) * d.x + (tx * uy) * d.y + (ux * ty) * d.z + (tx * ty) * d.w;
            const float gix = cur.a * (uy * (d.y - d.x) + ty * (d.w - d.z));
            const float giy = cur.a * (ux * (d.z - d.x) + tx * (d.w - d.y));
            // d ix / d loc_x = (W_l / 2) * 2   (GridSampler.h:51 times d(2*loc-1)/d loc)
            reinterpret_cast<float2 *>(grad_loc)[si] = make_float2(gix * (float)sl.Wl, giy * (float)sl.Hl);
            grad_attn[si] = ga;
        }
        if (tap_weight_sum) {
            // sum of the in-bounds tap weights A*w_k of this (query, head): the column sums of grad_value -- i.e. the
            // value_proj bias gradient -- follow from it as sum_q tap_weight_sum[q,h] * grad_out[q,h,:] without ever
            // reading the dense [B, Lv, d] gradient back (ops.py::_ValueProjFn)
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, m);
            if (lane == 0) tap_weight_sum[it.qh] = wsum;
        }
        if (!more) break;
        __syncwarp();   // taps / dots are rewritten by the next item
        cur = nxt;
        graw = gnext;
        it.next(LqH, H, slab);
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
// TAMTR_MSDA_VARIANT (environment, read once): selects an alternative (loads in flight, CTAs per SM) instantiation of
// the sampler kernels for tuning runs; unset / 0 = the default.
static int tuning_variant() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("TAMTR_MSDA_VARIANT");
        v = e ? atoi(e) : 0;
    }
    return v;
}

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
    const int variant = tuning_variant();
    if (variant == 1 && lv.S == 12) {       // tuning experiment: 6 loads in flight per lane, 4 CTAs per SM
        KernelTimer timer(K_MSDA_FWD, st);
        msda_fwd_kernel<T, LPC, 12, 6, 4><<<persistent_grid(total, 4), kWarpsPerCta * 32, 0, st>>>(
            (const T *)value, loc, attn, (T *)out, lv, Lq, H, Lv, (int)total, tok_stride);
        count_launch();
        TAMTR_CUDA_OK(cudaGetLastError());
        return 0;
    }
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

template <typename T, typename GT, int LPC>
static int launch_bwd(const void *grad_out, const void *value, const float *loc, const float *attn, void *grad_value,
                      float *grad_loc, float *grad_attn, float *tap_weight_sum, const Levels &lv, int B, int Lq, int H,
                      int Lv, int tok_stride, int zero_grad_value, cudaStream_t st) {
    const long total = (long)B * Lq * H;
    const int grid = persistent_grid(total, 2);
    if (zero_grad_value) {
        TAMTR_CUDA_OK(cudaMemsetAsync(grad_value, 0, (size_t)B * Lv * tok_stride * sizeof(GT), st));
        count_launch();
    }
    KernelTimer timer(K_MSDA_BWD, st);
    if (tuning_variant() == 1 && lv.S == 12)      // tuning experiment: 6 loads in flight per lane, 3 CTAs per SM
        msda_bwd_kernel<T, GT, LPC, 12, 6, 3><<<persistent_grid(total, 3), kWarpsPerCta * 32, 0, st>>>(
            (const T *)grad_out, (const T *)value, loc, attn, (GT *)grad_value, grad_loc, grad_attn, tap_weight_sum, lv,
            Lq, H, Lv, (int)total, tok_stride);
    else if (lv.S == 12)
        msda_bwd_kernel<T, GT, LPC, 12><<<grid, kWarpsPerCta * 32, 0, st>>>(
            (const T *)grad_out, (const T *)value, loc, attn, (GT *)grad_value, grad_loc, grad_attn, tap_weight_sum, lv,
            Lq, H, Lv, (int)total, tok_stride);
    else
        msda_bwd_kernel<T, GT, LPC, 0><<<grid, kWarpsPerCta * 32, 0, st>>>(
            (const T *)grad_out, (const T *)value, loc, attn, (GT *)grad_value, grad_loc, grad_attn, tap_weight_sum, lv,
            Lq, H, Lv, (int)total, tok_stride);
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
                              int value_token_stride, int zero_grad_value, float *tap_weight_sum, int grad_value_dtype,
                              void *stream) {
    TAMTR_CHECK_ARG(grad_out && value && loc && attn && grad_value && grad_loc && grad_attn, TAMTR_E_BADARG,
                    "msda_backward: null pointer");
    TAMTR_CHECK_ARG(grad_value_dtype == dtype || (dtype == TAMTR_BF16 && grad_value_dtype == TAMTR_F32),
                    TAMTR_E_UNSUPPORTED, "msda_backward: grad_value dtype %d with value dtype %d (same, or f32 for bf16)",
                    grad_value_dtype, dtype);
    Levels lv;
    int lpc = 0;
    int ts = value_token_stride;
    const int rc = check_common(dtype, B, Lv, H, Dh, Lq, L, P, level_shapes_host, ts, lv, lpc, points_host);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
#define BWD(T, GT, N)                                                                                              \
    return launch_bwd<T, GT, N>(grad_out, value, loc, attn, grad_value, grad_loc, grad_attn, tap_weight_sum, lv, B, \
                                Lq, H, Lv, ts, zero_grad_value, st)
    if (dtype == TAMTR_F32) {
        switch (lpc) {
            case 2: BWD(float, float, 2); case 4: BWD(float, float, 4);
            case 8: BWD(float, float, 8); case 16: BWD(float, float, 16);
        }
    } else if (grad_value_dtype == TAMTR_BF16) {
        switch (lpc) {
            case 2: BWD(__nv_bfloat16, __nv_bfloat16, 2); case 4: BWD(__nv_bfloat16, __nv_bfloat16, 4);
            case 8: BWD(__nv_bfloat16, __nv_bfloat16, 8); case 16: BWD(__nv_bfloat16, __nv_bfloat16, 16);
        }
    } else {
        switch (lpc) {
            case 2: BWD(__nv_bfloat16, float, 2); case 4: BWD(__nv_bfloat16, float, 4);
            case 8: BWD(__nv_bfloat16, float, 8); case 16: BWD(__nv_bfloat16, float, 16);
        }
    }
#undef BWD
    return TAMTR_E_UNSUPPORTED;
}

extern "C" int tamtr_msda_backward(const void *grad_out, const void *value, const float *loc, const float *attn,
                                   void *grad_value, float *grad_loc, float *grad_attn, int dtype, int B, int Lv, int H,
                                   int Dh, int Lq, int L, int P, const int32_t *level_shapes_host,
                                   int value_token_stride, int zero_grad_value, float *tap_weight_sum,
                                   int grad_value_dtype, void *stream) {
    return msda_backward_impl(grad_out, value, loc, attn, grad_value, grad_loc, grad_attn, dtype, B, Lv, H, Dh, Lq, L, P,
                              nullptr, level_shapes_host, value_token_stride, zero_grad_value, tap_weight_sum,
                              grad_value_dtype, stream);
}

extern "C" int tamtr_msda_backward_ragged(const void *grad_out, const void *value, const float *loc, const float *attn,
                                          void *grad_value, float *grad_loc, float *grad_attn, int dtype, int B, int Lv,
                                          int H, int Dh, int Lq, int L, const int32_t *points_host,
                                          const int32_t *level_shapes_host, int value_token_stride, int zero_grad_value,
                                          float *tap_weight_sum, int grad_value_dtype, void *stream) {
    TAMTR_CHECK_ARG(points_host, TAMTR_E_BADARG, "msda_backward_ragged: null points_host");
    return msda_backward_impl(grad_out, value, loc, attn, grad_value, grad_loc, grad_attn, dtype, B, Lv, H, Dh, Lq, L, 0,
                              points_host, level_shapes_host, value_token_stride, zero_grad_value, tap_weight_sum,
                              grad_value_dtype, stream);
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
