// Selective scan (S6) of VMamba's SS2D inside TAM-TR's MEH head (ultralytics/nn/modules/head.py:1092-1098,1134 ->
// nn/extra_modules/VManba/vmamba.py:962-990 -> csms6s.py:252-270, which calls the third-party CUDA extension
// `selective_scan_cuda_core` that is not part of the reference tree).  Written from the published recurrence
// (Gu & Dao, "Mamba", 2023):
//     delta_t = softplus(dt_t + bias)   (identity above 20)
//     h_t     = exp(delta_t * A) * h_{t-1} + delta_t * B_t * u_t        per channel and state, h_{-1} = 0
//     y_t     = <C_t, h_t> + D * u_t
// Shapes as the reference passes them (vmamba.py:977-990): u, dt, y [b, K*D, L]; A [K*D, N]; B, C [b, K, N, L];
// D, bias [K*D]; fp32 throughout (vmamba.py:985-986 forces fp32 into the scan).  N = 16 states.
//
// Mapping: 4 lanes per channel, each with 4 of the 16 states in registers; CTA = 32 consecutive channels of one (image,
// scan direction) = 128 threads, walking the L positions in tiles of 32.  (A first version with one thread per channel
// and 16 states each left one warp per scheduler at the head's largest level -- 128 CTAs -- and ran at 7 % of the SFU
// bound; splitting the states over lanes quadruples the resident warps.)  Tiles of u / delta / dy are staged through
// shared memory with coalesced row reads (L is the contiguous dimension; softplus is applied once per element while
// staging) and read back as conflict-free broadcasts (pitch 33); the B / C rows of the group are staged once per tile.
// The nominal bound is the SFU (16 exp per position and channel); in practice issue slots and shared-memory wavefronts
// (DESIGN.md).  The forward also writes the state every 16 positions; the backward walks these segments in reverse,
// recomputes the states of a segment from its checkpoint (sub-checkpoints every 4 positions in shared memory, the 4 positions
// in registers) and reduces dB / dC over a warp's channels with a transpose-reduction into a per-warp slot; the CTA's four
// slots are added into global memory once per (state, position).  Inference on small grids: chunk-parallel forward (MODE 1/2).
#include "common.cuh"

namespace tamtr {

constexpr int kScN = 16;          // states
constexpr int kScSg = 4;          // lanes per channel
constexpr int kScNs = kScN / kScSg;   // states per lane
constexpr int kScThreads = 128;
constexpr int kScCh = kScThreads / kScSg;   // 32 channels per CTA
constexpr int kScT = 32;          // positions per forward tile
constexpr int kScSeg = 16;        // positions per checkpoint segment = one backward tile
constexpr int kScSub = 4;         // positions recomputed into registers at a time (backward)
constexpr float kLog2e = 1.4426950408889634f;

// softplus(x) = log1p(e^x), identity above 20 (torch's threshold).  The staging passes evaluate it once per element, which
// with log1pf() was a quarter of the forward kernel's instructions (ncu: 56 warp-instructions per position against 26 in
// the scan loop itself).  e = ex2.approx(x * log2 e); for e < 1/8 the alternating series to e^6 / 6 (relative truncation
// error e^6 / 7 <= 6e-7), otherwise lg2.approx(1 + e) * ln 2 on a result >= 0.118 (relative error <= 3e-6).
__device__ __forceinline__ float softplus20(float x) {
    const float e = __expf(fminf(x, 20.0f));
    float p = fmaf(e, -0.16666667f, 0.2f);
    p = fmaf(e, p, -0.25f);
    p = fmaf(e, p, 0.33333334f);
    p = fmaf(e, p, -0.5f);
    p = fmaf(e, p, 1.0f);
    const float r = e < 0.125f ? e * p : __logf(1.0f + e);
    return x > 20.0f ? x : r;
}
// exp2 on the SFU (ex2.approx: <= 2 ulp; arguments here are <= 0, results in (0, 1])
__device__ __forceinline__ float ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ---- asynchronous staging (cp.async, 4-byte elements: the pitch-33 tiles are not 16-byte aligned): the next tile is
// in flight while the current one is scanned.  ncu on the synchronous version: 6.1 warps per issue stalled on the long
// scoreboard (global loads of the staging phase), SFU pipe 17 %, DRAM 8 % -- pure exposed memory latency.
__device__ __forceinline__ void cp_async4(float *dst, const float *src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int n = valid ? 4 : 0;                       // src-size 0: the 4 bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- tile layouts (per buffer).  The LSU pipe was the co-bottleneck of the first cp.async version (60 % busy: ten
// 4-byte shared loads per lane and position), so operands that are consumed together are stored together:
//   ud[c][t]   = {u, delta}            one 8-byte load per lane and position   (pitch 33 elements: conflict-free rows)
//   udyr[c][t] = {u, delta, dy, raw}   one 16-byte load (backward)
//   bn[t][n], cn[t][n] (pitch 20)      one 16-byte load for the lane's 4 states; the warp's 4 state groups read 64
//                                      contiguous bytes, the 8 channels of the warp share them (broadcast)
constexpr int kScBcPitch = 20;

__device__ __forceinline__ void stage_bc_async(float (*tile)[kScBcPitch], const float *__restrict__ src, size_t grp, int L,
                                               int t0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int n = warp; n < kScN; n += kScThreads / 32) {
        const int t = t0 + lane;
        cp_async4(&tile[lane][n], src + (grp * kScN + n) * (size_t)L + min(t, L - 1), t < L);
    }
}
// rows [ch0, ch0+32) x positions [t0, t0+32) of a [rows, L] array into component `comp` of a tile of NC-float elements
template <int NC>
__device__ __forceinline__ void stage_rows_async(float *tile, int comp, const float *__restrict__ src, size_t row0, int L,
                                                 int t0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < kScCh; r += kScThreads / 32) {
        const int t = t0 + lane;
        cp_async4(tile + ((size_t)r * (kScT + 1) + lane) * NC + comp, src + (row0 + r) * (size_t)L + min(t, L - 1), t < L);
    }
}

// bf16 rows: positions in pairs (4-byte cp.async; L even), tile16[32][16] words; lanes 0-15 / 16-31 take two rows at once
__device__ __forceinline__ void stage_rows16_async(uint32_t (*tile16)[kScT / 2], const __nv_bfloat16 *__restrict__ src,
                                                   size_t row0, int L, int t0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = lane >> 4, l16 = lane & 15;
    for (int r = warp * 2 + half; r < kScCh; r += 2 * (kScThreads / 32)) {
        const int t = t0 + 2 * l16;
        const unsigned d = (unsigned)__cvta_generic_to_shared(&tile16[r][l16]);
        const __nv_bfloat16 *g = src + (row0 + r) * (size_t)L + min(t, L - 2);
        const int n = t < L ? 4 : 0;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(g), "r"(n) : "memory");
    }
}
__device__ __forceinline__ float bf16_at(const uint32_t (*tile16)[kScT / 2], int r, int t) {
    const uint32_t w = tile16[r][t >> 1];
    return __uint_as_float((t & 1) ? (w & 0xffff0000u) : (w << 16));
}
template <bool BF16> struct ScIn { using type = float; };
template <> struct ScIn<true> { using type = __nv_bfloat16; };
template <bool BF16> __device__ __forceinline__ void sc_store(void *p, size_t i, float v) {
    if (BF16) reinterpret_cast<__nv_bfloat16 *>(p)[i] = __float2bfloat16_rn(v); else reinterpret_cast<float *>(p)[i] = v;
}

// MODE 0: the whole sequence in one CTA (training: with checkpoints).
// Chunk-parallel inference, for grids too small to fill the GPU (batch 1 at 1280x1280: 32 CTAs of 4 warps walking 102 400
// positions): the sequence is cut into n_chunks pieces of chunk_len positions, blockIdx.z = the piece.
//   MODE 1: state pass -- the recurrence without outputs on pieces 0 .. n_chunks-2, started from h = 0; writes the
//           piece's end state (carry_h [Bn, KD, n_chunks, 16]) and its sum of delta (carry_s [Bn, KD, n_chunks])
//   MODE 2: output pass -- h_in of piece k folded from the pieces before it, h <- exp(A * sum delta_j) * h + end_j (the
//           recurrence is linear in h, and the decay over a piece is exp(A * sum of its deltas)), then the normal loop
template <bool BF16, int MODE>
__global__ void __launch_bounds__(kScThreads)
sscan_fwd_kernel(const void *__restrict__ u_, const void *__restrict__ dt_, const float *__restrict__ A,
                 const float *__restrict__ Bm, const float *__restrict__ Cm, const float *__restrict__ Dv,
                 const float *__restrict__ bias, float *__restrict__ y, float *__restrict__ ckpt, int KD, int Dg, int L,
                 int n_seg, float *__restrict__ carry_h, float *__restrict__ carry_s, int n_chunks, int chunk_len) {
    using TIn = typename ScIn<BF16>::type;
    const TIn *__restrict__ u = reinterpret_cast<const TIn *>(u_);
    const TIn *__restrict__ dt = reinterpret_cast<const TIn *>(dt_);
    __shared__ __align__(16) float2 s_ud[BF16 ? 1 : 2][kScCh][kScT + 1];   // bf16 inputs: filled by the conversion pass only
    __shared__ uint32_t s_u16[BF16 ? 2 : 1][BF16 ? kScCh : 1][kScT / 2], s_d16[BF16 ? 2 : 1][BF16 ? kScCh : 1][kScT / 2];
    __shared__ __align__(16) float s_b[2][kScT][kScBcPitch], s_c[2][kScT][kScBcPitch];
    __shared__ float s_yp[kScThreads][kScT + 1];        // per-lane partial outputs: summed over the 4 lanes when stored
    const int b = blockIdx.y, ch0 = blockIdx.x * kScCh;
    const int c = threadIdx.x >> 2, sg = threadIdx.x & 3, ch = ch0 + c, n0 = sg * kScNs;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t row0 = (size_t)b * KD + ch0;
    const size_t grp = (size_t)b * (KD / Dg) + ch0 / Dg;
    float a2[kScNs], h[kScNs];
#pragma unroll
    for (int j = 0; j < kScNs; ++j) { a2[j] = __ldg(A + (size_t)ch * kScN + n0 + j) * kLog2e; h[j] = 0.0f; }
    const float dsk = (Dv != nullptr && sg == 0) ? __ldg(Dv + ch) : 0.0f;     // the skip term rides on lane 0's partial
    float *ck = (MODE == 0 && ckpt != nullptr) ? ckpt + ((size_t)b * KD + ch) * (size_t)n_seg * kScN + n0 : nullptr;
    const int piece = MODE == 0 ? 0 : (int)blockIdx.z;
    const int tb = MODE == 0 ? 0 : piece * chunk_len;                    // chunk_len is a multiple of the tile length
    const int te = MODE == 0 ? L : min(L, tb + chunk_len);
    const size_t carry0 = ((size_t)b * KD + ch) * (size_t)n_chunks;
    float sum_delta = 0.0f;
    if constexpr (MODE == 2) {
        for (int j = 0; j < piece; ++j) {
            const float sj = __ldg(carry_s + carry0 + j);
            const float4 e = __ldg(reinterpret_cast<const float4 *>(carry_h + (carry0 + j) * kScN + n0));
            h[0] = fmaf(ex2(sj * a2[0]), h[0], e.x);
            h[1] = fmaf(ex2(sj * a2[1]), h[1], e.y);
            h[2] = fmaf(ex2(sj * a2[2]), h[2], e.z);
            h[3] = fmaf(ex2(sj * a2[3]), h[3], e.w);
        }
    }

    auto prefetch = [&](int buf, int t0) {
        if constexpr (BF16) {
            stage_rows16_async(s_u16[buf], u, row0, L, t0);
            stage_rows16_async(s_d16[buf], dt, row0, L, t0);
        } else {
            stage_rows_async<2>(&s_ud[BF16 ? 0 : buf][0][0].x, 0, u, row0, L, t0);
            stage_rows_async<2>(&s_ud[BF16 ? 0 : buf][0][0].x, 1, dt, row0, L, t0);
        }
        stage_bc_async(s_b[buf], Bm, grp, L, t0);
        if constexpr (MODE != 1) stage_bc_async(s_c[buf], Cm, grp, L, t0);
        cp_async_commit();
    };
    prefetch(0, tb);
    int buf = 0;
    for (int t0 = tb; t0 < te; t0 += kScT, buf ^= 1) {
        if (t0 + kScT < te) { prefetch(buf ^ 1, t0 + kScT); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        // delta = softplus(dt + bias), once per element, in place (zero past L: h is then left unchanged)
        for (int r = warp; r < kScCh; r += kScThreads / 32) {
            float v;
            if constexpr (BF16) {
                v = bf16_at(s_d16[buf], r, lane);
                s_ud[BF16 ? 0 : buf][r][lane].x = t0 + lane < L ? bf16_at(s_u16[buf], r, lane) : 0.0f;
            } else {
                v = s_ud[BF16 ? 0 : buf][r][lane].y;
            }
            v += bias != nullptr ? __ldg(bias + ch0 + r) : 0.0f;
            s_ud[BF16 ? 0 : buf][r][lane].y = t0 + lane < L ? softplus20(v) : 0.0f;
        }
        __syncthreads();
        // fixed trip counts: the loads / exps of 8 positions are hoisted ahead of the only true dependency, the 4-cycle
        // FFMA chain on h
#pragma unroll
        for (int sgm = 0; sgm < kScT / kScSeg; ++sgm) {
            if (ck != nullptr && t0 + sgm * kScSeg < L)         // state BEFORE the segment's first position
                *reinterpret_cast<float4 *>(ck + (size_t)(t0 / kScSeg + sgm) * kScN) = make_float4(h[0], h[1], h[2], h[3]);
#pragma unroll 8
            for (int t = sgm * kScSeg; t < (sgm + 1) * kScSeg; ++t) {
                const float2 ud = s_ud[BF16 ? 0 : buf][c][t];
                const float4 b4 = *reinterpret_cast<const float4 *>(&s_b[buf][t][n0]);
                const float du = ud.y * ud.x;
                h[0] = fmaf(ex2(ud.y * a2[0]), h[0], du * b4.x);
                h[1] = fmaf(ex2(ud.y * a2[1]), h[1], du * b4.y);
                h[2] = fmaf(ex2(ud.y * a2[2]), h[2], du * b4.z);
                h[3] = fmaf(ex2(ud.y * a2[3]), h[3], du * b4.w);
                if constexpr (MODE == 1) {
                    sum_delta += ud.y;
                } else {
                    const float4 c4 = *reinterpret_cast<const float4 *>(&s_c[buf][t][n0]);
                    s_yp[threadIdx.x][t] = fmaf(c4.x, h[0], fmaf(c4.y, h[1], fmaf(c4.z, h[2], fmaf(c4.w, h[3], dsk * ud.x))));
                }
            }
        }
        __syncthreads();
        if constexpr (MODE != 1) {
            for (int r = warp; r < kScCh; r += kScThreads / 32)
                if (t0 + lane < L)
                    y[(row0 + r) * (size_t)L + t0 + lane] = (s_yp[4 * r][lane] + s_yp[4 * r + 1][lane]) + (s_yp[4 * r + 2][lane] + s_yp[4 * r + 3][lane]);
        }
        // (the next iteration's first __syncthreads orders these reads of s_yp before its writes)
    }
    if constexpr (MODE == 1) {
        *reinterpret_cast<float4 *>(carry_h + (carry0 + piece) * kScN + n0) = make_float4(h[0], h[1], h[2], h[3]);
        if (sg == 0) carry_s[carry0 + piece] = sum_delta;
    }
}

// ---- backward tiles: kScSeg (16) positions, so that a CTA needs ~33 KB of shared memory and <= 128 registers per thread
// and FOUR CTAs fit an SM.  (The first version -- 32-position tiles, 81 KB, 163 registers -- fitted two: the 512 CTAs of
// the head's largest level ran as two waves of 296 + 216 at IPC 1.2.)  A warp-wide staging instruction covers 32 / T rows.
template <int T>
__device__ __forceinline__ void stage_rows_f32_t(float (*tile)[T + 1], const float *__restrict__ src, size_t row0, int L,
                                                 int t0) {
    constexpr int RPW = 32 / T;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane / T, tl = lane % T;
    for (int r = warp * RPW + sub; r < kScCh; r += RPW * (kScThreads / 32)) {
        const int t = t0 + tl;
        cp_async4(&tile[r][tl], src + (row0 + r) * (size_t)L + min(t, L - 1), t < L);
    }
}
template <int T>
__device__ __forceinline__ void stage_rows_bf16_t(uint32_t (*tile)[T / 2], const __nv_bfloat16 *__restrict__ src, size_t row0,
                                                  int L, int t0) {
    constexpr int WPR = T / 2, RPW = 32 / WPR;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane / WPR, wl = lane % WPR;
    for (int r = warp * RPW + sub; r < kScCh; r += RPW * (kScThreads / 32)) {
        const int t = t0 + 2 * wl;
        const unsigned d = (unsigned)__cvta_generic_to_shared(&tile[r][wl]);
        const __nv_bfloat16 *g = src + (row0 + r) * (size_t)L + min(t, L - 2);
        const int n = t < L ? 4 : 0;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(g), "r"(n) : "memory");
    }
}
template <int T>
__device__ __forceinline__ void stage_bc_t(float (*tile)[kScBcPitch], const float *__restrict__ src, size_t grp, int L, int t0) {
    constexpr int RPW = 32 / T;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane / T, tl = lane % T;
    for (int n = warp * RPW + sub; n < kScN; n += RPW * (kScThreads / 32)) {
        const int t = t0 + tl;
        cp_async4(&tile[tl][n], src + (grp * kScN + n) * (size_t)L + min(t, L - 1), t < L);
    }
}

constexpr int kBwT = kScSeg;
template <bool BF16> struct ScBwdSmem {
    float4 udyr[kScCh][kBwT + 1];                               // {u, delta, dy, sigmoid(dt + bias)}; .z / .w become d_u / d_raw
    float dy[2][kScCh][kBwT + 1];                               // raw tiles as they arrive, double-buffered
    float b[2][kBwT][kScBcPitch], c[2][kBwT][kScBcPitch];
    float dbc[kScThreads / 32][2 * kScN][kBwT + 1];             // per-warp dB (rows 0-15) / dC (rows 16-31) of the segment
    float sub[kBwT / kScSub][kScThreads][kScNs];                // states before every 4th position of the segment
    uint32_t u16[BF16 ? 2 : 1][BF16 ? kScCh : 1][kBwT / 2], d16[BF16 ? 2 : 1][BF16 ? kScCh : 1][kBwT / 2];
    float u32[BF16 ? 1 : 2][BF16 ? 1 : kScCh][kBwT + 1], d32[BF16 ? 1 : 2][BF16 ? 1 : kScCh][kBwT + 1];
};

// One segment (16 positions) at a time, last to first: its tiles are prefetched (cp.async) while the previous one is
// processed; the states inside the segment are recomputed from the forward's checkpoint (pass 1, keeping the state before
// every 4th position), then the segment is walked backwards in groups of 4 positions held in registers (pass 2).
// Positions past L are staged as zeros and contribute nothing, so every segment is walked in full.
template <bool BF16>
__global__ void __launch_bounds__(kScThreads, 4)
sscan_bwd_kernel(const void *__restrict__ u_, const void *__restrict__ dt_, const float *__restrict__ A,
                 const float *__restrict__ Bm, const float *__restrict__ Cm, const float *__restrict__ Dv,
                 const float *__restrict__ bias, const float *__restrict__ dy, const float *__restrict__ ckpt,
                 void *__restrict__ g_u, void *__restrict__ g_dt, float *__restrict__ g_A, float *__restrict__ g_B,
                 float *__restrict__ g_C, float *__restrict__ g_D, float *__restrict__ g_bias, int KD, int Dg, int L,
                 int n_seg) {
    using TIn = typename ScIn<BF16>::type;
    constexpr int T = kBwT, RPW = 32 / T;
    const TIn *__restrict__ u = reinterpret_cast<const TIn *>(u_);
    const TIn *__restrict__ dt = reinterpret_cast<const TIn *>(dt_);
    extern __shared__ __align__(16) unsigned char sc_raw[];
    ScBwdSmem<BF16> &sm = *reinterpret_cast<ScBwdSmem<BF16> *>(sc_raw);
    const int b = blockIdx.y, ch0 = blockIdx.x * kScCh;
    const int c = threadIdx.x >> 2, sg = threadIdx.x & 3, ch = ch0 + c, n0 = sg * kScNs;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int srow = lane / T, tl = lane % T;                      // staging / store role of this lane
    const size_t row0 = (size_t)b * KD + ch0;
    const size_t grp = (size_t)b * (KD / Dg) + ch0 / Dg;
    float a1[kScNs], a2[kScNs], dh[kScNs], dA[kScNs];
#pragma unroll
    for (int j = 0; j < kScNs; ++j) {
        a1[j] = __ldg(A + (size_t)ch * kScN + n0 + j);
        a2[j] = a1[j] * kLog2e;
        dh[j] = dA[j] = 0.0f;
    }
    const float dsk = Dv != nullptr ? __ldg(Dv + ch) : 0.0f;
    float dD = 0.0f, dbias = 0.0f;
    const float *ck = ckpt + ((size_t)b * KD + ch) * (size_t)n_seg * kScN + n0;
    // after the transposition-reduction lane (v, sg) holds dB[n0 + v] (v < 4) or dC[n0 + v - 4]
    float *my_dbc = sm.dbc[warp][(((lane >> 2) & 7) < kScNs ? 0 : kScN - kScNs) + n0 + ((lane >> 2) & 7)];

    auto prefetch = [&](int buf, int t0) {
        if constexpr (BF16) {
            stage_rows_bf16_t<T>(sm.u16[buf], u, row0, L, t0);
            stage_rows_bf16_t<T>(sm.d16[buf], dt, row0, L, t0);
        } else {
            stage_rows_f32_t<T>(sm.u32[buf], u, row0, L, t0);
            stage_rows_f32_t<T>(sm.d32[buf], dt, row0, L, t0);
        }
        stage_rows_f32_t<T>(sm.dy[buf], dy, row0, L, t0);
        stage_bc_t<T>(sm.b[buf], Bm, grp, L, t0);
        stage_bc_t<T>(sm.c[buf], Cm, grp, L, t0);
        cp_async_commit();
    };
    prefetch(0, (n_seg - 1) * kScSeg);
    int buf = 0;
    for (int seg = n_seg - 1; seg >= 0; --seg, buf ^= 1) {
        const int t0 = seg * kScSeg;
        if (seg > 0) { prefetch(buf ^ 1, t0 - kScSeg); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();                                           // (also: the previous segment's outputs were stored)
        for (int r = warp * RPW + srow; r < kScCh; r += RPW * (kScThreads / 32)) {
            float4 e;
            if constexpr (BF16) {
                const uint32_t wu = sm.u16[buf][r][tl >> 1], wd = sm.d16[buf][r][tl >> 1];
                e.x = __uint_as_float((tl & 1) ? (wu & 0xffff0000u) : (wu << 16));
                e.w = __uint_as_float((tl & 1) ? (wd & 0xffff0000u) : (wd << 16));
            } else {
                e.x = sm.u32[buf][r][tl];
                e.w = sm.d32[buf][r][tl];
            }
            const bool live = t0 + tl < L;
            e.x = live ? e.x : 0.0f;
            e.z = sm.dy[buf][r][tl];                               // zero-filled past L
            e.w += bias != nullptr ? __ldg(bias + ch0 + r) : 0.0f;
            e.y = live ? softplus20(e.w) : 0.0f;
            e.w = e.w > 20.0f ? 1.0f : 1.0f / (1.0f + __expf(-e.w));      // d softplus / d raw, once per element
            sm.udyr[r][tl] = e;
        }
        __syncthreads();
        // ---- pass 1: recompute the segment forward from its checkpoint
        {
            const float4 v = *reinterpret_cast<const float4 *>(ck + (size_t)seg * kScN);
            float h[kScNs] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int t = 0; t < T - kScSub; ++t) {                 // (the last group's entry state is all pass 2 needs)
                if ((t & (kScSub - 1)) == 0)
                    *reinterpret_cast<float4 *>(sm.sub[t / kScSub][threadIdx.x]) = make_float4(h[0], h[1], h[2], h[3]);
                const float4 e = sm.udyr[c][t];
                const float4 b4 = *reinterpret_cast<const float4 *>(&sm.b[buf][t][n0]);
                const float du = e.y * e.x;
                h[0] = fmaf(ex2(e.y * a2[0]), h[0], du * b4.x);
                h[1] = fmaf(ex2(e.y * a2[1]), h[1], du * b4.y);
                h[2] = fmaf(ex2(e.y * a2[2]), h[2], du * b4.z);
                h[3] = fmaf(ex2(e.y * a2[3]), h[3], du * b4.w);
            }
            *reinterpret_cast<float4 *>(sm.sub[T / kScSub - 1][threadIdx.x]) = make_float4(h[0], h[1], h[2], h[3]);
        }
        // ---- pass 2: groups of 4 positions, last to first (each thread reads back only its own sub-checkpoints)
        for (int g0 = T - kScSub; g0 >= 0; g0 -= kScSub) {
            float hist[kScSub + 1][kScNs];                       // hist[j] = state before position g0 + j
            float an[kScSub][kScNs];                             // exp(delta*A) of the group, reused by the reverse walk
            {
                const float4 v = *reinterpret_cast<const float4 *>(sm.sub[g0 / kScSub][threadIdx.x]);
                hist[0][0] = v.x; hist[0][1] = v.y; hist[0][2] = v.z; hist[0][3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < kScSub; ++j) {
                const float4 e = sm.udyr[c][g0 + j];
                const float4 b4 = *reinterpret_cast<const float4 *>(&sm.b[buf][g0 + j][n0]);
                const float bq[kScNs] = {b4.x, b4.y, b4.z, b4.w};
                const float du = e.y * e.x;
#pragma unroll
                for (int q = 0; q < kScNs; ++q) {
                    an[j][q] = ex2(e.y * a2[q]);
                    hist[j + 1][q] = fmaf(an[j][q], hist[j][q], du * bq[q]);
                }
            }
#pragma unroll
            for (int j = kScSub - 1; j >= 0; --j) {
                const int t = g0 + j;
                // operands are re-read from shared memory (two 16-byte loads) instead of being kept in 32 registers
                const float4 e = sm.udyr[c][t];
                const float ut = e.x, dl = e.y, gy = e.z;
                const float4 b4 = *reinterpret_cast<const float4 *>(&sm.b[buf][t][n0]);
                const float4 c4 = *reinterpret_cast<const float4 *>(&sm.c[buf][t][n0]);
                const float bq[kScNs] = {b4.x, b4.y, b4.z, b4.w};
                const float cq[kScNs] = {c4.x, c4.y, c4.z, c4.w};
                float d_dl = 0.0f, d_u = 0.0f;
                float red[2 * kScNs];                            // dB then dC contributions of this lane's 4 states
                const float dlu = dl * ut;
#pragma unroll
                for (int q = 0; q < kScNs; ++q) {
                    dh[q] = fmaf(cq[q], gy, dh[q]);                            // dL/dh_t
                    red[kScNs + q] = gy * hist[j + 1][q];
                    red[q] = dh[q] * dlu;
                    const float dah = dh[q] * an[j][q] * hist[j][q];           // dh * a * h_{t-1}
                    const float dhb = dh[q] * bq[q];
                    d_dl = fmaf(dah, a1[q], fmaf(dhb, ut, d_dl));
                    dA[q] = fmaf(dah, dl, dA[q]);
                    d_u = fmaf(dhb, dl, d_u);
                    dh[q] *= an[j][q];                                         // dL/dh_{t-1}
                }
                // sums over the 16 states of the channel: across its 4 lanes
                d_dl += __shfl_xor_sync(0xffffffffu, d_dl, 1);
                d_u += __shfl_xor_sync(0xffffffffu, d_u, 1);
                d_dl += __shfl_xor_sync(0xffffffffu, d_dl, 2);
                d_u += __shfl_xor_sync(0xffffffffu, d_u, 2);
                // dB / dC: sum over the warp's 8 channels (lane bits 2..4), 8 values -> 1 per lane in 7 shuffles
#pragma unroll
                for (int s = 4; s >= 1; s >>= 1) {
                    const bool up = (lane & (s << 2)) != 0;
#pragma unroll
                    for (int i = 0; i < s; ++i) {
                        const float send = up ? red[i] : red[i + s];
                        const float got = __shfl_xor_sync(0xffffffffu, send, s << 2);
                        red[i] = (up ? red[i + s] : red[i]) + got;
                    }
                }
                // this lane now holds value index v = (lane >> 2) & 7 of state group sg: v < 4 -> dB, else dC.  Every warp
                // owns a slot per (value, position) -- a plain store; fp32 atomicAdd on shared memory is a CAS spin loop
                // (ATOMS.CAST.SPIN) and the CTA's four warps would contend on every address
                my_dbc[t] = red[0];
                {   // branch-free: all four lanes of the channel hold the same sums, lane 0 of them stores
                    // (every lane of the channel has read udyr[c][t] before the full-mask shuffles above)
                    d_u = fmaf(dsk, gy, d_u);
                    dD = fmaf(gy, ut, dD);
                    const float d_raw = d_dl * e.w;
                    dbias += d_raw;
                    if (sg == 0) *reinterpret_cast<float2 *>(&sm.udyr[c][t].z) = make_float2(d_u, d_raw);
                }
            }
        }
        __syncthreads();
        for (int r = warp * RPW + srow; r < kScCh; r += RPW * (kScThreads / 32)) {
            if (t0 + tl < L) {
                const float4 e = sm.udyr[r][tl];
                sc_store<BF16>(g_u, (row0 + r) * (size_t)L + t0 + tl, e.z);
                sc_store<BF16>(g_dt, (row0 + r) * (size_t)L + t0 + tl, e.w);
            }
        }
        for (int n = warp * RPW + srow; n < 2 * kScN; n += RPW * (kScThreads / 32)) {       // rows 0-15: dB, 16-31: dC
            if (t0 + tl < L) {
                const float v = (sm.dbc[0][n][tl] + sm.dbc[1][n][tl]) + (sm.dbc[2][n][tl] + sm.dbc[3][n][tl]);
                float *dst = n < kScN ? g_B + (grp * kScN + n) * (size_t)L : g_C + (grp * kScN + n - kScN) * (size_t)L;
                atomicAdd(dst + t0 + tl, v);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < kScNs; ++j) atomicAdd(g_A + (size_t)ch * kScN + n0 + j, dA[j]);   // over the batch
    if (sg == 0) {
        if (g_D != nullptr) atomicAdd(g_D + ch, dD);
        if (g_bias != nullptr) atomicAdd(g_bias + ch, dbias);
    }
}

}  // namespace tamtr

using namespace tamtr;

static int sscan_check(int Bn, int KD, int Dg, int N, int L) {
    TAMTR_CHECK_ARG(Bn > 0 && KD > 0 && Dg > 0 && L > 0, TAMTR_E_BADARG, "selective_scan: non-positive size");
    TAMTR_CHECK_ARG(N == kScN, TAMTR_E_UNSUPPORTED, "selective_scan: d_state = %d unsupported (16)", N);
    TAMTR_CHECK_ARG(KD % Dg == 0 && Dg % kScCh == 0, TAMTR_E_UNSUPPORTED,
                    "selective_scan: channels per direction (%d) must be a multiple of %d", Dg, kScCh);
    TAMTR_CHECK_ARG(KD / kScCh <= 2147483647 / 1, TAMTR_E_UNSUPPORTED, "selective_scan: too many channels");
    TAMTR_CHECK_ARG(Bn <= 65535, TAMTR_E_UNSUPPORTED, "selective_scan: batch too large");
    return 0;
}

extern "C" int tamtr_selective_scan_segments(int L) { return L > 0 ? (L + kScSeg - 1) / kScSeg : 0; }

static int sscan_check_in(int in_dtype, const void *u, const void *dt, int L) {
    TAMTR_CHECK_ARG(in_dtype == TAMTR_F32 || in_dtype == TAMTR_BF16, TAMTR_E_UNSUPPORTED, "selective_scan: dtype %d", in_dtype);
    if (in_dtype == TAMTR_BF16)
        TAMTR_CHECK_ARG(L % 2 == 0 && (((uintptr_t)u | (uintptr_t)dt) & 3) == 0, TAMTR_E_UNSUPPORTED,
                        "selective_scan: bf16 inputs need an even L (%d) and 4-byte aligned rows", L);
    return 0;
}

extern "C" int tamtr_selective_scan_forward(const void *u, const void *dt, int in_dtype, const float *A, const float *Bm,
                                            const float *Cm, const float *D, const float *bias, float *y, float *ckpt,
                                            int Bn, int KD, int Dg, int N, int L, void *stream) {
    TAMTR_CHECK_ARG(u && dt && A && Bm && Cm && y, TAMTR_E_BADARG, "selective_scan_forward: null pointer");
    int rc = sscan_check(Bn, KD, Dg, N, L);
    if (rc) return rc;
    rc = sscan_check_in(in_dtype, u, dt, L);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    {
        KernelTimer timer(K_SSCAN_FWD, st);
        if (in_dtype == TAMTR_BF16)
            sscan_fwd_kernel<true, 0><<<dim3(KD / kScCh, Bn), kScThreads, 0, st>>>(
                u, dt, A, Bm, Cm, D, bias, y, ckpt, KD, Dg, L, tamtr_selective_scan_segments(L), nullptr, nullptr, 1, L);
        else
            sscan_fwd_kernel<false, 0><<<dim3(KD / kScCh, Bn), kScThreads, 0, st>>>(
                u, dt, A, Bm, Cm, D, bias, y, ckpt, KD, Dg, L, tamtr_selective_scan_segments(L), nullptr, nullptr, 1, L);
    }
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

// How many pieces the chunk-parallel inference forward should cut the sequence into: 1 (= use the plain forward) when the
// (channel block, image) grid already fills the GPU, otherwise enough pieces for about four CTAs per SM, at most 32, each
// at least 1024 positions.
extern "C" int tamtr_selective_scan_chunks(int Bn, int KD, int L) {
    if (Bn <= 0 || KD < kScCh || L <= 0) return 1;
    const long ctas = (long)(KD / kScCh) * Bn;
    const long want = 4L * ::tamtr::sm_count();
    if (ctas * 2 > want) return 1;
    long k = (want + ctas - 1) / ctas;
    if (k > 32) k = 32;
    while (k > 1 && L / k < 1024) --k;
    return (int)k;
}

extern "C" int tamtr_selective_scan_forward_chunked(const void *u, const void *dt, int in_dtype, const float *A,
                                                    const float *Bm, const float *Cm, const float *D, const float *bias,
                                                    float *y, float *carry, int n_chunks, int Bn, int KD, int Dg, int N,
                                                    int L, void *stream) {
    TAMTR_CHECK_ARG(u && dt && A && Bm && Cm && y && carry, TAMTR_E_BADARG, "selective_scan_forward_chunked: null pointer");
    TAMTR_CHECK_ARG(n_chunks >= 2 && n_chunks <= 64, TAMTR_E_BADARG, "selective_scan_forward_chunked: n_chunks = %d", n_chunks);
    int rc = sscan_check(Bn, KD, Dg, N, L);
    if (rc) return rc;
    rc = sscan_check_in(in_dtype, u, dt, L);
    if (rc) return rc;
    TAMTR_CHECK_ARG(n_chunks <= 65535, TAMTR_E_UNSUPPORTED, "selective_scan_forward_chunked: too many chunks");
    cudaStream_t st = (cudaStream_t)stream;
    const int chunk_len = (((L + n_chunks - 1) / n_chunks + kScT - 1) / kScT) * kScT;
    const int pieces = (L + chunk_len - 1) / chunk_len;              // <= n_chunks; the carry arrays keep stride n_chunks
    float *carry_h = carry, *carry_s = carry + (size_t)Bn * KD * n_chunks * kScN;
    const int nseg = tamtr_selective_scan_segments(L);
    KernelTimer timer(K_SSCAN_FWD, st);
    if (pieces > 1) {
        const dim3 ga(KD / kScCh, Bn, pieces - 1);
        if (in_dtype == TAMTR_BF16)
            sscan_fwd_kernel<true, 1><<<ga, kScThreads, 0, st>>>(u, dt, A, Bm, Cm, D, bias, y, nullptr, KD, Dg, L, nseg, carry_h,
                                                               carry_s, n_chunks, chunk_len);
        else
            sscan_fwd_kernel<false, 1><<<ga, kScThreads, 0, st>>>(u, dt, A, Bm, Cm, D, bias, y, nullptr, KD, Dg, L, nseg, carry_h,
                                                                carry_s, n_chunks, chunk_len);
        count_launch();
    }
    const dim3 gc(KD / kScCh, Bn, pieces);
    if (in_dtype == TAMTR_BF16)
        sscan_fwd_kernel<true, 2><<<gc, kScThreads, 0, st>>>(u, dt, A, Bm, Cm, D, bias, y, nullptr, KD, Dg, L, nseg, carry_h,
                                                           carry_s, n_chunks, chunk_len);
    else
        sscan_fwd_kernel<false, 2><<<gc, kScThreads, 0, st>>>(u, dt, A, Bm, Cm, D, bias, y, nullptr, KD, Dg, L, nseg, carry_h,
                                                            carry_s, n_chunks, chunk_len);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_selective_scan_backward(const void *u, const void *dt, int in_dtype, const float *A, const float *Bm,
                                             const float *Cm, const float *D, const float *bias, const float *dy,
                                             const float *ckpt, void *g_u, void *g_dt, float *g_A, float *g_B,
                                             float *g_C, float *g_D, float *g_bias, int Bn, int KD, int Dg, int N, int L,
                                             void *stream) {
    TAMTR_CHECK_ARG(u && dt && A && Bm && Cm && dy && ckpt && g_u && g_dt && g_A && g_B && g_C, TAMTR_E_BADARG,
                    "selective_scan_backward: null pointer");
    int rc = sscan_check(Bn, KD, Dg, N, L);
    if (rc) return rc;
    rc = sscan_check_in(in_dtype, u, dt, L);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t grp_elems = (size_t)Bn * (KD / Dg) * kScN * L;
    TAMTR_CUDA_OK(cudaMemsetAsync(g_A, 0, (size_t)KD * kScN * sizeof(float), st));
    TAMTR_CUDA_OK(cudaMemsetAsync(g_B, 0, grp_elems * sizeof(float), st));
    TAMTR_CUDA_OK(cudaMemsetAsync(g_C, 0, grp_elems * sizeof(float), st));
    if (g_D) TAMTR_CUDA_OK(cudaMemsetAsync(g_D, 0, (size_t)KD * sizeof(float), st));
    if (g_bias) TAMTR_CUDA_OK(cudaMemsetAsync(g_bias, 0, (size_t)KD * sizeof(float), st));
    static bool attr_set[64] = {false};          // cudaFuncSetAttribute is per device
    int dev_id = 0;
    TAMTR_CUDA_OK(cudaGetDevice(&dev_id));
    if (dev_id < 0 || dev_id >= 64 || !attr_set[dev_id]) {
        TAMTR_CUDA_OK(cudaFuncSetAttribute(sscan_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(ScBwdSmem<false>)));
        TAMTR_CUDA_OK(cudaFuncSetAttribute(sscan_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(ScBwdSmem<true>)));
        if (dev_id >= 0 && dev_id < 64) attr_set[dev_id] = true;
    }
    {
        KernelTimer timer(K_SSCAN_BWD, st);
        if (in_dtype == TAMTR_BF16)
            sscan_bwd_kernel<true><<<dim3(KD / kScCh, Bn), kScThreads, sizeof(ScBwdSmem<true>), st>>>(
                u, dt, A, Bm, Cm, D, bias, dy, ckpt, g_u, g_dt, g_A, g_B, g_C, g_D, g_bias, KD, Dg, L,
                tamtr_selective_scan_segments(L));
        else
            sscan_bwd_kernel<false><<<dim3(KD / kScCh, Bn), kScThreads, sizeof(ScBwdSmem<false>), st>>>(
                u, dt, A, Bm, Cm, D, bias, dy, ckpt, g_u, g_dt, g_A, g_B, g_C, g_D, g_bias, KD, Dg, L,
                tamtr_selective_scan_segments(L));
    }
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}
