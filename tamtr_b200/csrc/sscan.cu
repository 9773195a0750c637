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
// Mapping (round 2).  A thread owns TWO neighbouring channels x FOUR states; the four lanes `sg` of a channel pair cover the
// 16 states, a warp covers 16 channels, a CTA of NW warps covers 16*NW consecutive channels of one (image, scan direction).
//   * The nominal bound is the SFU: 16 ex2 per (channel, position), 16 results / clk / SM -> 1 clk per (channel, position)
//     and SM.  The round-1 mapping (one channel x four states per thread) needed 11 shared-memory wavefronts and 26 issue
//     slots per 8 (channel, position)s and ran at 0.32 of that bound.
//   * All arithmetic is packed over the CHANNEL PAIR (FMUL2 / FFMA2: two IEEE fp32 operations per issue slot): h2[k] =
//     {h of channel 0, h of channel 1} for the lane's state k.  B_t[k] / C_t[k] are then scalar (broadcast) operands, read
//     from tiles that keep the global [state][position] layout -- a lane fetches 4 positions of one state with one 16-byte
//     load, the "transposition" is register naming -- and the outputs' sums over states need no horizontal adds.
//   * Work is done in blocks of 4 positions, written as separate loops per phase (exponentials and input terms of all 4
//     positions; then the only true recurrence, one FFMA2 per state pair and position; then outputs; then the lane sums of
//     all 4 positions) so that the in-order instruction stream interleaves independent work: with two warps per scheduler
//     (all there is at the head's largest level) a position's dependent chain load -> ex2 -> FMA -> shuffle -> shuffle ->
//     store (~180 clk) was exactly what the first version of this mapping ran at (ncu: IPC 0.42, 92 instructions per
//     warp-position, half of them staging / conversion overhead).
//   * Sums over the 16 states of a channel (y; d_u, d_delta) = over the 4 lanes sg: a transposing butterfly (each lane ends
//     up owning ONE of the pair's values and stores it).
//   * Backward, sums over channels (dB_t, dC_t): the two channels of a thread are one horizontal add; the warp's 8 channel
//     pairs are folded with a SELECT-FREE transposing butterfly: register slot k of lane (sg, cp) holds state
//     4 sg + (k ^ (cp & 3)), so at every step each lane keeps the slots whose index bit is 0 and receives its partner's
//     slots whose bit is 1 -- static register names, no predicated moves (4 shuffles + 4 adds per 4 values and 16
//     channels; the classic butterfly with selects costs 14).  The permuted state order is free: it only changes which
//     ROW of the B / C tile a slot reads.
//   * Tiles are staged with 16-byte cp.async (4-byte when L is not a multiple of 4), double-buffered: the next tile is in
//     flight while the current one is scanned.  u / dt / dy are converted ONCE per element into {u0, u1, delta0, delta1} /
//     {dy0, dy1, sig0, sig1} entries per channel pair (softplus on ex2 / lg2).  Outputs go through shared memory so that
//     global stores are row-contiguous 16-byte stores.
//   * The forward writes the state every 16 positions; the backward walks these segments last to first, recomputes a
//     segment from its checkpoint (pass 1: states before every 4th position into shared memory; pass 2: 4 positions in
//     registers, their operands kept in registers for the reverse walk).
// Inference on small grids: chunk-parallel forward (MODE 1/2).
#include <initializer_list>

#include "common.cuh"

namespace tamtr {

constexpr int kScN = 16;          // states
constexpr int kScT = 32;          // positions per forward tile
constexpr int kScSeg = 16;        // positions per checkpoint segment = one backward tile
constexpr int kScSub = 4;         // positions recomputed into registers at a time (backward)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// softplus(x) = log1p(e^x), identity above 20 (torch's threshold), and its derivative sigmoid(x).
// e = ex2.approx(x * log2 e); for e < 1/8 the alternating series to e^6 / 6 (relative truncation error e^6 / 7 <= 6e-7),
// otherwise lg2.approx(1 + e) * ln 2 on a result >= 0.118 (relative error <= 3e-6).
__device__ __forceinline__ float softplus20(float x, float *sig = nullptr) {
    const float e = __expf(fminf(x, 20.0f));
    float p = fmaf(e, -0.16666667f, 0.2f);
    p = fmaf(e, p, -0.25f);
    p = fmaf(e, p, 0.33333334f);
    p = fmaf(e, p, -0.5f);
    p = fmaf(e, p, 1.0f);
    const float r = e < 0.125f ? e * p : __logf(1.0f + e);
    if (sig != nullptr) *sig = x > 20.0f ? 1.0f : __fdividef(e, 1.0f + e);
    return x > 20.0f ? x : r;
}
// exp2 on the SFU (ex2.approx: <= 2 ulp; arguments here are <= 0, results in (0, 1])
__device__ __forceinline__ float ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float2 ex2(float2 x) { return make_float2(ex2(x.x), ex2(x.y)); }
__device__ __forceinline__ float2 dup(float x) { return make_float2(x, x); }

__device__ __forceinline__ void cp_async4(void *dst, const void *src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int n = valid ? 4 : 0;                       // src-size 0: the 4 bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <bool BF16> struct ScIn { using type = float; };
template <> struct ScIn<true> { using type = __nv_bfloat16; };
template <bool BF16> __device__ __forceinline__ void sc_store(void *p, size_t i, float v) {
    if (BF16) reinterpret_cast<__nv_bfloat16 *>(p)[i] = __float2bfloat16_rn(v); else reinterpret_cast<float *>(p)[i] = v;
}

__device__ __forceinline__ void cp_async16(void *dst, const void *src, int bytes) {   // bytes < 16: the rest is zero-filled
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ float comp(const float4 &v, int j) { return j == 0 ? v.x : j == 1 ? v.y : j == 2 ? v.z : v.w; }
__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// ---- staging.  Input rows are arrays of 32-bit words (fp32: one position per word, bf16: two), copied in 16-byte chunks
// when the row length and the base addresses allow (`vec`), else word by word; a tile row is TW words + 4 of padding, so
// that 16-byte accesses at a stride of one or two rows spread over the banks.
template <int ROWS, int TW, int NT>
__device__ __forceinline__ void stage_rows(uint32_t *tile, const uint32_t *__restrict__ src, size_t row0, int LW, int w0, bool vec) {
    constexpr int PITCH = TW + 4;
    if (vec) {
        constexpr int CPR = TW / 4;
        for (int i = threadIdx.x; i < ROWS * CPR; i += NT) {
            const int r = i / CPR, c = i % CPR, w = w0 + 4 * c;
            cp_async16(tile + r * PITCH + 4 * c, src + (row0 + r) * (size_t)LW + min(w, LW - 4), w < LW ? 16 : 0);
        }
    } else {
        for (int i = threadIdx.x; i < ROWS * TW; i += NT) {
            const int r = i / TW, wl = i % TW, w = w0 + wl;
            cp_async4(tile + r * PITCH + wl, src + (row0 + r) * (size_t)LW + min(w, LW - 1), w < LW);
        }
    }
}
// B_t / C_t of the group keep their global layout [state n][position]: tile[n][T], the 16-byte chunk c of row n stored at
// chunk position c ^ (n >> 2) -- the four lanes sg = n >> 2 of a channel pair read four different rows at the same
// position, and rows are a multiple of 64 bytes apart.
template <int T, int NT>
__device__ __forceinline__ void stage_bc(float *tile, const float *__restrict__ src, size_t grp, int L, int t0, bool vec) {
    if (vec) {
        constexpr int CPR = T / 4;
        for (int i = threadIdx.x; i < kScN * CPR; i += NT) {
            const int n = i / CPR, c = i % CPR, t = t0 + 4 * c;
            cp_async16(tile + n * T + ((c ^ (n >> 2)) << 2), src + (grp * kScN + n) * (size_t)L + min(t, L - 4), t < L ? 16 : 0);
        }
    } else {
        for (int i = threadIdx.x; i < kScN * T; i += NT) {
            const int n = i / T, tl = i % T, t = t0 + tl;
            cp_async4(tile + n * T + ((((tl >> 2) ^ (n >> 2)) << 2) | (tl & 3)), src + (grp * kScN + n) * (size_t)L + min(t, L - 1), t < L);
        }
    }
}

template <bool BF16, int NW> struct ScFwdSmem {
    static constexpr int CH = 16 * NW, TWIN = BF16 ? kScT / 2 : kScT, PIN = TWIN + 4;
    float4 ud[CH / 2][kScT + 1];                                 // {u0, u1, delta0, delta1} of a channel pair
    float b[2][kScN][kScT], c[2][kScN][kScT];
    float y[CH][kScT + 4];
    uint32_t uin[2][CH][PIN], din[2][CH][PIN];                   // input rows as they arrive, double-buffered
};

// MODE 0: the whole sequence in one CTA (training: with checkpoints).
// Chunk-parallel inference, for grids too small to fill the GPU (batch 1 at 1280x1280): the sequence is cut into n_chunks
// pieces of chunk_len positions, blockIdx.z = the piece.
//   MODE 1: state pass -- the recurrence without outputs on pieces 0 .. n_chunks-2, started from h = 0; writes the
//           piece's end state (carry_h [Bn, KD, n_chunks, 16]) and its sum of delta (carry_s [Bn, KD, n_chunks])
//   MODE 2: output pass -- h_in of piece k folded from the pieces before it, h <- exp(A * sum delta_j) * h + end_j (the
//           recurrence is linear in h, and the decay over a piece is exp(A * sum of its deltas)), then the normal loop
template <bool BF16, int MODE, int NW>
__global__ void __launch_bounds__(32 * NW)
sscan_fwd_kernel(const void *__restrict__ u_, const void *__restrict__ dt_, const float *__restrict__ A,
                 const float *__restrict__ Bm, const float *__restrict__ Cm, const float *__restrict__ Dv,
                 const float *__restrict__ bias, float *__restrict__ y, float *__restrict__ ckpt, int KD, int Dg, int L,
                 int n_seg, float *__restrict__ carry_h, float *__restrict__ carry_s, int n_chunks, int chunk_len, int vec_) {
    using SM = ScFwdSmem<BF16, NW>;
    constexpr int CH = 16 * NW, NT = 32 * NW, PAIRS = CH / 2;
    const bool vec = vec_ != 0;
    extern __shared__ __align__(16) unsigned char sc_raw[];
    SM &sm = *reinterpret_cast<SM *>(sc_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sg = lane & 3, cp = lane >> 2, pair = warp * 8 + cp;
    const int b = blockIdx.y, ch0 = blockIdx.x * CH, chA = ch0 + 2 * pair, n0 = 4 * sg;
    const size_t row0 = (size_t)b * KD + ch0;
    const size_t grp = (size_t)b * (KD / Dg) + ch0 / Dg;
    const int LW = BF16 ? L / 2 : L;
    float2 A2[4], h2[4];                                         // [state n0 + k] = {channel chA, channel chA + 1}
    {
        const float4 a0 = __ldg(reinterpret_cast<const float4 *>(A + (size_t)chA * kScN + n0));
        const float4 a1 = __ldg(reinterpret_cast<const float4 *>(A + (size_t)(chA + 1) * kScN + n0));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            A2[k] = make_float2(comp(a0, k) * kLog2e, comp(a1, k) * kLog2e);
            h2[k] = make_float2(0.0f, 0.0f);
        }
    }
    // the skip term rides on lane 0's partial
    const float2 dsk2 = (Dv != nullptr && sg == 0) ? make_float2(__ldg(Dv + chA), __ldg(Dv + chA + 1)) : make_float2(0.0f, 0.0f);
    // conversion pass: thread <-> channel pair cpair (NT is a multiple of PAIRS, so the pair is fixed)
    const int cpair = tid % PAIRS;
    const float cbias0 = bias != nullptr ? __ldg(bias + ch0 + 2 * cpair) : 0.0f;
    const float cbias1 = bias != nullptr ? __ldg(bias + ch0 + 2 * cpair + 1) : 0.0f;
    // checkpoints: [b][channel pair][segment][state][2 channels] -- a lane's 4 state pairs are 32 contiguous bytes, in the
    // order the registers hold them
    float *ck = (MODE == 0 && ckpt != nullptr) ? ckpt + ((size_t)b * KD + chA) * (size_t)n_seg * kScN + 2 * n0 : nullptr;
    const int piece = MODE == 0 ? 0 : (int)blockIdx.z;
    const int tb = MODE == 0 ? 0 : piece * chunk_len;                    // chunk_len is a multiple of the tile length
    const int te = MODE == 0 ? L : min(L, tb + chunk_len);
    const size_t carry0 = ((size_t)b * KD + chA) * (size_t)n_chunks;     // channel chA + e: + e * n_chunks
    float2 sum_delta2 = make_float2(0.0f, 0.0f);
    if constexpr (MODE == 2) {
        for (int j = 0; j < piece; ++j) {
            const float2 s2 = make_float2(__ldg(carry_s + carry0 + j), __ldg(carry_s + carry0 + n_chunks + j));
            const float4 e0 = __ldg(reinterpret_cast<const float4 *>(carry_h + (carry0 + j) * kScN + n0));
            const float4 e1 = __ldg(reinterpret_cast<const float4 *>(carry_h + (carry0 + n_chunks + j) * kScN + n0));
#pragma unroll
            for (int k = 0; k < 4; ++k)
                h2[k] = __ffma2_rn(ex2(__fmul2_rn(s2, A2[k])), h2[k], make_float2(comp(e0, k), comp(e1, k)));
        }
    }

    auto prefetch = [&](int buf, int t0) {
        const int w0 = BF16 ? t0 / 2 : t0;
        stage_rows<CH, SM::TWIN, NT>(&sm.uin[buf][0][0], reinterpret_cast<const uint32_t *>(u_), row0, LW, w0, vec);
        stage_rows<CH, SM::TWIN, NT>(&sm.din[buf][0][0], reinterpret_cast<const uint32_t *>(dt_), row0, LW, w0, vec);
        stage_bc<kScT, NT>(&sm.b[buf][0][0], Bm, grp, L, t0, vec);
        if constexpr (MODE != 1) stage_bc<kScT, NT>(&sm.c[buf][0][0], Cm, grp, L, t0, vec);
        cp_async_commit();
    };
    prefetch(0, tb);
    int buf = 0;
    for (int t0 = tb; t0 < te; t0 += kScT, buf ^= 1) {
        if (t0 + kScT < te) { prefetch(buf ^ 1, t0 + kScT); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        // ---- conversion pass: {u0, u1, delta0, delta1}, delta = softplus(dt + bias), once per element (zeros past L: h is
        // then left unchanged).  A thread converts 4 positions of a channel pair: 16-byte loads of the two rows, 16-byte
        // stores of the four entries; lanes run over pairs.
        for (int c4 = tid / PAIRS; c4 < kScT / 4; c4 += NT / PAIRS) {
            float uu[2][4], dd[2][4];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int r = 2 * cpair + e;
                if constexpr (BF16) {
                    const uint2 wu = *reinterpret_cast<const uint2 *>(&sm.uin[buf][r][2 * c4]);
                    const uint2 wd = *reinterpret_cast<const uint2 *>(&sm.din[buf][r][2 * c4]);
                    uu[e][0] = bf16lo(wu.x); uu[e][1] = bf16hi(wu.x); uu[e][2] = bf16lo(wu.y); uu[e][3] = bf16hi(wu.y);
                    dd[e][0] = bf16lo(wd.x); dd[e][1] = bf16hi(wd.x); dd[e][2] = bf16lo(wd.y); dd[e][3] = bf16hi(wd.y);
                } else {
                    const float4 fu = *reinterpret_cast<const float4 *>(&sm.uin[buf][r][4 * c4]);
                    const float4 fd = *reinterpret_cast<const float4 *>(&sm.din[buf][r][4 * c4]);
                    uu[e][0] = fu.x; uu[e][1] = fu.y; uu[e][2] = fu.z; uu[e][3] = fu.w;
                    dd[e][0] = fd.x; dd[e][1] = fd.y; dd[e][2] = fd.z; dd[e][3] = fd.w;
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool live = t0 + 4 * c4 + k < L;
                const float d0 = softplus20(dd[0][k] + cbias0), d1 = softplus20(dd[1][k] + cbias1);
                sm.ud[cpair][4 * c4 + k] = make_float4(live ? uu[0][k] : 0.0f, live ? uu[1][k] : 0.0f, live ? d0 : 0.0f, live ? d1 : 0.0f);
            }
        }
        __syncthreads();
        // ---- blocks of 4 positions, one loop per phase (file header)
#pragma unroll 1
        for (int tq = 0; tq < kScT; tq += 4) {
            if (ck != nullptr && (tq % kScSeg) == 0 && t0 + tq < L) {      // state BEFORE the segment's first position
                float *dst = ck + (size_t)((t0 + tq) / kScSeg) * (2 * kScN);
                *reinterpret_cast<float4 *>(dst) = make_float4(h2[0].x, h2[0].y, h2[1].x, h2[1].y);
                *reinterpret_cast<float4 *>(dst + 4) = make_float4(h2[2].x, h2[2].y, h2[3].x, h2[3].y);
            }
            float4 uv[4], bk[4], cq[4];
            const int chunk = ((tq >> 2) ^ sg) << 2;
#pragma unroll
            for (int j = 0; j < 4; ++j) uv[j] = sm.ud[pair][tq + j];
#pragma unroll
            for (int k = 0; k < 4; ++k) bk[k] = *reinterpret_cast<const float4 *>(&sm.b[buf][n0 + k][chunk]);
            if constexpr (MODE != 1) {
#pragma unroll
                for (int k = 0; k < 4; ++k) cq[k] = *reinterpret_cast<const float4 *>(&sm.c[buf][n0 + k][chunk]);
            }
            float2 ea[4][4], db[4][4];                                                       // [position][state]
#pragma unroll
            for (int j = 0; j < 4; ++j) {                                                    // exponents
                const float2 D2 = make_float2(uv[j].z, uv[j].w);
#pragma unroll
                for (int k = 0; k < 4; ++k) ea[j][k] = __fmul2_rn(D2, A2[k]);
                if constexpr (MODE == 1) sum_delta2 = __fadd2_rn(sum_delta2, D2);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {                                                    // the SFU work
#pragma unroll
                for (int k = 0; k < 4; ++k) ea[j][k] = ex2(ea[j][k]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {                                                    // input terms delta u B
                const float2 DU2 = __fmul2_rn(make_float2(uv[j].z, uv[j].w), make_float2(uv[j].x, uv[j].y));
#pragma unroll
                for (int k = 0; k < 4; ++k) db[j][k] = __fmul2_rn(DU2, dup(comp(bk[k], j)));
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {                                                    // the recurrence; ea <- h_t
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    h2[k] = __ffma2_rn(ea[j][k], h2[k], db[j][k]);
                    ea[j][k] = h2[k];
                }
            }
            if constexpr (MODE != 1) {
                float2 y2[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) y2[j] = __fmul2_rn(dsk2, make_float2(uv[j].x, uv[j].y));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) y2[j] = __ffma2_rn(ea[j][k], dup(comp(cq[k], j)), y2[j]);
                }
                // sums over the 4 lanes of the pair, transposing: lane sg ends up with channel sg & 1
                const bool odd = (sg & 1) != 0;
                float w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    w[j] = (odd ? y2[j].y : y2[j].x) + __shfl_xor_sync(0xffffffffu, odd ? y2[j].x : y2[j].y, 1);
#pragma unroll
                for (int j = 0; j < 4; ++j) w[j] += __shfl_xor_sync(0xffffffffu, w[j], 2);
                if (sg < 2) *reinterpret_cast<float4 *>(&sm.y[2 * pair + sg][tq]) = make_float4(w[0], w[1], w[2], w[3]);
            }
        }
        __syncthreads();                                           // every warp is done with this buffer's B / C tiles
        if constexpr (MODE != 1) {
            for (int i = tid; i < CH * (kScT / 4); i += NT) {
                const int r = i / (kScT / 4), t = t0 + 4 * (i % (kScT / 4));
                const float4 v = *reinterpret_cast<const float4 *>(&sm.y[r][4 * (i % (kScT / 4))]);
                float *dst = y + (row0 + r) * (size_t)L + t;
                if (vec) {
                    if (t < L) *reinterpret_cast<float4 *>(dst) = v;
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (t + k < L) dst[k] = comp(v, k);
                }
            }
        }
        // (the next iteration's first __syncthreads orders these reads, and the scan's reads of `ud`, before the next writes)
    }
    if constexpr (MODE == 1) {
        *reinterpret_cast<float4 *>(carry_h + (carry0 + piece) * kScN + n0) = make_float4(h2[0].x, h2[1].x, h2[2].x, h2[3].x);
        *reinterpret_cast<float4 *>(carry_h + (carry0 + n_chunks + piece) * kScN + n0) = make_float4(h2[0].y, h2[1].y, h2[2].y, h2[3].y);
        if (sg == 0) {
            carry_s[carry0 + piece] = sum_delta2.x;
            carry_s[carry0 + n_chunks + piece] = sum_delta2.y;
        }
    }
}

// ---- backward
template <bool BF16, int NW> struct ScBwdSmem {
    static constexpr int CH = 16 * NW, NT = 32 * NW, T = kScSeg, TWIN = BF16 ? T / 2 : T, PIN = TWIN + 4;
    float4 ud[CH / 2][T + 1];                                    // {u0, u1, delta0, delta1} of a channel pair
    float4 gs[CH / 2][T + 1];                                    // {dy0, dy1, sigmoid0, sigmoid1}: d softplus / d raw
    float4 sub[T / kScSub][2][NT];                               // states before every 4th position, per thread
    float b[2][kScN][T], c[2][kScN][T];
    float out[NT][T + 4];                                        // row 4 pair + 2 e + kind: d_u (kind 0) / d_raw (1) of channel 2 pair + e
    float dbc[NW][2 * kScN][T + 4];                              // per-warp dB (rows 0-15) / dC (rows 16-31) of the segment
    uint32_t uin[2][CH][PIN], din[2][CH][PIN];                   // input rows as they arrive, double-buffered
    float dyin[2][CH][T + 4];
};

// slot k <- q[k ^ p]   (p in 0..3)
__device__ __forceinline__ float4 permute4(float4 q, int p) {
    if (p & 1) q = make_float4(q.y, q.x, q.w, q.z);
    if (p & 2) q = make_float4(q.z, q.w, q.x, q.y);
    return q;
}
__device__ __forceinline__ void red_add_v4(float *p, const float4 &v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// One segment (16 positions) at a time, last to first: its tiles are prefetched (cp.async) while the previous one is
// processed; positions past L are staged as zeros and contribute nothing, so every segment is walked in full.
template <bool BF16, int NW>
__global__ void __launch_bounds__(32 * NW, 2)
sscan_bwd_kernel(const void *__restrict__ u_, const void *__restrict__ dt_, const float *__restrict__ A,
                 const float *__restrict__ Bm, const float *__restrict__ Cm, const float *__restrict__ Dv,
                 const float *__restrict__ bias, const float *__restrict__ dy, const float *__restrict__ ckpt,
                 void *__restrict__ g_u, void *__restrict__ g_dt, float *__restrict__ g_A, float *__restrict__ g_B,
                 float *__restrict__ g_C, float *__restrict__ g_D, float *__restrict__ g_bias, int KD, int Dg, int L,
                 int n_seg, int vec_) {
    using SM = ScBwdSmem<BF16, NW>;
    constexpr int CH = 16 * NW, NT = 32 * NW, T = kScSeg, PAIRS = CH / 2;
    const bool vec = vec_ != 0;
    extern __shared__ __align__(16) unsigned char sc_raw[];
    SM &sm = *reinterpret_cast<SM *>(sc_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sg = lane & 3, cp = lane >> 2, pair = warp * 8 + cp, p = cp & 3;
    const int b = blockIdx.y, ch0 = blockIdx.x * CH, chA = ch0 + 2 * pair, n0 = 4 * sg;
    const size_t row0 = (size_t)b * KD + ch0;
    const size_t grp = (size_t)b * (KD / Dg) + ch0 / Dg;
    const int LW = BF16 ? L / 2 : L;
    // register slot k of this lane <-> state n0 + (k ^ p); values packed over the channel pair
    float2 A2[4], dh2[4], dA2[4];
    {
        const float4 a0 = permute4(__ldg(reinterpret_cast<const float4 *>(A + (size_t)chA * kScN + n0)), p);
        const float4 a1 = permute4(__ldg(reinterpret_cast<const float4 *>(A + (size_t)(chA + 1) * kScN + n0)), p);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            A2[k] = make_float2(comp(a0, k) * kLog2e, comp(a1, k) * kLog2e);
            dh2[k] = dA2[k] = make_float2(0.0f, 0.0f);
        }
    }
    // after the sum over the pair's 4 lanes, lane sg owns (channel sg >> 1, kind sg & 1): kind 0 = d_u, 1 = d_delta
    const int own_e = sg >> 1, own_kind = sg & 1;
    const float own_dsk = Dv != nullptr ? __ldg(Dv + chA + own_e) : 0.0f;
    float own_acc = 0.0f;                                          // kind 0: dD of the channel, kind 1: d bias
    const int cpair = tid % PAIRS;                                 // conversion pass: thread <-> (channel pair, 4 positions)
    const float cbias0 = bias != nullptr ? __ldg(bias + ch0 + 2 * cpair) : 0.0f;
    const float cbias1 = bias != nullptr ? __ldg(bias + ch0 + 2 * cpair + 1) : 0.0f;
    const float *ck = ckpt + ((size_t)b * KD + chA) * (size_t)n_seg * kScN + 2 * n0;       // layout: forward kernel
    float *out_row = &sm.out[tid][0];                              // (4 pair + 2 own_e + own_kind = tid)
    // after the sum over channels lanes cp < 4 hold dB[n0 + p], lanes cp >= 4 dC[n0 + p]
    float *dbc_row = &sm.dbc[warp][(cp < 4 ? 0 : kScN) + n0 + p][0];
    int brow[4];                                                   // slot k reads row n0 + (k ^ p) of the B / C tiles
#pragma unroll
    for (int k = 0; k < 4; ++k) brow[k] = (n0 + (k ^ p)) * T;

    auto prefetch = [&](int buf, int t0) {
        const int w0 = BF16 ? t0 / 2 : t0;
        stage_rows<CH, SM::TWIN, NT>(&sm.uin[buf][0][0], reinterpret_cast<const uint32_t *>(u_), row0, LW, w0, vec);
        stage_rows<CH, SM::TWIN, NT>(&sm.din[buf][0][0], reinterpret_cast<const uint32_t *>(dt_), row0, LW, w0, vec);
        stage_rows<CH, T, NT>(reinterpret_cast<uint32_t *>(&sm.dyin[buf][0][0]), reinterpret_cast<const uint32_t *>(dy), row0, L, t0, vec);
        stage_bc<T, NT>(&sm.b[buf][0][0], Bm, grp, L, t0, vec);
        stage_bc<T, NT>(&sm.c[buf][0][0], Cm, grp, L, t0, vec);
        cp_async_commit();
    };
    auto load_ck = [&](int seg, float4 (&q)[2]) {
        q[0] = __ldg(reinterpret_cast<const float4 *>(ck + (size_t)seg * (2 * kScN)));      // states n0, n0 + 1 x 2 channels
        q[1] = __ldg(reinterpret_cast<const float4 *>(ck + (size_t)seg * (2 * kScN) + 4));  // states n0 + 2, n0 + 3
    };
    prefetch(0, (n_seg - 1) * kScSeg);
    float4 ck_cur[2], ck_nxt[2];
    load_ck(n_seg - 1, ck_cur);
    ck_nxt[0] = ck_nxt[1] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    int buf = 0;
    for (int seg = n_seg - 1; seg >= 0; --seg, buf ^= 1) {
        const int t0 = seg * kScSeg;
        if (seg > 0) { prefetch(buf ^ 1, t0 - kScSeg); load_ck(seg - 1, ck_nxt); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();                                           // (also: the previous segment's outputs were stored)
        // ---- conversion pass: a thread converts 4 positions of a channel pair (lanes run over pairs)
        for (int c4 = tid / PAIRS; c4 < T / 4; c4 += NT / PAIRS) {
            float uu[2][4], dd[2][4], gg[2][4];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int r = 2 * cpair + e;
                if constexpr (BF16) {
                    const uint2 wu = *reinterpret_cast<const uint2 *>(&sm.uin[buf][r][2 * c4]);
                    const uint2 wd = *reinterpret_cast<const uint2 *>(&sm.din[buf][r][2 * c4]);
                    uu[e][0] = bf16lo(wu.x); uu[e][1] = bf16hi(wu.x); uu[e][2] = bf16lo(wu.y); uu[e][3] = bf16hi(wu.y);
                    dd[e][0] = bf16lo(wd.x); dd[e][1] = bf16hi(wd.x); dd[e][2] = bf16lo(wd.y); dd[e][3] = bf16hi(wd.y);
                } else {
                    const float4 fu = *reinterpret_cast<const float4 *>(&sm.uin[buf][r][4 * c4]);
                    const float4 fd = *reinterpret_cast<const float4 *>(&sm.din[buf][r][4 * c4]);
                    uu[e][0] = fu.x; uu[e][1] = fu.y; uu[e][2] = fu.z; uu[e][3] = fu.w;
                    dd[e][0] = fd.x; dd[e][1] = fd.y; dd[e][2] = fd.z; dd[e][3] = fd.w;
                }
                const float4 fg = *reinterpret_cast<const float4 *>(&sm.dyin[buf][r][4 * c4]);   // zero-filled past L
                gg[e][0] = fg.x; gg[e][1] = fg.y; gg[e][2] = fg.z; gg[e][3] = fg.w;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool live = t0 + 4 * c4 + k < L;
                float s0, s1;
                const float d0 = softplus20(dd[0][k] + cbias0, &s0), d1 = softplus20(dd[1][k] + cbias1, &s1);
                sm.ud[cpair][4 * c4 + k] = make_float4(live ? uu[0][k] : 0.0f, live ? uu[1][k] : 0.0f, live ? d0 : 0.0f, live ? d1 : 0.0f);
                sm.gs[cpair][4 * c4 + k] = make_float4(gg[0][k], gg[1][k], s0, s1);
            }
        }
        __syncthreads();
        const float *bt = &sm.b[buf][0][0], *ct = &sm.c[buf][0][0];
        // ---- pass 1: recompute the segment forward from its checkpoint, keeping the state before every 4th position
        {
            float2 h2[4];
            {
                float2 q[4] = {make_float2(ck_cur[0].x, ck_cur[0].y), make_float2(ck_cur[0].z, ck_cur[0].w),
                               make_float2(ck_cur[1].x, ck_cur[1].y), make_float2(ck_cur[1].z, ck_cur[1].w)};
                if (p & 1) { const float2 t0_ = q[0], t2_ = q[2]; q[0] = q[1]; q[1] = t0_; q[2] = q[3]; q[3] = t2_; }
                if (p & 2) { const float2 t0_ = q[0], t1_ = q[1]; q[0] = q[2]; q[1] = q[3]; q[2] = t0_; q[3] = t1_; }
#pragma unroll
                for (int k = 0; k < 4; ++k) h2[k] = q[k];                                     // slot k <- state n0 + (k ^ p)
            }
#pragma unroll 1
            for (int g0 = 0; g0 < T - kScSub; g0 += kScSub) {      // (the last group's entry state is all pass 2 needs)
                sm.sub[g0 / kScSub][0][tid] = make_float4(h2[0].x, h2[0].y, h2[1].x, h2[1].y);
                sm.sub[g0 / kScSub][1][tid] = make_float4(h2[2].x, h2[2].y, h2[3].x, h2[3].y);
                float4 uv[4], bk[4];
                const int chunk = ((g0 >> 2) ^ sg) << 2;
#pragma unroll
                for (int j = 0; j < 4; ++j) uv[j] = sm.ud[pair][g0 + j];
#pragma unroll
                for (int k = 0; k < 4; ++k) bk[k] = *reinterpret_cast<const float4 *>(bt + brow[k] + chunk);
                float2 ea[4][4], db[4][4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 D2 = make_float2(uv[j].z, uv[j].w);
#pragma unroll
                    for (int k = 0; k < 4; ++k) ea[j][k] = __fmul2_rn(D2, A2[k]);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) ea[j][k] = ex2(ea[j][k]);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 DU2 = __fmul2_rn(make_float2(uv[j].z, uv[j].w), make_float2(uv[j].x, uv[j].y));
#pragma unroll
                    for (int k = 0; k < 4; ++k) db[j][k] = __fmul2_rn(DU2, dup(comp(bk[k], j)));
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) h2[k] = __ffma2_rn(ea[j][k], h2[k], db[j][k]);
                }
            }
            sm.sub[T / kScSub - 1][0][tid] = make_float4(h2[0].x, h2[0].y, h2[1].x, h2[1].y);
            sm.sub[T / kScSub - 1][1][tid] = make_float4(h2[2].x, h2[2].y, h2[3].x, h2[3].y);
        }
        // ---- pass 2: groups of 4 positions, last to first (each thread reads back only its own sub-checkpoints):
        //   F  forward recompute of the group (exponentials first, then the recurrence), operands kept in registers;
        //   R  reverse walk: the only chain is dh <- (dh + C dy) a; everything else hangs off it per position;
        //   S  the sums over lanes (states: 3 shuffles deep 2; channels: 8 shuffles deep 3) for the 4 positions at once.
#pragma unroll 1
        for (int g0 = T - kScSub; g0 >= 0; g0 -= kScSub) {
            float2 hist[kScSub + 1][4];                          // hist[j] = state before position g0 + j
            float2 an[kScSub][4];                                // exp(delta*A) of the group, reused by the reverse walk
            float4 uv[kScSub], gv[kScSub], bk[4], cq[4];
            const int chunk = ((g0 >> 2) ^ sg) << 2;
            {
                const float4 q0 = sm.sub[g0 / kScSub][0][tid], q1 = sm.sub[g0 / kScSub][1][tid];
                hist[0][0] = make_float2(q0.x, q0.y);
                hist[0][1] = make_float2(q0.z, q0.w);
                hist[0][2] = make_float2(q1.x, q1.y);
                hist[0][3] = make_float2(q1.z, q1.w);
            }
#pragma unroll
            for (int j = 0; j < kScSub; ++j) uv[j] = sm.ud[pair][g0 + j];
#pragma unroll
            for (int k = 0; k < 4; ++k) bk[k] = *reinterpret_cast<const float4 *>(bt + brow[k] + chunk);
#pragma unroll
            for (int j = 0; j < kScSub; ++j) gv[j] = sm.gs[pair][g0 + j];
#pragma unroll
            for (int k = 0; k < 4; ++k) cq[k] = *reinterpret_cast<const float4 *>(ct + brow[k] + chunk);
#pragma unroll
            for (int j = 0; j < kScSub; ++j) {
                const float2 D2 = make_float2(uv[j].z, uv[j].w);
#pragma unroll
                for (int k = 0; k < 4; ++k) an[j][k] = __fmul2_rn(D2, A2[k]);
            }
#pragma unroll
            for (int j = 0; j < kScSub; ++j) {
#pragma unroll
                for (int k = 0; k < 4; ++k) an[j][k] = ex2(an[j][k]);
            }
            {
                float2 db[kScSub][4];
#pragma unroll
                for (int j = 0; j < kScSub; ++j) {
                    const float2 DU2 = __fmul2_rn(make_float2(uv[j].z, uv[j].w), make_float2(uv[j].x, uv[j].y));
#pragma unroll
                    for (int k = 0; k < 4; ++k) db[j][k] = __fmul2_rn(DU2, dup(comp(bk[k], j)));
                }
#pragma unroll
                for (int j = 0; j < kScSub; ++j) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) hist[j + 1][k] = __ffma2_rn(an[j][k], hist[j][k], db[j][k]);
                }
            }
            float rB[kScSub][4], rC[kScSub][4];                  // this thread's dB / dC contributions (both channels added)
            float2 vdu[kScSub], vdd[kScSub];                     // {channel 0, channel 1} partial sums of d_u, d_delta
#pragma unroll
            for (int j = kScSub - 1; j >= 0; --j) {
                const float2 U2 = make_float2(uv[j].x, uv[j].y), D2 = make_float2(uv[j].z, uv[j].w);
                const float2 GY2 = make_float2(gv[j].x, gv[j].y), DLU2 = __fmul2_rn(D2, U2);
                float2 dah[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) dh2[k] = __ffma2_rn(GY2, dup(comp(cq[k], j)), dh2[k]);        // dL/dh_t
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 c2 = __fmul2_rn(GY2, hist[j + 1][k]);
                    rC[j][k] = c2.x + c2.y;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 b2 = __fmul2_rn(dh2[k], DLU2);
                    rB[j][k] = b2.x + b2.y;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) dah[k] = __fmul2_rn(dh2[k], an[j][k]);
#pragma unroll
                for (int k = 0; k < 4; ++k) dah[k] = __fmul2_rn(dah[k], hist[j][k]);                      // dh * a * h_{t-1}
                float2 s1a = __fmul2_rn(dh2[0], dup(comp(bk[0], j))), s1b = __fmul2_rn(dh2[1], dup(comp(bk[1], j)));
                float2 s2a = __fmul2_rn(dah[0], A2[0]), s2b = __fmul2_rn(dah[1], A2[1]);
                s1a = __ffma2_rn(dh2[2], dup(comp(bk[2], j)), s1a);
                s1b = __ffma2_rn(dh2[3], dup(comp(bk[3], j)), s1b);
                s2a = __ffma2_rn(dah[2], A2[2], s2a);
                s2b = __ffma2_rn(dah[3], A2[3], s2b);
#pragma unroll
                for (int k = 0; k < 4; ++k) dA2[k] = __ffma2_rn(dah[k], D2, dA2[k]);
#pragma unroll
                for (int k = 0; k < 4; ++k) dh2[k] = __fmul2_rn(dh2[k], an[j][k]);                        // dL/dh_{t-1}
                const float2 s1 = __fadd2_rn(s1a, s1b), s2 = __fadd2_rn(s2a, s2b);                        // sum dh B | sum dah A log2 e
                vdu[j] = __fmul2_rn(D2, s1);                                                 // d_u (without the skip term)
                vdd[j] = __ffma2_rn(dup(kLn2), s2, __fmul2_rn(U2, s1));                      // d_delta
            }
            // sums over the 16 states = over the pair's 4 lanes, transposing: lane sg ends up with value index sg =
            // 2 channel + kind
            {
                const bool o1 = (sg & 1) != 0, o2 = (sg & 2) != 0;
                float w0[kScSub], w1[kScSub], r[kScSub];
#pragma unroll
                for (int j = 0; j < kScSub; ++j) {
                    w0[j] = (o1 ? vdd[j].x : vdu[j].x) + __shfl_xor_sync(0xffffffffu, o1 ? vdu[j].x : vdd[j].x, 1);
                    w1[j] = (o1 ? vdd[j].y : vdu[j].y) + __shfl_xor_sync(0xffffffffu, o1 ? vdu[j].y : vdd[j].y, 1);
                }
#pragma unroll
                for (int j = 0; j < kScSub; ++j)
                    r[j] = (o2 ? w1[j] : w0[j]) + __shfl_xor_sync(0xffffffffu, o2 ? w0[j] : w1[j], 2);
#pragma unroll
                for (int j = 0; j < kScSub; ++j) {
                    const float gy_o = own_e ? gv[j].y : gv[j].x, sig_o = own_e ? gv[j].w : gv[j].z;
                    const float ut_o = own_e ? uv[j].y : uv[j].x;
                    r[j] = own_kind ? r[j] * sig_o : fmaf(own_dsk, gy_o, r[j]);              // d_raw = d_delta * sigmoid | d_u + D dy
                    own_acc += own_kind ? r[j] : gy_o * ut_o;                                // d bias | dD
                }
                *reinterpret_cast<float4 *>(out_row + g0) = make_float4(r[0], r[1], r[2], r[3]);
            }
            // sums over the warp's 16 channels: slot bit <-> lane bit, no selects (file header)
#pragma unroll
            for (int j = 0; j < kScSub; ++j) {
                rB[j][0] += __shfl_xor_sync(0xffffffffu, rB[j][2], 8);
                rB[j][1] += __shfl_xor_sync(0xffffffffu, rB[j][3], 8);
                rC[j][0] += __shfl_xor_sync(0xffffffffu, rC[j][2], 8);
                rC[j][1] += __shfl_xor_sync(0xffffffffu, rC[j][3], 8);
            }
#pragma unroll
            for (int j = 0; j < kScSub; ++j) {
                rB[j][0] += __shfl_xor_sync(0xffffffffu, rB[j][1], 4);
                rC[j][0] += __shfl_xor_sync(0xffffffffu, rC[j][1], 4);
            }
#pragma unroll
            for (int j = 0; j < kScSub; ++j) {
                rB[j][0] += __shfl_xor_sync(0xffffffffu, rB[j][0], 16);
                rC[j][0] += __shfl_xor_sync(0xffffffffu, rC[j][0], 16);
            }
            *reinterpret_cast<float4 *>(dbc_row + g0) = cp < 4 ? make_float4(rB[0][0], rB[1][0], rB[2][0], rB[3][0])
                                                               : make_float4(rC[0][0], rC[1][0], rC[2][0], rC[3][0]);
        }
        __syncthreads();
        // ---- the segment's outputs: d_u / d_dt rows (16-byte chunks), dB / dC summed over the CTA's warps
        for (int i = tid; i < NT * (T / 4); i += NT) {
            const int rho = i / (T / 4), c4 = i % (T / 4), t = t0 + 4 * c4;                  // rho = out row
            const int r = 2 * (rho >> 2) + ((rho >> 1) & 1);
            const float4 v = *reinterpret_cast<const float4 *>(&sm.out[rho][4 * c4]);
            void *dst = (rho & 1) ? g_dt : g_u;
            const size_t at = (row0 + r) * (size_t)L + t;
            if (vec) {
                if (t < L) {
                    if constexpr (BF16) {
                        const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
                        *reinterpret_cast<uint2 *>(reinterpret_cast<__nv_bfloat16 *>(dst) + at) =
                            make_uint2(*reinterpret_cast<const uint32_t *>(&lo), *reinterpret_cast<const uint32_t *>(&hi));
                    } else {
                        *reinterpret_cast<float4 *>(reinterpret_cast<float *>(dst) + at) = v;
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (t + k < L) sc_store<BF16>(dst, at + k, comp(v, k));
            }
        }
        for (int i = tid; i < 2 * kScN * (T / 4); i += NT) {                                 // rows 0-15: dB, 16-31: dC
            const int n = i / (T / 4), c4 = i % (T / 4), t = t0 + 4 * c4;
            float4 v = *reinterpret_cast<const float4 *>(&sm.dbc[0][n][4 * c4]);
#pragma unroll
            for (int w = 1; w < NW; ++w) {
                const float4 o = *reinterpret_cast<const float4 *>(&sm.dbc[w][n][4 * c4]);
                v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
            }
            float *dst = (n < kScN ? g_B + (grp * kScN + n) * (size_t)L : g_C + (grp * kScN + n - kScN) * (size_t)L) + t;
            if (vec) {
                if (t < L) red_add_v4(dst, v);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (t + k < L) atomicAdd(dst + k, comp(v, k));
            }
        }
        ck_cur[0] = ck_nxt[0];
        ck_cur[1] = ck_nxt[1];
    }
    {                                                                                        // over the batch
        float *d0 = g_A + (size_t)chA * kScN + n0, *d1 = d0 + kScN;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            atomicAdd(d0 + (k ^ p), dA2[k].x);
            atomicAdd(d1 + (k ^ p), dA2[k].y);
        }
    }
    if (own_kind == 0) {
        if (g_D != nullptr) atomicAdd(g_D + chA + own_e, own_acc);
    } else {
        if (g_bias != nullptr) atomicAdd(g_bias + chA + own_e, own_acc);
    }
}

}  // namespace tamtr

using namespace tamtr;

static int sscan_warps(int Dg) { return Dg % 64 == 0 ? 4 : 2; }            // channels per CTA = 16 * warps

static int sscan_check(int Bn, int KD, int Dg, int N, int L) {
    TAMTR_CHECK_ARG(Bn > 0 && KD > 0 && Dg > 0 && L > 0, TAMTR_E_BADARG, "selective_scan: non-positive size");
    TAMTR_CHECK_ARG(N == kScN, TAMTR_E_UNSUPPORTED, "selective_scan: d_state = %d unsupported (16)", N);
    TAMTR_CHECK_ARG(KD % Dg == 0 && Dg % 32 == 0, TAMTR_E_UNSUPPORTED,
                    "selective_scan: channels per direction (%d) must be a multiple of 32", Dg);
    TAMTR_CHECK_ARG(Bn <= 65535, TAMTR_E_UNSUPPORTED, "selective_scan: batch too large");
    return 0;
}

extern "C" int tamtr_selective_scan_segments(int L) { return L > 0 ? (L + kScSeg - 1) / kScSeg : 0; }

// 16-byte staging / stores: rows of 4 (fp32) or 8 (bf16) positions at a time, every base address 16-byte aligned
static int sscan_vec(int in_dtype, int L, std::initializer_list<const void *> ptrs) {
    if (L % (in_dtype == TAMTR_BF16 ? 8 : 4) != 0) return 0;
    for (const void *q : ptrs)
        if (q != nullptr && ((uintptr_t)q & 15) != 0) return 0;
    return 1;
}

static int sscan_check_in(int in_dtype, const void *u, const void *dt, int L) {
    TAMTR_CHECK_ARG(in_dtype == TAMTR_F32 || in_dtype == TAMTR_BF16, TAMTR_E_UNSUPPORTED, "selective_scan: dtype %d", in_dtype);
    if (in_dtype == TAMTR_BF16)
        TAMTR_CHECK_ARG(L % 2 == 0 && (((uintptr_t)u | (uintptr_t)dt) & 3) == 0, TAMTR_E_UNSUPPORTED,
                        "selective_scan: bf16 inputs need an even L (%d) and 4-byte aligned rows", L);
    return 0;
}

// Dynamic shared memory above 48 KB needs an opt-in per kernel and device; `done` is the caller's per-instantiation flag
// array (a function-local static of a template on the kernel's TYPE would be shared by all kernels of one signature).
static cudaError_t sscan_opt_in(const void *kernel, size_t bytes, bool (&done)[64]) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true;
    return e;
}

template <bool BF16, int MODE, int NW>
static cudaError_t sscan_fwd_launch(dim3 grid, cudaStream_t st, const void *u, const void *dt, const float *A, const float *Bm,
                                    const float *Cm, const float *D, const float *bias, float *y, float *ckpt, int KD, int Dg,
                                    int L, int n_seg, float *carry_h, float *carry_s, int n_chunks, int chunk_len, int vec) {
    auto kernel = sscan_fwd_kernel<BF16, MODE, NW>;
    static bool done[64] = {false};
    cudaError_t e = sscan_opt_in((const void *)kernel, sizeof(ScFwdSmem<BF16, NW>), done);
    if (e != cudaSuccess) return e;
    kernel<<<grid, 32 * NW, sizeof(ScFwdSmem<BF16, NW>), st>>>(u, dt, A, Bm, Cm, D, bias, y, ckpt, KD, Dg, L, n_seg, carry_h,
                                                              carry_s, n_chunks, chunk_len, vec);
    count_launch();
    return cudaGetLastError();
}
template <int MODE, typename... Args>
static cudaError_t sscan_fwd_dispatch(int in_dtype, int nw, Args... args) {
    if (in_dtype == TAMTR_BF16)
        return nw == 4 ? sscan_fwd_launch<true, MODE, 4>(args...) : sscan_fwd_launch<true, MODE, 2>(args...);
    return nw == 4 ? sscan_fwd_launch<false, MODE, 4>(args...) : sscan_fwd_launch<false, MODE, 2>(args...);
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
    const int nw = sscan_warps(Dg);
    KernelTimer timer(K_SSCAN_FWD, st);
    TAMTR_CUDA_OK(sscan_fwd_dispatch<0>(in_dtype, nw, dim3(KD / (16 * nw), Bn), st, u, dt, A, Bm, Cm, D, bias, y, ckpt, KD, Dg, L,
                                        tamtr_selective_scan_segments(L), (float *)nullptr, (float *)nullptr, 1, L,
                                        sscan_vec(in_dtype, L, {u, dt, Bm, Cm, y})));
    return 0;
}

// How many pieces the chunk-parallel inference forward should cut the sequence into: 1 (= use the plain forward) when the
// (channel block, image) grid already fills the GPU, otherwise enough pieces for about four CTAs per SM, at most 32, each
// at least 1024 positions.
extern "C" int tamtr_selective_scan_chunks(int Bn, int KD, int L) {
    if (Bn <= 0 || KD < 32 || L <= 0) return 1;
    const long ctas = (long)(KD / 64 > 0 ? KD / 64 : 1) * Bn;
    const long want = 2L * ::tamtr::sm_count();
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
    cudaStream_t st = (cudaStream_t)stream;
    const int chunk_len = (((L + n_chunks - 1) / n_chunks + kScT - 1) / kScT) * kScT;
    const int pieces = (L + chunk_len - 1) / chunk_len;              // <= n_chunks; the carry arrays keep stride n_chunks
    float *carry_h = carry, *carry_s = carry + (size_t)Bn * KD * n_chunks * kScN;
    const int nseg = tamtr_selective_scan_segments(L);
    const int nw = sscan_warps(Dg), vec = sscan_vec(in_dtype, L, {u, dt, Bm, Cm, y});
    KernelTimer timer(K_SSCAN_FWD, st);
    if (pieces > 1)
        TAMTR_CUDA_OK(sscan_fwd_dispatch<1>(in_dtype, nw, dim3(KD / (16 * nw), Bn, pieces - 1), st, u, dt, A, Bm, Cm, D, bias, y,
                                            (float *)nullptr, KD, Dg, L, nseg, carry_h, carry_s, n_chunks, chunk_len, vec));
    TAMTR_CUDA_OK(sscan_fwd_dispatch<2>(in_dtype, nw, dim3(KD / (16 * nw), Bn, pieces), st, u, dt, A, Bm, Cm, D, bias, y,
                                        (float *)nullptr, KD, Dg, L, nseg, carry_h, carry_s, n_chunks, chunk_len, vec));
    return 0;
}

template <bool BF16, int NW, typename... Args>
static cudaError_t sscan_bwd_launch(dim3 grid, cudaStream_t st, Args... args) {
    auto kernel = sscan_bwd_kernel<BF16, NW>;
    static bool done[64] = {false};
    cudaError_t e = sscan_opt_in((const void *)kernel, sizeof(ScBwdSmem<BF16, NW>), done);
    if (e != cudaSuccess) return e;
    kernel<<<grid, 32 * NW, sizeof(ScBwdSmem<BF16, NW>), st>>>(args...);
    count_launch();
    return cudaGetLastError();
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
    const int nw = sscan_warps(Dg), nseg = tamtr_selective_scan_segments(L);
    const int vec = sscan_vec(in_dtype, L, {u, dt, Bm, Cm, dy, g_u, g_dt, g_B, g_C});
    const dim3 grid(KD / (16 * nw), Bn);
    KernelTimer timer(K_SSCAN_BWD, st);
    cudaError_t e;
    if (in_dtype == TAMTR_BF16)
        e = nw == 4 ? sscan_bwd_launch<true, 4>(grid, st, u, dt, A, Bm, Cm, D, bias, dy, ckpt, g_u, g_dt, g_A, g_B, g_C, g_D, g_bias, KD, Dg, L, nseg, vec)
                    : sscan_bwd_launch<true, 2>(grid, st, u, dt, A, Bm, Cm, D, bias, dy, ckpt, g_u, g_dt, g_A, g_B, g_C, g_D, g_bias, KD, Dg, L, nseg, vec);
    else
        e = nw == 4 ? sscan_bwd_launch<false, 4>(grid, st, u, dt, A, Bm, Cm, D, bias, dy, ckpt, g_u, g_dt, g_A, g_B, g_C, g_D, g_bias, KD, Dg, L, nseg, vec)
                    : sscan_bwd_launch<false, 2>(grid, st, u, dt, A, Bm, Cm, D, bias, dy, ckpt, g_u, g_dt, g_A, g_B, g_C, g_D, g_bias, KD, Dg, L, nseg, vec);
    TAMTR_CUDA_OK(e);
    return 0;
}
