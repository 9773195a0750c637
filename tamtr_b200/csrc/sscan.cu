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
// Mapping: thread = one channel with its 16 states in registers, CTA = 128 consecutive channels of one (image, scan
// direction), walking the L positions in tiles of 32.  Tiles of u / dt / dy are staged through shared memory with
// coalesced row reads (L is the contiguous dimension) and read back conflict-free (pitch 33); the B / C rows of the
// group are staged once per tile and read as broadcasts.  The bound is the SFU: 16 exp per position and channel.
// The forward also writes the state every 64 positions; the backward walks the segments in reverse, recomputes the
// states of a segment from its checkpoint (sub-checkpoints every 4 positions in shared memory, the 4 positions
// in registers) and accumulates dB / dC across the CTA's channels with a 32-value warp transpose-reduction.
#include "common.cuh"

namespace tamtr {

constexpr int kScN = 16;          // states
constexpr int kScCh = 128;        // channels (threads) per CTA
constexpr int kScT = 32;          // positions per tile
constexpr int kScSeg = 64;        // positions per checkpoint segment
constexpr int kScSub = 4;         // positions recomputed into registers at a time (backward)
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float softplus20(float x) { return x > 20.0f ? x : log1pf(__expf(x)); }

// stage rows [ch0, ch0+128) x positions [t0, t0+32) of a [rows, L] array into tile[128][33] (zero past L)
__device__ __forceinline__ void stage_rows(float (*tile)[kScT + 1], const float *__restrict__ src, size_t row0, int L,
                                           int t0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < kScCh; r += kScCh / 32) {
        const int t = t0 + lane;
        tile[r][lane] = t < L ? __ldg(src + (row0 + r) * (size_t)L + t) : 0.0f;
    }
}

// tile[16][32] of B or C for group (b, k): src [b, K, N, L]
__device__ __forceinline__ void stage_bc(float (*tile)[kScT], const float *__restrict__ src, size_t grp, int L, int t0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int n = warp; n < kScN; n += kScCh / 32) {
        const int t = t0 + lane;
        tile[n][lane] = t < L ? __ldg(src + (grp * kScN + n) * (size_t)L + t) : 0.0f;
    }
}

__global__ void __launch_bounds__(kScCh)
sscan_fwd_kernel(const float *__restrict__ u, const float *__restrict__ dt, const float *__restrict__ A,
                 const float *__restrict__ Bm, const float *__restrict__ Cm, const float *__restrict__ Dv,
                 const float *__restrict__ bias, float *__restrict__ y, float *__restrict__ ckpt, int KD, int Dg, int L,
                 int n_seg) {
    __shared__ float s_u[kScCh][kScT + 1], s_dt[kScCh][kScT + 1];   // y overwrites u in place (static smem <= 48 KB)
    __shared__ float s_b[kScN][kScT], s_c[kScN][kScT];
    const int b = blockIdx.y, ch0 = blockIdx.x * kScCh, ch = ch0 + threadIdx.x;
    const size_t row0 = (size_t)b * KD + ch0;
    const size_t grp = (size_t)b * (KD / Dg) + ch0 / Dg;
    float a2[kScN], h[kScN];
#pragma unroll
    for (int n = 0; n < kScN; ++n) { a2[n] = __ldg(A + (size_t)ch * kScN + n) * kLog2e; h[n] = 0.0f; }
    const float dsk = Dv != nullptr ? __ldg(Dv + ch) : 0.0f, bs = bias != nullptr ? __ldg(bias + ch) : 0.0f;
    float *ck = ckpt != nullptr ? ckpt + ((size_t)b * KD + ch) * (size_t)n_seg * kScN : nullptr;

    for (int t0 = 0; t0 < L; t0 += kScT) {
        __syncthreads();
        stage_rows(s_u, u, row0, L, t0);
        stage_rows(s_dt, dt, row0, L, t0);
        stage_bc(s_b, Bm, grp, L, t0);
        stage_bc(s_c, Cm, grp, L, t0);
        __syncthreads();
        if (ck != nullptr && t0 % kScSeg == 0) {            // state BEFORE position t0
            float4 *dst = reinterpret_cast<float4 *>(ck + (size_t)(t0 / kScSeg) * kScN);
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[q] = make_float4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
        }
        const int tn = min(kScT, L - t0);
        for (int t = 0; t < tn; ++t) {
            const float ut = s_u[threadIdx.x][t];
            const float dl = softplus20(s_dt[threadIdx.x][t] + bs);
            const float du = dl * ut;
            float acc = dsk * ut;
#pragma unroll
            for (int n = 0; n < kScN; ++n) {
                h[n] = fmaf(exp2f(dl * a2[n]), h[n], du * s_b[n][t]);
                acc = fmaf(s_c[n][t], h[n], acc);
            }
            s_u[threadIdx.x][t] = acc;
        }
        __syncthreads();
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int r = warp; r < kScCh; r += kScCh / 32)
            if (t0 + lane < L) y[(row0 + r) * (size_t)L + t0 + lane] = s_u[r][lane];
    }
}

// sum over the 32 lanes of v[i] -> returned to lane i (31 shuffles for 32 values)
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float send = up ? v[i] : v[i + s];
            const float got = __shfl_xor_sync(0xffffffffu, send, s);
            v[i] = (up ? v[i + s] : v[i]) + got;
        }
    }
    return v[0];
}

struct ScBwdSmem {
    float u[kScCh][kScT + 1], dt[kScCh][kScT + 1], dy[kScCh][kScT + 1];   // current tile (reverse order)
    float du[kScCh][kScT + 1], ddt[kScCh][kScT + 1];                      // outputs of the tile
    float b[kScN][kScT], c[kScN][kScT];
    float db[kScN][kScT], dc[kScN][kScT];                                // CTA-level dB / dC of the tile
    float sub[kScSeg / kScSub][kScN][kScCh];                             // states before every 4th position of a segment
};

__global__ void __launch_bounds__(kScCh)
sscan_bwd_kernel(const float *__restrict__ u, const float *__restrict__ dt, const float *__restrict__ A,
                 const float *__restrict__ Bm, const float *__restrict__ Cm, const float *__restrict__ Dv,
                 const float *__restrict__ bias, const float *__restrict__ dy, const float *__restrict__ ckpt,
                 float *__restrict__ g_u, float *__restrict__ g_dt, float *__restrict__ g_A, float *__restrict__ g_B,
                 float *__restrict__ g_C, float *__restrict__ g_D, float *__restrict__ g_bias, int KD, int Dg, int L,
                 int n_seg) {
    extern __shared__ unsigned char sc_raw[];
    ScBwdSmem &sm = *reinterpret_cast<ScBwdSmem *>(sc_raw);
    const int b = blockIdx.y, ch0 = blockIdx.x * kScCh, ch = ch0 + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t row0 = (size_t)b * KD + ch0;
    const size_t grp = (size_t)b * (KD / Dg) + ch0 / Dg;
    float a1[kScN], a2[kScN], dh[kScN], dA[kScN];
#pragma unroll
    for (int n = 0; n < kScN; ++n) {
        a1[n] = __ldg(A + (size_t)ch * kScN + n);
        a2[n] = a1[n] * kLog2e;
        dh[n] = dA[n] = 0.0f;
    }
    const float dsk = Dv != nullptr ? __ldg(Dv + ch) : 0.0f, bs = bias != nullptr ? __ldg(bias + ch) : 0.0f;
    float dD = 0.0f, dbias = 0.0f;
    const float *ck = ckpt + ((size_t)b * KD + ch) * (size_t)n_seg * kScN;

    for (int seg = n_seg - 1; seg >= 0; --seg) {
        const int s0 = seg * kScSeg, sn = min(kScSeg, L - s0);
        // ---- pass 1: recompute the segment forward, keep the state before every 4th position
        float h[kScN];
        {
            const float4 *src = reinterpret_cast<const float4 *>(ck + (size_t)seg * kScN);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 v = src[q];
                h[4 * q] = v.x; h[4 * q + 1] = v.y; h[4 * q + 2] = v.z; h[4 * q + 3] = v.w;
            }
        }
        for (int t0 = s0; t0 < s0 + sn; t0 += kScT) {
            __syncthreads();
            stage_rows(sm.u, u, row0, L, t0);
            stage_rows(sm.dt, dt, row0, L, t0);
            stage_bc(sm.b, Bm, grp, L, t0);
            __syncthreads();
            const int tn = min(kScT, s0 + sn - t0);
            for (int t = 0; t < tn; ++t) {
                if (((t0 - s0 + t) & (kScSub - 1)) == 0) {
#pragma unroll
                    for (int n = 0; n < kScN; ++n) sm.sub[(t0 - s0 + t) / kScSub][n][threadIdx.x] = h[n];
                }
                const float dl = softplus20(sm.dt[threadIdx.x][t] + bs);
                const float du = dl * sm.u[threadIdx.x][t];
#pragma unroll
                for (int n = 0; n < kScN; ++n) h[n] = fmaf(exp2f(dl * a2[n]), h[n], du * sm.b[n][t]);
            }
        }
        // ---- pass 2: tiles of the segment in reverse; inside a tile, groups of 4 positions in reverse
        for (int t0 = s0 + ((sn - 1) / kScT) * kScT; t0 >= s0; t0 -= kScT) {
            __syncthreads();
            stage_rows(sm.u, u, row0, L, t0);
            stage_rows(sm.dt, dt, row0, L, t0);
            stage_rows(sm.dy, dy, row0, L, t0);
            stage_bc(sm.b, Bm, grp, L, t0);
            stage_bc(sm.c, Cm, grp, L, t0);
            for (int i = threadIdx.x; i < kScN * kScT; i += kScCh) { (&sm.db[0][0])[i] = 0.0f; (&sm.dc[0][0])[i] = 0.0f; }
            __syncthreads();
            const int tn = min(kScT, s0 + sn - t0);
            for (int g0 = ((tn - 1) / kScSub) * kScSub; g0 >= 0; g0 -= kScSub) {
                const int gn = min(kScSub, tn - g0);
                // states h_{t-1} (hist[j]) for the positions of the group, from the sub-checkpoint
                float hist[kScSub + 1][kScN];
#pragma unroll
                for (int n = 0; n < kScN; ++n) hist[0][n] = sm.sub[(t0 - s0 + g0) / kScSub][n][threadIdx.x];
#pragma unroll
                for (int j = 0; j < kScSub; ++j) {
                    const float dl = j < gn ? softplus20(sm.dt[threadIdx.x][g0 + j] + bs) : 0.0f;
                    const float du = dl * (j < gn ? sm.u[threadIdx.x][g0 + j] : 0.0f);
#pragma unroll
                    for (int n = 0; n < kScN; ++n)
                        hist[j + 1][n] = j < gn ? fmaf(exp2f(dl * a2[n]), hist[j][n], du * sm.b[n][g0 + j]) : hist[j][n];
                }
#pragma unroll
                for (int j = kScSub - 1; j >= 0; --j) {
                    if (j >= gn) continue;
                    const int t = g0 + j;
                    const float ut = sm.u[threadIdx.x][t], raw = sm.dt[threadIdx.x][t] + bs;
                    const float dl = softplus20(raw);
                    const float gy = sm.dy[threadIdx.x][t];
                    float d_dl = 0.0f, d_u = dsk * gy;
                    float red[32];
#pragma unroll
                    for (int n = 0; n < kScN; ++n) {
                        const float an = exp2f(dl * a2[n]);
                        const float bn = sm.b[n][t];
                        dh[n] = fmaf(sm.c[n][t], gy, dh[n]);                    // dL/dh_t
                        red[kScN + n] = gy * hist[j + 1][n];                     // dC contribution
                        red[n] = dh[n] * dl * ut;                                // dB contribution
                        const float dah = dh[n] * an * hist[j][n];               // dh * a * h_{t-1}
                        d_dl = fmaf(dah, a1[n], fmaf(dh[n] * bn, ut, d_dl));
                        dA[n] = fmaf(dah, dl, dA[n]);
                        d_u = fmaf(dh[n] * bn, dl, d_u);
                        dh[n] *= an;                                             // dL/dh_{t-1}
                    }
                    const float r = warp_transpose_sum(red, lane);               // lane i: sum over the warp's channels
                    if (lane < kScN) atomicAdd(&sm.db[lane][t], r); else atomicAdd(&sm.dc[lane - kScN][t], r);
                    dD = fmaf(gy, ut, dD);
                    const float d_raw = raw > 20.0f ? d_dl : d_dl * (1.0f / (1.0f + __expf(-raw)));
                    dbias += d_raw;
                    sm.du[threadIdx.x][t] = d_u;
                    sm.ddt[threadIdx.x][t] = d_raw;
                }
            }
            __syncthreads();
            for (int r = warp; r < kScCh; r += kScCh / 32) {
                if (t0 + lane < s0 + sn) {
                    g_u[(row0 + r) * (size_t)L + t0 + lane] = sm.du[r][lane];
                    g_dt[(row0 + r) * (size_t)L + t0 + lane] = sm.ddt[r][lane];
                }
            }
            for (int n = warp; n < kScN; n += kScCh / 32) {
                if (t0 + lane < s0 + sn) {
                    atomicAdd(g_B + (grp * kScN + n) * (size_t)L + t0 + lane, sm.db[n][lane]);
                    atomicAdd(g_C + (grp * kScN + n) * (size_t)L + t0 + lane, sm.dc[n][lane]);
                }
            }
        }
    }
#pragma unroll
    for (int n = 0; n < kScN; ++n) atomicAdd(g_A + (size_t)ch * kScN + n, dA[n]);   // over the batch
    if (g_D != nullptr) atomicAdd(g_D + ch, dD);
    if (g_bias != nullptr) atomicAdd(g_bias + ch, dbias);
}

}  // namespace tamtr

using namespace tamtr;

static int sscan_check(int Bn, int KD, int Dg, int N, int L) {
    TAMTR_CHECK_ARG(Bn > 0 && KD > 0 && Dg > 0 && L > 0, TAMTR_E_BADARG, "selective_scan: non-positive size");
    TAMTR_CHECK_ARG(N == kScN, TAMTR_E_UNSUPPORTED, "selective_scan: d_state = %d unsupported (16)", N);
    TAMTR_CHECK_ARG(KD % Dg == 0 && Dg % kScCh == 0, TAMTR_E_UNSUPPORTED,
                    "selective_scan: channels per direction (%d) must be a multiple of %d", Dg, kScCh);
    TAMTR_CHECK_ARG(Bn <= 65535, TAMTR_E_UNSUPPORTED, "selective_scan: batch too large");
    return 0;
}

extern "C" int tamtr_selective_scan_segments(int L) { return L > 0 ? (L + kScSeg - 1) / kScSeg : 0; }

extern "C" int tamtr_selective_scan_forward(const float *u, const float *dt, const float *A, const float *Bm,
                                            const float *Cm, const float *D, const float *bias, float *y, float *ckpt,
                                            int Bn, int KD, int Dg, int N, int L, void *stream) {
    TAMTR_CHECK_ARG(u && dt && A && Bm && Cm && y, TAMTR_E_BADARG, "selective_scan_forward: null pointer");
    const int rc = sscan_check(Bn, KD, Dg, N, L);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    {
        KernelTimer timer(K_SSCAN_FWD, st);
        sscan_fwd_kernel<<<dim3(KD / kScCh, Bn), kScCh, 0, st>>>(u, dt, A, Bm, Cm, D, bias, y, ckpt, KD, Dg, L,
                                                                tamtr_selective_scan_segments(L));
    }
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_selective_scan_backward(const float *u, const float *dt, const float *A, const float *Bm,
                                             const float *Cm, const float *D, const float *bias, const float *dy,
                                             const float *ckpt, float *g_u, float *g_dt, float *g_A, float *g_B,
                                             float *g_C, float *g_D, float *g_bias, int Bn, int KD, int Dg, int N, int L,
                                             void *stream) {
    TAMTR_CHECK_ARG(u && dt && A && Bm && Cm && dy && ckpt && g_u && g_dt && g_A && g_B && g_C, TAMTR_E_BADARG,
                    "selective_scan_backward: null pointer");
    const int rc = sscan_check(Bn, KD, Dg, N, L);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t grp_elems = (size_t)Bn * (KD / Dg) * kScN * L;
    TAMTR_CUDA_OK(cudaMemsetAsync(g_A, 0, (size_t)KD * kScN * sizeof(float), st));
    TAMTR_CUDA_OK(cudaMemsetAsync(g_B, 0, grp_elems * sizeof(float), st));
    TAMTR_CUDA_OK(cudaMemsetAsync(g_C, 0, grp_elems * sizeof(float), st));
    if (g_D) TAMTR_CUDA_OK(cudaMemsetAsync(g_D, 0, (size_t)KD * sizeof(float), st));
    if (g_bias) TAMTR_CUDA_OK(cudaMemsetAsync(g_bias, 0, (size_t)KD * sizeof(float), st));
    static bool attr_set = false;
    if (!attr_set) {
        TAMTR_CUDA_OK(cudaFuncSetAttribute(sscan_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(ScBwdSmem)));
        attr_set = true;
    }
    {
        KernelTimer timer(K_SSCAN_BWD, st);
        sscan_bwd_kernel<<<dim3(KD / kScCh, Bn), kScCh, sizeof(ScBwdSmem), st>>>(
            u, dt, A, Bm, Cm, D, bias, dy, ckpt, g_u, g_dt, g_A, g_B, g_C, g_D, g_bias, KD, Dg, L,
            tamtr_selective_scan_segments(L));
    }
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}
