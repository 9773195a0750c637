// Folded encoder-side projections of the detection heads on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
// The reference runs, on every pyramid level l (head.py:1202-1218, transformer.py:273, head.py:1229-1237):
//     Y = conv1x1(X_l)  ->  M = BatchNorm(Y)  ->  value_i = value_proj_i(M) (every decoder layer i),  E = enc_output.0(M)
// Every consumer of M starts with a Linear layer, and BatchNorm is an affine map per channel once its statistics are known,
// so  value_i = X_l^T (W_i diag(s) Wc)^T + (W_i t + b_i)  with  s = gamma * rstd, t = beta - mu * s  (tamtr_b200/fold.py has
// the algebra, including the statistics: mu = Wc mean(X), var = diag(Wc Cov(X) Wc^T)).  M is never materialised:
//
//   tamtr_tok_project   out[b, tok, :] = X_l[b, :, tok]^T W_fold^T + bias      K = C_l (128 / 256 / 512) instead of d = 512
//                       X is read where the backbone left it (NCHW: tokens contiguous = an MN-major A operand), the
//                       columns go to up to two bf16 tensors (value arena, ranking embedding) and an fp32 tail (scores)
//   tamtr_tok_reduce    D[m, n] = sum over all tokens of A[m; tok] * X_l[n; tok]  (+ row sums of A through a ones block)
//                       A = grad_value (weight gradient of the folded projection, replaces the dgrad + BatchNorm backward +
//                       conv wgrad chain) or A = X_l (second moments for the BatchNorm statistics)
//
// Both are warp-specialised: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, the rest = epilogue (TMEM -> registers ->
// swizzled staging tile -> TMA store / fp32 stores).  HBM-bound by construction.
#include <stdlib.h>

#include "tc_ptx.cuh"

namespace tamtr {

// ------------------------------------------------------------------------------------------------ shared helpers
__device__ __forceinline__ void tk_tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tk_tma_store_3d(const CUtensorMap *map, const void *src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tk_tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// K-major operand, SWIZZLE_128B: rows of 64 elements (128 B), 8-row atoms every 1024 B (SBO); a K step of 16 = +32 B
__device__ __forceinline__ uint64_t tk_desc_k(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major operand, SWIZZLE_128B, staged as 64-wide MN blocks of `64 K rows x 128 B`: 8 K rows = one 1024 B atom (SBO),
// the next 64 MN elements sit 8192 B further (LBO); a K step of 16 = +2048 B  (cute/atom/mma_traits_sm100.hpp:165-187,
// same construction as csrc/maxsig_tc.cu)
__device__ __forceinline__ uint64_t tk_desc_mn(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tk_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tk_tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tk_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t tk_pack2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

constexpr int kTkThreads = 64 + 256;         // producer warp, MMA warp, 8 epilogue warps
constexpr int kTkSlot = 128 * 64 * 2;        // one 128 x 64 bf16 operand block = 16 KB

// Shared-memory descriptors split into 32-bit halves: only the low word (start address >> 4 | LBO << 16) changes between
// MMAs, so the issue loop adds small constants to it.  The issuing warp walks its loops with uniform control flow and
// elects one lane per instruction group: with a single divergent lane (`if (lane == 0)`) every tcgen05 / TMA instruction
// costs an ELECT + a handful of R2UR moves, ~40 instructions per MMA -- which bounded the first version of these
// kernels (2 500 clk per K block whatever was loaded or multiplied: profiles/tokgemm_r2_notes.txt).
constexpr uint32_t kTkDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);      // SBO = 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t tk_lo_k(uint32_t addr) { return (addr >> 4) | (1u << 16); }
__device__ __forceinline__ uint32_t tk_lo_mn(uint32_t addr) { return (addr >> 4) | ((8192u >> 4) << 16); }
__device__ __forceinline__ uint64_t tk_desc(uint32_t lo) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(kTkDescHi));
    return d;
}
__device__ __forceinline__ bool tk_elect() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ================================================================================================ forward projection
// One CTA owns MT * 128 consecutive tokens of one image: their X columns (all K blocks) stay resident in shared memory
// while the CTA walks the N_all output columns 128 at a time, streaming W_fold through a ring (W_fold is a few hundred
// KB: L2 hits).  Two TMEM accumulator stages (MT x 128 columns each) overlap the MMAs of step n+1 with the epilogue of
// step n.  Shared memory: 11 operand slots of 16 KB (A: MT * C / 64 of them, the rest is the W ring) + two staging tiles.
constexpr int kPjSlots = 13;                 // 16 KB slots: A (MT * C / 64), the W ring, 2 or 4 staging tiles
constexpr int kPjMaxWStages = 8;

struct PjBars {
    uint64_t a_full, a_empty, w_full[kPjMaxWStages], w_empty[kPjMaxWStages], acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};

struct PjGeom {
    int B, HW, C, n_kb;            // level geometry; n_kb = C / 64
    int w_stages;                  // depth of the W ring = kPjSlots - MT * n_kb - 2 * stg_depth (capped)
    int stg_depth;                 // staging tiles per epilogue group (2 when the A block leaves room)
    int n_rot;                     // the first n_rot column steps (the stored value columns) are walked in a per-CTA rotated
                                   // order, so that the CTAs do not all pull the same W tile out of L2 at the same time
    int pairs_per_img, n_pairs;    // CTA tiles of MT * 128 tokens
    int N0, N1, NT, n_steps;       // bf16 columns of out0 / out1, fp32 tail columns, ceil((N0 + N1 + NT) / 128)
    long raw_row, raw_img;         // strides (elements) of the fp32 tail tensor
    // ranking mode (rank_mode = 1): the N1 columns are the ranking embedding E and the tail its class scores; neither is
    // stored -- the epilogue reduces them to one score per token (see tok_project_kernel)
    int rank_mode, nc;
    int defer;                     // ranking mode, NT == 16: the tail values wait in registers and the score is finished at the
                                   // end of the tile, so that ALL column steps can be walked in the rotated order
    int zero_fill;                 // also write zeros over the same (token, value column) range of a second tensor (map_z)
    long rank_img;
    float eps, inv_d;
};

template <int MT>
__global__ void __launch_bounds__(kTkThreads, 1)
tok_project_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                   const __grid_constant__ CUtensorMap map_o0, const __grid_constant__ CUtensorMap map_o1,
                   const __grid_constant__ CUtensorMap map_z, const float *__restrict__ bias, float *__restrict__ raw,
                   const uint8_t *__restrict__ valid,
                   const float *__restrict__ rconst, float *__restrict__ rank_out, const PjGeom g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sA = base;                                         // [MT][n_kb] slots of 16 KB
    uint8_t *sW = base + (size_t)MT * g.n_kb * kTkSlot;         // [w_stages] 16 KB
    uint8_t *sO = base + (size_t)(kPjSlots - 2 * g.stg_depth) * kTkSlot;     // staging tiles: [group][stg_depth]
    uint8_t *sEnd = base + (size_t)kPjSlots * kTkSlot;
    PjBars &bars = *reinterpret_cast<PjBars *>(sEnd);
    float *s_rc = reinterpret_cast<float *>(sEnd + 256);                   // ranking constants: 2 + 3 * NT floats
    float2 *s_part = reinterpret_cast<float2 *>(s_rc + 2 + 3 * 64);        // MT == 1: partial row moments of the other group
    uint8_t *s_zero = sEnd + 4096;                                         // 32 rows x 128 B of zeros (zero_fill)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (g.rank_mode)
        for (int i = threadIdx.x; i < 2 + 3 * g.NT; i += kTkThreads) s_rc[i] = __ldg(rconst + i);
    if (g.zero_fill) {
        for (int i = threadIdx.x; i < 4096 / 16; i += kTkThreads) reinterpret_cast<uint4 *>(s_zero)[i] = make_uint4(0u, 0u, 0u, 0u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        mbar_init(&bars.a_full, 1);
        mbar_init(&bars.a_empty, 1);
        for (int s = 0; s < kPjMaxWStages; ++s) { mbar_init(&bars.w_full[s], 1); mbar_init(&bars.w_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&bars.acc_full[a], 1); mbar_init(&bars.acc_empty[a], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars.tmem_base)),
                     "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = bars.tmem_base;
    constexpr int tile_tok = MT * 128;
    const int n0 = g.n_rot > 0 ? (int)(blockIdx.x % (unsigned)g.n_rot) : 0;

    if (warp == 0) {
        // ===== TMA producer (whole warp walks the loops, one elected lane issues)
        uint32_t ws = 0, wph = 0;
        int it = 0;
        for (int p = blockIdx.x; p < g.n_pairs; p += gridDim.x, ++it) {
            const int b = p / g.pairs_per_img, tok0 = (p - b * g.pairs_per_img) * tile_tok;
            mbar_wait(&bars.a_empty, (it & 1) ^ 1);
            if (tk_elect()) {
                // 64-token boxes that start inside the image; the others are skipped (their accumulator rows are never stored)
                int boxes = 0;
#pragma unroll
                for (int h = 0; h < 2 * MT; ++h) boxes += (tok0 + 64 * h < g.HW) ? 1 : 0;
                mbar_expect_tx(&bars.a_full, (uint32_t)(boxes * g.n_kb * 8192));
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
                    for (int kb = 0; kb < g.n_kb; ++kb)
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int t = tok0 + mt * 128 + h * 64;
                            if (t < g.HW)
                                tk_tma_load_3d(sA + (size_t)(mt * g.n_kb + kb) * kTkSlot + h * 8192, &map_x, &bars.a_full, t,
                                               kb * 64, b);
                        }
            }
            __syncwarp();
            for (int ns = 0; ns < g.n_steps; ++ns) {
                const int n = ns < g.n_rot ? (ns + n0 < g.n_rot ? ns + n0 : ns + n0 - g.n_rot) : ns;
                for (int kb = 0; kb < g.n_kb; ++kb) {
                    mbar_wait(&bars.w_empty[ws], wph ^ 1);
                    if (tk_elect()) {
                        mbar_expect_tx(&bars.w_full[ws], (uint32_t)kTkSlot);
                        tma_load_2d(sW + (size_t)ws * kTkSlot, &map_w, &bars.w_full[ws], kb * 64, n * 128);
                    }
                    __syncwarp();
                    if (++ws == (uint32_t)g.w_stages) { ws = 0; wph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: D[128 tokens, 128 columns] += A (MN-major) x W (K-major), K = 16 per instruction
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a_lo0 = tk_lo_mn(smem_u32(sA)), w_lo0 = tk_lo_k(smem_u32(sW));
        const uint32_t a_mt = (uint32_t)g.n_kb * (kTkSlot >> 4);          // descriptor units between the token blocks
        uint32_t ws = 0, wph = 0, ai = 0;
        int it = 0;
        for (int p = blockIdx.x; p < g.n_pairs; p += gridDim.x, ++it) {
            mbar_wait(&bars.a_full, it & 1);
            for (int n = 0; n < g.n_steps; ++n, ++ai) {
                const uint32_t as = ai & 1;
                mbar_wait(&bars.acc_empty[as], ((ai >> 1) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d0 = tmem_base + as * 256;
                for (int kb = 0; kb < g.n_kb; ++kb) {
                    mbar_wait(&bars.w_full[ws], wph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (tk_elect()) {
                        const uint32_t w_lo = w_lo0 + ws * (kTkSlot >> 4), a_lo = a_lo0 + (uint32_t)kb * (kTkSlot >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t bd = tk_desc(w_lo + 2 * k);
#pragma unroll
                            for (int mt = 0; mt < MT; ++mt)
                                umma_f16(d0 + mt * 128, tk_desc(a_lo + mt * a_mt + k * (2048 >> 4)), bd, idesc,
                                         (kb | k) ? 1u : 0u);
                        }
                        umma_commit(&bars.w_empty[ws]);
                    }
                    __syncwarp();
                    if (++ws == (uint32_t)g.w_stages) { ws = 0; wph ^= 1; }
                }
                if (tk_elect()) umma_commit(&bars.acc_full[as]);
                __syncwarp();
            }
            if (tk_elect()) umma_commit(&bars.a_empty);
            __syncwarp();
        }
    } else {
        // ===== epilogue: two groups of four warps; warp % 4 = TMEM lane quarter.  MT == 2: group = token block, both
        // 64-column chunks of the step; MT == 1: group = chunk.  Per chunk: TMEM -> + bias -> bf16 -> the token's 128-byte row
        // of the group's swizzled staging tile -> one TMA store (the map clips rows past the image's last token).
        const int quarter = warp & 3, grp = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;                       // TMEM lane = token within the 128-block
        uint8_t *stage0 = sO + (size_t)grp * g.stg_depth * kTkSlot;
        uint32_t sc = 0;                                           // stores issued by this group
        const uint32_t row_off = (uint32_t)row * 128, sw = (uint32_t)(row & 7);
        const bool leader = (warp - 2) % 4 == 0 && lane == 0;      // issues the group's TMA stores
        const int mt = MT == 2 ? grp : 0;
        const int chunk0 = MT == 2 ? 0 : grp, n_chunks = MT == 2 ? 2 : 1;
        const int Nmain = g.N0 + g.N1, Nall = Nmain + g.NT;
        // ranking mode: where the tail sits (NT <= 64: one chunk), and -- MT == 1, where two threads share a token's columns --
        // which group finishes the token
        const int n_tail = Nmain >> 7, tail_grp = (Nmain >> 6) & 1;
        uint32_t ai = 0;
        for (int p = blockIdx.x; p < g.n_pairs; p += gridDim.x) {
            const int b = p / g.pairs_per_img, tok0 = (p - b * g.pairs_per_img) * tile_tok + mt * 128;
            float s1 = 0.f, s2 = 0.f;                                        // sum / sum of squares of the token's E row
            float tv[16];                                                    // deferred mode: the token's 16 tail values
#pragma unroll
            for (int j = 0; j < 16; ++j) tv[j] = 0.f;
            for (int ns = 0; ns < g.n_steps; ++ns, ++ai) {
                const int n = ns < g.n_rot ? (ns + n0 < g.n_rot ? ns + n0 : ns + n0 - g.n_rot) : ns;
                const uint32_t as = ai & 1;
                mbar_wait(&bars.acc_full[as], (ai >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (as * 2 + mt) * 128;
                for (int ci = 0; ci < n_chunks; ++ci) {
                    const int chunk = chunk0 + ci, col0 = n * 128 + chunk * 64;
                    if (col0 >= Nall || tok0 >= g.HW) continue;              // uniform over the group
                    if (col0 >= Nmain) {
                        const int tok = tok0 + row, toff = col0 - Nmain;
                        uint32_t v[32];
                        if (g.rank_mode && g.defer) {
                            tk_tmem_ld32(taddr + chunk * 64, v);
#pragma unroll
                            for (int j = 0; j < 16; ++j) tv[j] = __uint_as_float(v[j]) + __ldg(bias + Nmain + j);
                            continue;
                        }
                        if (g.rank_mode) {
                            // ===== ranking score of the token (head.py:1229-1237): LayerNorm statistics of E + enc bias from
                            // (sum E, sum E^2, E . enc_bias), folded with the class scores of the tail columns
                            if (MT == 1) {
                                asm volatile("bar.sync 3, 256;" ::: "memory");      // the other group's partial moments
                                const float2 o = s_part[row];
                                s1 += o.x; s2 += o.y;
                            }
                            const bool ok = tok < g.HW && __ldg(valid + (tok < g.HW ? tok : 0)) != 0;
                            float best = -INFINITY, dot = 0.f, mean = 0.f, rstd = 0.f;
                            // the dot product sits in the LAST tail column: read its 32-column piece first
                            const int last = g.NT - 1;
                            tk_tmem_ld32(taddr + chunk * 64 + (last & ~31), v);
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j == (last & 31)) dot = __uint_as_float(v[j]) + __ldg(bias + Nmain + last);
                            if (!ok) { s1 = 0.f; s2 = 0.f; dot = 0.f; }
                            mean = (s1 + s_rc[0]) * g.inv_d;
                            rstd = rsqrtf(fmaxf((s2 + 2.f * dot + s_rc[1]) * g.inv_d - mean * mean, 0.f) + g.eps);
                            const float *bw = s_rc + 2, *sw_ = bw + g.NT, *ck = sw_ + g.NT;
                            int loaded = last >> 5;                          // which 32-column piece v holds
                            for (int h = 0; h * 32 < g.nc; ++h) {
                                if (h != loaded) { tk_tmem_ld32(taddr + chunk * 64 + h * 32, v); loaded = h; }
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    const int k = h * 32 + j;
                                    if (k < g.nc) {
                                        const float r = ok ? __uint_as_float(v[j]) + __ldg(bias + Nmain + k) : 0.f;
                                        best = fmaxf(best, rstd * (r + bw[k] - mean * sw_[k]) + ck[k]);
                                    }
                                }
                            }
                            if (tok < g.HW) rank_out[(size_t)b * g.rank_img + tok] = best;
                            if (MT == 1) asm volatile("bar.sync 3, 256;" ::: "memory");     // partials consumed
                            (void)toff;
                            continue;
                        }
                        // fp32 tail stored as it is: direct stores of the chunk's share of the NT columns
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int cb = toff + h * 32;
                            if (cb >= g.NT) break;                                  // uniform
                            tk_tmem_ld32(taddr + chunk * 64 + h * 32, v);
                            if (tok < g.HW) {
                                float *dst = raw + (size_t)b * g.raw_img + (size_t)tok * g.raw_row + cb;
#pragma unroll
                                for (int j = 0; j < 32; j += 4) {
                                    if (cb + j < g.NT) {
                                        const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bias + Nmain + cb + j));
                                        *reinterpret_cast<float4 *>(dst + j) =
                                            make_float4(__uint_as_float(v[j]) + b4.x, __uint_as_float(v[j + 1]) + b4.y,
                                                        __uint_as_float(v[j + 2]) + b4.z, __uint_as_float(v[j + 3]) + b4.w);
                                    }
                                }
                            }
                        }
                        continue;
                    }
                    if (g.rank_mode && col0 >= g.N0) {
                        // ===== ranking embedding: only its row moments are needed
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            uint32_t v[32];
                            tk_tmem_ld32(taddr + chunk * 64 + h * 32, v);
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bias + col0 + h * 32 + j));
                                const float e0 = __uint_as_float(v[j]) + b4.x, e1 = __uint_as_float(v[j + 1]) + b4.y;
                                const float e2 = __uint_as_float(v[j + 2]) + b4.z, e3 = __uint_as_float(v[j + 3]) + b4.w;
                                s1 += (e0 + e1) + (e2 + e3);
                                s2 = fmaf(e0, e0, fmaf(e1, e1, fmaf(e2, e2, fmaf(e3, e3, s2))));
                            }
                        }
                        continue;
                    }
                    // both halves of the chunk leave TMEM together while the group makes sure its staging tile is free
                    uint32_t v0[32], v1[32];
                    tk_tmem_ld32_nowait(taddr + chunk * 64, v0);
                    tk_tmem_ld32_nowait(taddr + chunk * 64 + 32, v1);
                    uint8_t *stage = stage0 + (size_t)(g.stg_depth == 2 ? (sc & 1) : 0) * kTkSlot;
                    ++sc;
                    if (leader) {       // the store that last read this tile (stg_depth stores ago) has finished reading
                        if (g.stg_depth == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
                    tk_tmem_wait_ld();
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t(&v)[32] = h == 0 ? v0 : v1;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 b0 = __ldg(reinterpret_cast<const float4 *>(bias + col0 + h * 32 + q * 8));
                            const float4 b1 = __ldg(reinterpret_cast<const float4 *>(bias + col0 + h * 32 + q * 8 + 4));
                            const uint32_t o0 = tk_pack2(__uint_as_float(v[q * 8 + 0]) + b0.x, __uint_as_float(v[q * 8 + 1]) + b0.y);
                            const uint32_t o1 = tk_pack2(__uint_as_float(v[q * 8 + 2]) + b0.z, __uint_as_float(v[q * 8 + 3]) + b0.w);
                            const uint32_t o2 = tk_pack2(__uint_as_float(v[q * 8 + 4]) + b1.x, __uint_as_float(v[q * 8 + 5]) + b1.y);
                            const uint32_t o3 = tk_pack2(__uint_as_float(v[q * 8 + 6]) + b1.z, __uint_as_float(v[q * 8 + 7]) + b1.w);
                            const uint32_t piece = (uint32_t)(h * 4 + q) ^ sw;
                            *reinterpret_cast<uint4 *>(stage + row_off + piece * 16) = make_uint4(o0, o1, o2, o3);
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
                    if (leader) {       // one bulk group per chunk: the data tile (+ four 32-row zero tiles over the same range)
                        if (col0 < g.N0) {
                            tk_tma_store_3d(&map_o0, stage, col0, tok0, b);
                            if (g.zero_fill) {
#pragma unroll
                                for (int r4 = 0; r4 < 4; ++r4)
                                    if (tok0 + 32 * r4 < g.HW) tk_tma_store_3d(&map_z, s_zero, col0, tok0 + 32 * r4, b);
                            }
                        } else {
                            tk_tma_store_3d(&map_o1, stage, col0 - g.N0, tok0, b);
                        }
                        tk_tma_commit();
                    }
                }
                if (MT == 1 && g.rank_mode && !g.defer && n == n_tail && grp != tail_grp) {
                    // this group's share of the token's moments -> the group that owns the tail chunk
                    s_part[row] = make_float2(s1, s2);
                    asm volatile("bar.sync 3, 256;" ::: "memory");
                    asm volatile("bar.sync 3, 256;" ::: "memory");
                }
                // every TMEM read of this accumulator stage by this warp is complete (tcgen05.wait::ld in tk_tmem_ld32)
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars.acc_empty[as]);
            }
            if (g.rank_mode && g.defer) {
                // ===== ranking score of the tile's tokens, after all of their column steps (same arithmetic as above)
                if (MT == 1) {
                    if (grp != tail_grp) s_part[row] = make_float2(s1, s2);
                    asm volatile("bar.sync 3, 256;" ::: "memory");
                    if (grp == tail_grp) { const float2 o = s_part[row]; s1 += o.x; s2 += o.y; }
                }
                const int tok = tok0 + row;
                if ((MT == 2 || grp == tail_grp) && tok < g.HW) {
                    const bool ok = __ldg(valid + tok) != 0;
                    const float dot = ok ? tv[15] : 0.f;
                    if (!ok) { s1 = 0.f; s2 = 0.f; }
                    const float mean = (s1 + s_rc[0]) * g.inv_d;
                    const float rstd = rsqrtf(fmaxf((s2 + 2.f * dot + s_rc[1]) * g.inv_d - mean * mean, 0.f) + g.eps);
                    const float *bw = s_rc + 2, *sw_ = bw + 16, *ck = sw_ + 16;
                    float best = -INFINITY;
#pragma unroll
                    for (int k = 0; k < 15; ++k)
                        if (k < g.nc) best = fmaxf(best, rstd * ((ok ? tv[k] : 0.f) + bw[k] - mean * sw_[k]) + ck[k]);
                    rank_out[(size_t)b * g.rank_img + tok] = best;
                }
                if (MT == 1) asm volatile("bar.sync 3, 256;" ::: "memory");
            }
        }
        if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

// ================================================================================================ reduction over tokens
// D[m, n] = sum_{b, tok} A[m; b, tok] * X[b, n, tok],   rs[m] = sum_{b, tok} A[m; b, tok]
// CTA (blockIdx.x = m group * n chunks + n chunk, blockIdx.y = split): MT row blocks of 128 x NB <= 256 columns, over the
// split's share of the (image, 64-token block) sequence; partial results are written per split (deterministic; the caller
// adds the splits).  The row sums come from one more MMA per K step against a block of ones (N = 16).
struct RdBars {
    uint64_t full[4], empty[4], acc_full;
    uint32_t tmem_base;
};

struct RdGeom {
    int B, HW, C, M;               // X [B, C, HW]; A has M rows
    int MT, NB, n_cchunks;         // row blocks per CTA, columns per CTA, C / NB
    int kb_per_img, total_kb, splits, stages;
    int a_mn;                      // 1: A given token-major [B, HW, M] (MN-major operand); 0: channel-major [B, M, HW]
    int tmem_cols;
};

constexpr uint32_t kRdOnesStep = 32;   // TMEM columns between the row-sum accumulators (32-column aligned: at 16 the sums
                                       // of the first row block came out nondeterministic)

template <int MT, bool AMN>
__global__ void __launch_bounds__(64 + 128, 1)
tok_reduce_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_x,
                  float *__restrict__ part_d, float *__restrict__ part_rs, const RdGeom g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int a_bytes = MT * kTkSlot;
    const int b_bytes = g.NB * 128, stage_bytes = a_bytes + b_bytes;
    uint8_t *ones = base + (size_t)g.stages * stage_bytes;                 // 16 rows x 128 B of bf16 1.0
    RdBars &bars = *reinterpret_cast<RdBars *>(ones + 2048);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mg = blockIdx.x / g.n_cchunks, cc = blockIdx.x - mg * g.n_cchunks;
    const int m0 = mg * MT * 128, c0 = cc * g.NB;
    const int split = blockIdx.y;
    const int kb0 = (int)((long)g.total_kb * split / g.splits), kb1 = (int)((long)g.total_kb * (split + 1) / g.splits);

    if (threadIdx.x == 0) {
        for (int s = 0; s < 4; ++s) { mbar_init(&bars.full[s], 1); mbar_init(&bars.empty[s], 1); }
        mbar_init(&bars.acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars.tmem_base)),
                     "r"(g.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = threadIdx.x; i < 2048 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(ones)[i] = 0x3F803F80u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = bars.tmem_base;
    const uint32_t ones_col = (uint32_t)(MT * g.NB);             // TMEM columns of the row-sum accumulators

    if (warp == 0) {
        // ===== TMA producer
        int b = kb0 / g.kb_per_img, tb = kb0 - b * g.kb_per_img;
        uint32_t s = 0, ph = 0;
        int boxes = 0;                                             // A boxes that start inside the M rows
        if (AMN) { for (int h = 0; h < 2 * MT; ++h) boxes += (m0 + 64 * h < g.M) ? 1 : 0; }
        else { for (int mt = 0; mt < MT; ++mt) boxes += (m0 + 128 * mt < g.M) ? 2 : 0; }
        const uint32_t tx = (uint32_t)(boxes * 8192 + b_bytes);
        for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&bars.empty[s], ph ^ 1);
            if (tk_elect()) {
                uint8_t *a = base + (size_t)s * stage_bytes;
                const int t = tb * 64;
                mbar_expect_tx(&bars.full[s], tx);
                if (AMN) {
#pragma unroll
                    for (int h = 0; h < 2 * MT; ++h)
                        if (m0 + 64 * h < g.M) tk_tma_load_3d(a + (size_t)h * 8192, &map_a, &bars.full[s], m0 + 64 * h, t, b);
                } else {
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
                        if (m0 + 128 * mt < g.M)
                            tk_tma_load_3d(a + (size_t)mt * kTkSlot, &map_a, &bars.full[s], t, m0 + 128 * mt, b);
                }
                tk_tma_load_3d(a + a_bytes, &map_x, &bars.full[s], t, c0, b);
            }
            __syncwarp();
            if (++tb == g.kb_per_img) { tb = 0; ++b; }
            if (++s == (uint32_t)g.stages) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        // ===== MMA issuer
        constexpr uint32_t major = AMN ? (1u << 15) : 0u;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | major | ((uint32_t)(g.NB >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | major | ((16u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a_lo0 = AMN ? tk_lo_mn(smem_u32(base)) : tk_lo_k(smem_u32(base));
        const uint32_t b_lo0 = tk_lo_k(smem_u32(base + a_bytes)), o_lo = tk_lo_k(smem_u32(ones));
        constexpr uint32_t a_k = AMN ? (2048u >> 4) : 2u;         // descriptor units per K step of 16
        const uint32_t stage_units = (uint32_t)stage_bytes >> 4;
        uint32_t s = 0, ph = 0, first = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&bars.full[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tk_elect()) {
                const uint32_t a_lo = a_lo0 + s * stage_units, b_lo = b_lo0 + s * stage_units;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t bd = tk_desc(b_lo + 2 * k), od = tk_desc(o_lo + 2 * k);
                    const uint32_t acc = (first | k) ? 1u : 0u;
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const uint64_t ad = tk_desc(a_lo + mt * (kTkSlot >> 4) + k * a_k);
                        umma_f16(tmem_base + mt * g.NB, ad, bd, idesc, acc);
                        umma_f16(tmem_base + ones_col + mt * kRdOnesStep, ad, od, idesc1, acc);
                    }
                }
                umma_commit(&bars.empty[s]);
            }
            __syncwarp();
            first = 1;
            if (++s == (uint32_t)g.stages) { s = 0; ph ^= 1; }
        }
        if (tk_elect()) umma_commit(&bars.acc_full);
        __syncwarp();
    } else {
        // ===== epilogue: 4 warps, lane = row of the 128-block; fp32 partials of this split
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        if (kb1 > kb0) {
            mbar_wait(&bars.acc_full, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const int m = m0 + mt * 128 + row;
            float *drow = part_d + ((size_t)split * g.M + (m < g.M ? m : 0)) * g.C + c0;
            for (int j0 = 0; j0 < g.NB; j0 += 32) {
                uint32_t v[32];
                if (kb1 > kb0) tk_tmem_ld32(trow + mt * g.NB + j0, v);
                else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0u;
                }
                if (m < g.M) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4 *>(drow + j0 + j) =
                            make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                        __uint_as_float(v[j + 3]));
                }
            }
            if (cc == 0) {
                uint32_t r[4] = {0u, 0u, 0u, 0u};
                if (kb1 > kb0) {
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                                 : "r"(trow + ones_col + mt * kRdOnesStep));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                }
                if (m < g.M) part_rs[(size_t)split * g.M + m] = __uint_as_float(r[0]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(g.tmem_cols));
    }
}

}  // namespace tamtr

using namespace tamtr;

static int tk_encode(CUtensorMap *map, const void *ptr, int rank, const cuuint64_t *dims, const cuuint64_t *strides,
                     const cuuint32_t *box, const char *what) {
    EncodeTiledFn encode = get_encode();
    TAMTR_CHECK_ARG(encode != nullptr, TAMTR_E_NODEVICE, "%s: cuTensorMapEncodeTiled unavailable", what);
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult cr = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void *>(ptr), dims, strides,
                               box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TAMTR_CHECK_ARG(cr == CUDA_SUCCESS, TAMTR_E_BADARG, "%s: cuTensorMapEncodeTiled failed (%d)", what, (int)cr);
    return 0;
}

static bool tk_attr_once(const void *fn, bool *flags) {
    int dev_id = 0;
    if (cudaGetDevice(&dev_id) != cudaSuccess) return false;
    if (dev_id < 0 || dev_id >= 64 || !flags[dev_id]) {
        if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return false;
        if (dev_id >= 0 && dev_id < 64) flags[dev_id] = true;
    }
    return true;
}

extern "C" int tamtr_tok_project_supported(int B, int C, int HW, int N0, int N1, int NT) {
    if (B <= 0 || C <= 0 || HW <= 0 || N0 <= 0 || N1 < 0 || NT < 0) return 0;
    if (C % 64 != 0 || C > 512 || HW % 8 != 0) return 0;               // K blocks of 64; TMA row stride of X = HW * 2 bytes
    if (N0 % 64 != 0 || N1 % 64 != 0 || NT % 4 != 0 || NT > 1024) return 0;
    return 1;
}

struct PjRank {                  // ranking mode (tamtr_tok_project_rank)
    float *rank;
    long rank_img;
    const uint8_t *valid;
    const float *consts;
    int nc;
    float eps;
};

static int tok_project_impl(const void *x_bf16, const void *w_bf16, const float *bias, void *out0, void *zero0, long out0_row,
                            long out0_img, void *out1, long out1_row, long out1_img, float *raw, long raw_row,
                            long raw_img, const PjRank *rk, int B, int C, int HW, int N0, int N1, int NT, void *stream) {
    TAMTR_CHECK_ARG(x_bf16 && w_bf16 && bias && out0, TAMTR_E_BADARG, "tok_project: null pointer");
    TAMTR_CHECK_ARG(tamtr_tok_project_supported(B, C, HW, N0, N1, NT), TAMTR_E_UNSUPPORTED,
                    "tok_project: unsupported problem (B=%d C=%d HW=%d N0=%d N1=%d NT=%d): need C %% 64 == 0, C <= 512, "
                    "HW %% 8 == 0, N0 %% 64 == 0, N1 %% 64 == 0, NT %% 4 == 0, NT <= 1024", B, C, HW, N0, N1, NT);
    TAMTR_CHECK_ARG(rk != nullptr || ((N1 == 0 || out1 != nullptr) && (NT == 0 || raw != nullptr)), TAMTR_E_BADARG,
                    "tok_project: missing output tensor");
    if (rk != nullptr) {
        TAMTR_CHECK_ARG(rk->rank && rk->valid && rk->consts, TAMTR_E_BADARG, "tok_project_rank: null pointer");
        TAMTR_CHECK_ARG(N1 > 0 && NT >= 16 && NT <= 64 && rk->nc > 0 && rk->nc < NT, TAMTR_E_UNSUPPORTED,
                        "tok_project_rank: need N1 > 0 and nc < NT <= 64 (nc = %d, NT = %d)", rk->nc, NT);
    }
    TAMTR_CHECK_ARG((((uintptr_t)x_bf16 | (uintptr_t)w_bf16 | (uintptr_t)bias | (uintptr_t)out0 | (uintptr_t)out1 |
                      (uintptr_t)raw) & 15) == 0, TAMTR_E_BADARG, "tok_project: pointers must be 16-byte aligned");
    TAMTR_CHECK_ARG(out0_row % 8 == 0 && out0_img % 8 == 0 && out1_row % 8 == 0 && out1_img % 8 == 0 && raw_row % 4 == 0 &&
                    raw_img % 4 == 0, TAMTR_E_BADARG, "tok_project: output strides must keep 16-byte alignment");
    const int Nall = N0 + N1 + NT;
    TAMTR_CHECK_ARG(((uintptr_t)zero0 & 15) == 0, TAMTR_E_BADARG, "tok_project: zero0 must be 16-byte aligned");
    CUtensorMap map_x, map_w, map_o0, map_o1, map_z;
    {
        const cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)C, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)HW * 2, (cuuint64_t)C * HW * 2};
        const cuuint32_t box[3] = {64, 64, 1};
        const int rc = tk_encode(&map_x, x_bf16, 3, dims, strides, box, "tok_project(x)");
        if (rc) return rc;
    }
    {
        const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)Nall};
        const cuuint64_t strides[1] = {(cuuint64_t)C * 2};
        const cuuint32_t box[2] = {64, 128};
        const int rc = tk_encode(&map_w, w_bf16, 2, dims, strides, box, "tok_project(w)");
        if (rc) return rc;
    }
    for (int which = 0; which < 2; ++which) {
        const bool live = which == 0 || (N1 > 0 && rk == nullptr);
        const cuuint64_t dims[3] = {(cuuint64_t)(which == 0 ? N0 : (live ? N1 : N0)), (cuuint64_t)HW, (cuuint64_t)B};
        const long row = which == 0 || !live ? out0_row : out1_row, img = which == 0 || !live ? out0_img : out1_img;
        const cuuint64_t strides[2] = {(cuuint64_t)row * 2, (cuuint64_t)img * 2};
        const cuuint32_t box[3] = {64, 128, 1};
        const int rc = tk_encode(which == 0 ? &map_o0 : &map_o1, which == 0 || !live ? out0 : out1, 3, dims, strides, box,
                                 "tok_project(out)");
        if (rc) return rc;
    }
    {       // the tensor zero-filled alongside out0 (same layout; plain 32-row boxes of the shared zero tile)
        EncodeTiledFn encode = get_encode();
        const cuuint64_t dims[3] = {(cuuint64_t)N0, (cuuint64_t)HW, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)out0_row * 2, (cuuint64_t)out0_img * 2};
        const cuuint32_t box[3] = {64, 32, 1}, estr[3] = {1, 1, 1};
        const CUresult cr = encode(&map_z, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, zero0 != nullptr ? zero0 : out0, dims, strides, box,
                                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TAMTR_CHECK_ARG(cr == CUDA_SUCCESS, TAMTR_E_BADARG, "tok_project: zero-fill tensor map failed (%d)", (int)cr);
    }
    PjGeom g;
    g.zero_fill = zero0 != nullptr ? 1 : 0;
    g.B = B; g.HW = HW; g.C = C; g.n_kb = C / 64;
    int MT = (2 * g.n_kb <= 8) ? 2 : 1;
    if (HW <= 128) MT = 1;
    if (g.n_kb > 2 && !getenv("TAMTR_TOK_MT2")) MT = 1;      // C > 128: a deeper W ring beats sharing W between two token blocks
    g.stg_depth = (kPjSlots - MT * g.n_kb - 4 >= 4) ? 2 : 1;
    if (getenv("TAMTR_TOK_STG1")) g.stg_depth = 1;      // (experiment switch)
    g.w_stages = kPjSlots - MT * g.n_kb - 2 * g.stg_depth;
    if (g.w_stages > kPjMaxWStages) g.w_stages = kPjMaxWStages;
    g.pairs_per_img = (HW + MT * 128 - 1) / (MT * 128);
    g.n_pairs = B * g.pairs_per_img;
    g.N0 = N0; g.N1 = N1; g.NT = NT;
    g.n_steps = (Nall + 127) / 128;
    g.raw_row = raw_row; g.raw_img = raw_img;
    g.n_rot = (N0 % 128 == 0 && !getenv("TAMTR_TOK_NOROT")) ? N0 / 128 : 0;
    g.defer = 0;
    g.rank_mode = rk != nullptr ? 1 : 0;
    g.nc = rk != nullptr ? rk->nc : 0;
    g.rank_img = rk != nullptr ? rk->rank_img : 0;
    g.eps = rk != nullptr ? rk->eps : 0.f;
    g.inv_d = N1 > 0 ? 1.0f / (float)N1 : 0.f;
    if (rk != nullptr && NT == 16 && g.n_rot > 0 && (N0 + N1) % 128 == 0 && !getenv("TAMTR_TOK_NODEFER")) {
        g.defer = 1;
        g.n_rot = g.n_steps;                // every column step, the E and tail steps included
    }
    const uint8_t *valid = rk != nullptr ? rk->valid : nullptr;
    const float *rconst = rk != nullptr ? rk->consts : nullptr;
    float *rank_out = rk != nullptr ? rk->rank : nullptr;
    static bool attr1[64] = {false}, attr2[64] = {false};
    TAMTR_CHECK_ARG(tk_attr_once((const void *)tok_project_kernel<1>, attr1) &&
                    tk_attr_once((const void *)tok_project_kernel<2>, attr2), TAMTR_E_NODEVICE,
                    "tok_project: cannot raise the dynamic shared memory limit");
    const size_t smem = (size_t)kPjSlots * kTkSlot + 4096 + 4096 + 1024;      // slots, barriers + ranking scratch, zero tile, alignment
    const int n_sm = ::tamtr::sm_count();
    const int grid = g.n_pairs < n_sm ? g.n_pairs : n_sm;
    cudaStream_t st = (cudaStream_t)stream;
    {
        KernelTimer timer(K_TOK_PROJECT, st);
        if (MT == 2)
            tok_project_kernel<2><<<grid, kTkThreads, smem, st>>>(map_x, map_w, map_o0, map_o1, map_z, bias, raw, valid, rconst, rank_out, g);
        else
            tok_project_kernel<1><<<grid, kTkThreads, smem, st>>>(map_x, map_w, map_o0, map_o1, map_z, bias, raw, valid, rconst, rank_out, g);
    }
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_tok_project(const void *x_bf16, const void *w_bf16, const float *bias, void *out0, void *zero0,
                                 long out0_row, long out0_img, void *out1, long out1_row, long out1_img, float *raw,
                                 long raw_row, long raw_img, int B, int C, int HW, int N0, int N1, int NT, void *stream) {
    return tok_project_impl(x_bf16, w_bf16, bias, out0, zero0, out0_row, out0_img, out1, out1_row, out1_img, raw, raw_row,
                            raw_img, nullptr, B, C, HW, N0, N1, NT, stream);
}

extern "C" int tamtr_tok_project_rank(const void *x_bf16, const void *w_bf16, const float *bias, void *out0, void *zero0,
                                      long out0_row, long out0_img, float *rank, long rank_img, const uint8_t *valid,
                                      const float *rank_consts, int nc, float eps, int B, int C, int HW, int N0, int N1,
                                      int NT, void *stream) {
    const PjRank rk = {rank, rank_img, valid, rank_consts, nc, eps};
    return tok_project_impl(x_bf16, w_bf16, bias, out0, zero0, out0_row, out0_img, nullptr, 0, 0, nullptr, 0, 0, &rk, B, C, HW, N0,
                            N1, NT, stream);
}

static int rd_geometry(RdGeom &g, int B, int C, int HW, int M, int a_mn) {
    g.B = B; g.HW = HW; g.C = C; g.M = M; g.a_mn = a_mn;
    g.NB = C <= 256 ? C : 256;
    if (C % g.NB != 0) return TAMTR_E_UNSUPPORTED;
    g.n_cchunks = C / g.NB;
    g.MT = (g.NB <= 128 && M > 128) ? 2 : 1;
    g.kb_per_img = (HW + 63) / 64;
    g.total_kb = B * g.kb_per_img;
    const int groups = ((M + g.MT * 128 - 1) / (g.MT * 128)) * g.n_cchunks;
    int splits = ::tamtr::sm_count() / groups;
    if (splits < 1) splits = 1;
    if (splits > g.total_kb) splits = g.total_kb;
    g.splits = splits;
    const int stage_bytes = g.MT * kTkSlot + g.NB * 128;
    g.stages = (200 * 1024) / stage_bytes;
    if (g.stages > 4) g.stages = 4;
    const int cols = g.MT * g.NB + g.MT * (int)kRdOnesStep;
    g.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
    return 0;
}

extern "C" int tamtr_tok_reduce_supported(int B, int C, int HW, int M, int a_token_major) {
    if (B <= 0 || C <= 0 || HW <= 0 || M <= 0) return 0;
    if (C % 64 != 0 || HW % 8 != 0) return 0;
    if (C > 256 && C % 256 != 0) return 0;
    if (a_token_major && M % 8 != 0) return 0;
    return 1;
}

/* number of per-split partial results tamtr_tok_reduce writes for this problem */
extern "C" int tamtr_tok_reduce_splits(int B, int C, int HW, int M, int a_token_major) {
    if (!tamtr_tok_reduce_supported(B, C, HW, M, a_token_major)) return 0;
    RdGeom g;
    if (rd_geometry(g, B, C, HW, M, a_token_major)) return 0;
    return g.splits;
}

extern "C" int tamtr_tok_reduce(const void *a_bf16, long a_row, long a_img, int a_token_major, const void *x_bf16,
                                float *part_d, float *part_rs, int B, int C, int HW, int M, void *stream) {
    TAMTR_CHECK_ARG(a_bf16 && x_bf16 && part_d && part_rs, TAMTR_E_BADARG, "tok_reduce: null pointer");
    TAMTR_CHECK_ARG(tamtr_tok_reduce_supported(B, C, HW, M, a_token_major), TAMTR_E_UNSUPPORTED,
                    "tok_reduce: unsupported problem (B=%d C=%d HW=%d M=%d): need C %% 64 == 0 (C > 256: C %% 256 == 0), "
                    "HW %% 8 == 0", B, C, HW, M);
    TAMTR_CHECK_ARG((((uintptr_t)a_bf16 | (uintptr_t)x_bf16 | (uintptr_t)part_d) & 15) == 0 && a_row % 8 == 0 && a_img % 8 == 0,
                    TAMTR_E_BADARG, "tok_reduce: pointers / strides must keep 16-byte alignment");
    RdGeom g;
    TAMTR_CHECK_ARG(rd_geometry(g, B, C, HW, M, a_token_major ? 1 : 0) == 0, TAMTR_E_UNSUPPORTED, "tok_reduce: bad geometry");
    CUtensorMap map_a, map_x;
    if (a_token_major) {       // A [B, HW, M] (row = token, stride a_row): box = 64 columns (MN) x 64 tokens (K)
        const cuuint64_t dims[3] = {(cuuint64_t)M, (cuuint64_t)HW, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)a_row * 2, (cuuint64_t)a_img * 2};
        const cuuint32_t box[3] = {64, 64, 1};
        const int rc = tk_encode(&map_a, a_bf16, 3, dims, strides, box, "tok_reduce(a)");
        if (rc) return rc;
    } else {                   // A [B, M, HW] (row = channel, stride a_row): box = 64 tokens (K) x 128 rows
        const cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)M, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)a_row * 2, (cuuint64_t)a_img * 2};
        const cuuint32_t box[3] = {64, 128, 1};
        const int rc = tk_encode(&map_a, a_bf16, 3, dims, strides, box, "tok_reduce(a)");
        if (rc) return rc;
    }
    {
        const cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)C, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)HW * 2, (cuuint64_t)C * HW * 2};
        const cuuint32_t box[3] = {64, (cuuint32_t)g.NB, 1};
        const int rc = tk_encode(&map_x, x_bf16, 3, dims, strides, box, "tok_reduce(x)");
        if (rc) return rc;
    }
    static bool attr[4][64] = {{false}};
    TAMTR_CHECK_ARG(tk_attr_once((const void *)tok_reduce_kernel<1, false>, attr[0]) &&
                    tk_attr_once((const void *)tok_reduce_kernel<2, false>, attr[1]) &&
                    tk_attr_once((const void *)tok_reduce_kernel<1, true>, attr[2]) &&
                    tk_attr_once((const void *)tok_reduce_kernel<2, true>, attr[3]), TAMTR_E_NODEVICE,
                    "tok_reduce: cannot raise the dynamic shared memory limit");
    const size_t smem = (size_t)g.stages * (g.MT * kTkSlot + g.NB * 128) + 2048 + sizeof(RdBars) + 1024;
    const int groups = ((M + g.MT * 128 - 1) / (g.MT * 128)) * g.n_cchunks;
    cudaStream_t st = (cudaStream_t)stream;
    {
        KernelTimer timer(K_TOK_REDUCE, st);
        const dim3 grid(groups, g.splits);
        if (g.MT == 2 && a_token_major) tok_reduce_kernel<2, true><<<grid, 64 + 128, smem, st>>>(map_a, map_x, part_d, part_rs, g);
        else if (g.MT == 2) tok_reduce_kernel<2, false><<<grid, 64 + 128, smem, st>>>(map_a, map_x, part_d, part_rs, g);
        else if (a_token_major) tok_reduce_kernel<1, true><<<grid, 64 + 128, smem, st>>>(map_a, map_x, part_d, part_rs, g);
        else tok_reduce_kernel<1, false><<<grid, 64 + 128, smem, st>>>(map_a, map_x, part_d, part_rs, g);
    }
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}
