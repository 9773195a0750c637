// Query self-attention of the decoder layers (ultralytics/nn/modules/transformer.py:544-548: nn.MultiheadAttention on
// q = k = embed + pos, v = embed, with the contrastive-denoising attention mask of models/utils/ops.py:273-284), after the
// packed input projection: softmax(q k^T / sqrt(Dh) + mask) v per (image, head), forward and backward.
//
// The problem is tiny -- 300..700 queries, Dh = 64 (32 for the 256-wide heads), 128 (image, head) pairs: 3 GFLOP forward --
// and the library's fused attention spends 86 us forward / 70 us backward per layer on it once a mask is present (17 us
// without one), plus a dozen layout copies per step around it (head split / merge, slice gradients).  Here q and k are read
// straight out of the packed projection [B, L, 2 d] and v out of [B, L, d], the output is written as [B, L, d], and the
// gradients come back in those same layouts.
//
// Forward (flash-attention style, one CTA per (64 queries, head, image), a warp per 16 query rows): S = Q K^T on
// mma.sync.m16n8k16 (bf16 in, fp32 accumulate) over key tiles of 64, mask and bounds applied to the accumulators, online
// softmax in registers (row statistics over the 4 lanes of a quad), the probabilities re-packed from the accumulator layout
// into A fragments for P V.  K / V tiles are double-buffered with cp.async; rows are padded by 16 bytes so that ldmatrix
// reads 8 rows from 8 different bank groups.
// Backward in two kernels, no atomics: (1) per query tile: recompute P from the saved log-sum-exp, dP = dO V^T,
// dS = P (dP - delta), dQ = dS K; P and dS leave as bf16 tiles in a scratch buffer [B, H, Lp, Lp] (2 x 26 MB at 300
// queries: 8 us of HBM time); (2) per key tile: dV = P^T dO, dK = dS^T Q, reading those tiles back through ldmatrix.trans.
// (The 5th-generation tensor cores are not used here on purpose: with 64-query tiles and 3 GFLOP the kernel is latency-,
// not throughput-bound, and a tcgen05 pipeline -- TMEM allocation, TMA descriptors, MN-major operands for the two
// transposed products -- has more fixed cost per CTA than this whole kernel runs for.)
#include "common.cuh"

namespace tamtr {

constexpr int kSaQ = 64;          // queries per CTA (4 warps x 16)
constexpr int kSaK = 64;          // keys per tile
constexpr int kSaThreads = 128;
constexpr float kSaLog2e = 1.4426950408889634f;

__device__ __forceinline__ void sa_ldsm4(uint32_t (&r)[4], const void *p) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void sa_ldsm4_t(uint32_t (&r)[4], const void *p) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
// D (16x8 fp32) += A (16x16 bf16, row) * B (16x8 bf16, col)
__device__ __forceinline__ void sa_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t sa_pack(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&v);
}
__device__ __forceinline__ float sa_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void sa_cp16(void *dst, const void *src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void sa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void sa_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// rows [r0, r0 + ROWS) x DH columns starting at column c0 of a [*, ld] bf16 matrix -> tile[ROWS][DH + 8] (zero rows past n_rows)
template <int ROWS, int DH>
__device__ __forceinline__ void sa_stage(__nv_bfloat16 (*tile)[DH + 8], const __nv_bfloat16 *__restrict__ src, size_t ld, int r0,
                                         int n_rows, int c0) {
    constexpr int CPR = DH / 8;
    for (int i = threadIdx.x; i < ROWS * CPR; i += kSaThreads) {
        const int r = i / CPR, c = i % CPR, row = r0 + r;
        sa_cp16(&tile[r][8 * c], src + (size_t)min(row, n_rows - 1) * ld + c0 + 8 * c, row < n_rows);
    }
}

// S accumulators of one warp (16 query rows x 64 keys) = Q (A fragments, in registers) x K^T (tile in shared memory)
template <int DH>
__device__ __forceinline__ void sa_qkT(float (&s)[8][4], const uint32_t (&qa)[DH / 16][4], const __nv_bfloat16 (*kt)[DH + 8], int lane) {
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.0f;
    const int row = (lane & 7) + ((lane >> 4) << 3), col = ((lane >> 3) & 1) << 3;
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks) {
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {                         // 16 keys per ldmatrix.x4: n-tiles 2 jp, 2 jp + 1
            uint32_t b[4];
            sa_ldsm4(b, &kt[16 * jp + row][16 * ks + col]);
            sa_mma(s[2 * jp], qa[ks], b[0], b[1]);
            sa_mma(s[2 * jp + 1], qa[ks], b[2], b[3]);
        }
    }
}

// acc (16 rows x DH) += P (A fragments built from the S-shaped accumulators `p`, 16 rows x 64 keys) x M (tile [64 keys][DH])
template <int DH>
__device__ __forceinline__ void sa_pm(float (&acc)[DH / 8][4], const float (&p)[8][4], const __nv_bfloat16 (*mt)[DH + 8], int lane) {
    const int row = (lane & 7) + (((lane >> 3) & 1) << 3), col = (lane >> 4) << 3;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {                             // 16 keys per step
        uint32_t a[4] = {sa_pack(p[2 * ks][0], p[2 * ks][1]), sa_pack(p[2 * ks][2], p[2 * ks][3]),
                         sa_pack(p[2 * ks + 1][0], p[2 * ks + 1][1]), sa_pack(p[2 * ks + 1][2], p[2 * ks + 1][3])};
#pragma unroll
        for (int np = 0; np < DH / 16; ++np) {                   // 16 output columns per ldmatrix.x4.trans
            uint32_t b[4];
            sa_ldsm4_t(b, &mt[16 * ks + row][16 * np + col]);
            sa_mma(acc[2 * np], a, b[0], b[1]);
            sa_mma(acc[2 * np + 1], a, b[2], b[3]);
        }
    }
}

// The mask arrives bit-packed (tamtr_self_attention_pack_mask): bits[q][t] holds the 64 keys of tile t, bit set = blocked, keys
// past L included -- two 8-byte loads per thread and tile instead of 32 byte loads.  Without a mask only the bounds remain.
__device__ __forceinline__ unsigned long long sa_tile_bits(const unsigned long long *__restrict__ bits, int n_tiles, int L, int q,
                                                           int t) {
    if (bits != nullptr) return q < L ? __ldg(bits + (size_t)q * n_tiles + t) : 0ull;
    const int live = L - t * kSaK;                               // keys of this tile that exist
    return live >= 64 ? 0ull : ~0ull << live;
}
// s <- s * scale2 (log2 units), -inf where blocked
__device__ __forceinline__ void sa_mask(float (&s)[8][4], unsigned long long w0, unsigned long long w1, int lane, float scale2) {
    const int sh = 2 * (lane & 3);
    const unsigned long long a = w0 >> sh, b = w1 >> sh;         // rows lane / 4 and lane / 4 + 8
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        s[j][0] = ((a >> (8 * j)) & 1ull) ? -INFINITY : s[j][0] * scale2;
        s[j][1] = ((a >> (8 * j + 1)) & 1ull) ? -INFINITY : s[j][1] * scale2;
        s[j][2] = ((b >> (8 * j)) & 1ull) ? -INFINITY : s[j][2] * scale2;
        s[j][3] = ((b >> (8 * j + 1)) & 1ull) ? -INFINITY : s[j][3] * scale2;
    }
}

__global__ void selfattn_pack_mask_kernel(const uint8_t *__restrict__ blocked, unsigned long long *__restrict__ bits, int L,
                                          int n_tiles) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L * n_tiles) return;
    const int q = i / n_tiles, t = i % n_tiles;
    unsigned long long w = 0ull;
    for (int c = 0; c < 64; ++c) {
        const int k = t * kSaK + c;
        const bool dead = k >= L || blocked[(size_t)q * L + k] != 0;
        w |= (unsigned long long)dead << c;
    }
    bits[i] = w;
}

template <int DH>
__global__ void __launch_bounds__(kSaThreads)
selfattn_fwd_kernel(const __nv_bfloat16 *__restrict__ qk, const __nv_bfloat16 *__restrict__ v,
                    const unsigned long long *__restrict__ bits,
                    __nv_bfloat16 *__restrict__ o, float *__restrict__ lse2, int L, int H, float scale2) {
    __shared__ __align__(16) __nv_bfloat16 qs[kSaQ][DH + 8];
    __shared__ __align__(16) __nv_bfloat16 ks[2][kSaK][DH + 8];
    __shared__ __align__(16) __nv_bfloat16 vs[2][kSaK][DH + 8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q0 = blockIdx.x * kSaQ, h = blockIdx.y, b = blockIdx.z, d = H * DH;
    const __nv_bfloat16 *qkb = qk + (size_t)b * L * 2 * d, *vb = v + (size_t)b * L * d;
    sa_stage<kSaQ, DH>(qs, qkb, 2 * d, q0, L, h * DH);
    sa_stage<kSaK, DH>(ks[0], qkb, 2 * d, 0, L, d + h * DH);
    sa_stage<kSaK, DH>(vs[0], vb, d, 0, L, h * DH);
    sa_commit();
    const int n_tiles = (L + kSaK - 1) / kSaK;
    uint32_t qa[DH / 16][4];
    float acc[DH / 8][4];
#pragma unroll
    for (int j = 0; j < DH / 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f;
    float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.0f, 0.0f};     // rows lane / 4 and lane / 4 + 8 of the warp's 16
    const int q_lo = q0 + 16 * warp + (lane >> 2);
    for (int t = 0; t < n_tiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < n_tiles) {
            sa_stage<kSaK, DH>(ks[buf ^ 1], qkb, 2 * d, (t + 1) * kSaK, L, d + h * DH);
            sa_stage<kSaK, DH>(vs[buf ^ 1], vb, d, (t + 1) * kSaK, L, h * DH);
            sa_commit();
            sa_wait<1>();
        } else {
            sa_wait<0>();
        }
        __syncthreads();
        if (t == 0) {
#pragma unroll
            for (int ksx = 0; ksx < DH / 16; ++ksx)
                sa_ldsm4(qa[ksx], &qs[16 * warp + (lane & 7) + (((lane >> 3) & 1) << 3)][16 * ksx + ((lane >> 4) << 3)]);
        }
        float s[8][4];
        sa_qkT<DH>(s, qa, ks[buf], lane);
        sa_mask(s, sa_tile_bits(bits, n_tiles, L, q_lo, t), sa_tile_bits(bits, n_tiles, L, q_lo + 8, t), lane, scale2);
#pragma unroll
        for (int r = 0; r < 2; ++r) {                             // online softmax, rows r = 0 (e 0,1) and 1 (e 2,3)
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 8; ++j) mx = fmaxf(mx, fmaxf(s[j][2 * r], s[j][2 * r + 1]));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            const float m_new = fmaxf(m[r], mx);
            const float base = m_new == -INFINITY ? 0.0f : m_new;        // a row that has seen no key yet
            const float corr = sa_ex2(m[r] - base);
            float sum = 0.0f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s[j][2 * r] = sa_ex2(s[j][2 * r] - base);
                s[j][2 * r + 1] = sa_ex2(s[j][2 * r + 1] - base);
                sum += s[j][2 * r] + s[j][2 * r + 1];
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            l[r] = l[r] * corr + sum;
            m[r] = m_new;
#pragma unroll
            for (int j = 0; j < DH / 8; ++j) {
                acc[j][2 * r] *= corr;
                acc[j][2 * r + 1] *= corr;
            }
        }
        sa_pm<DH>(acc, s, vs[buf], lane);
        __syncthreads();                                           // this buffer is refilled by the next iteration's prefetch
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int q = q_lo + 8 * r;
        if (q >= L) continue;
        const float inv = l[r] > 0.0f ? 1.0f / l[r] : 0.0f;
        __nv_bfloat16 *dst = o + ((size_t)b * L + q) * d + h * DH + 2 * (lane & 3);
#pragma unroll
        for (int j = 0; j < DH / 8; ++j)
            *reinterpret_cast<uint32_t *>(dst + 8 * j) = sa_pack(acc[j][2 * r] * inv, acc[j][2 * r + 1] * inv);
        if ((lane & 3) == 0) lse2[((size_t)b * H + h) * L + q] = l[r] > 0.0f ? m[r] + log2f(l[r]) : INFINITY;
    }
}

// backward 1: per query tile.  dQ, and the P / dS tiles (bf16) for backward 2.
template <int DH> struct SaBwdQSmem {
    __nv_bfloat16 qs[kSaQ][DH + 8];
    __nv_bfloat16 gs[kSaQ][DH + 8];                              // dO
    __nv_bfloat16 ks[2][kSaK][DH + 8];
    __nv_bfloat16 vs[2][kSaK][DH + 8];
    __nv_bfloat16 st[kSaThreads / 32][16][kSaK + 8];             // per warp: a 16 x 64 tile on its way to the scratch buffer
    float delta[kSaQ];
};

// a warp's 16 x 64 accumulator tile -> bf16 rows of the scratch matrix (row pitch Lp), as 16-byte stores
__device__ __forceinline__ void sa_store_tile(__nv_bfloat16 (*st)[kSaK + 8], const float (&x)[8][4], __nv_bfloat16 *__restrict__ dst,
                                              size_t pitch, int lane) {
    const int r = lane >> 2, c = 2 * (lane & 3);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        *reinterpret_cast<uint32_t *>(&st[r][8 * j + c]) = sa_pack(x[j][0], x[j][1]);
        *reinterpret_cast<uint32_t *>(&st[r + 8][8 * j + c]) = sa_pack(x[j][2], x[j][3]);
    }
    __syncwarp();
#pragma unroll
    for (int i = lane; i < 16 * 8; i += 32) {                    // 16 rows x 8 chunks of 16 bytes
        const int rr = i >> 3, cc = i & 7;
        *reinterpret_cast<uint4 *>(dst + (size_t)rr * pitch + 8 * cc) = *reinterpret_cast<const uint4 *>(&st[rr][8 * cc]);
    }
    __syncwarp();
}

template <int DH>
__global__ void __launch_bounds__(kSaThreads)
selfattn_bwd_q_kernel(const __nv_bfloat16 *__restrict__ qk, const __nv_bfloat16 *__restrict__ v,
                      const unsigned long long *__restrict__ bits,
                      const __nv_bfloat16 *__restrict__ o, const __nv_bfloat16 *__restrict__ d_o, const float *__restrict__ lse2,
                      __nv_bfloat16 *__restrict__ d_qk, __nv_bfloat16 *__restrict__ p_out, __nv_bfloat16 *__restrict__ ds_out,
                      int L, int Lp, int H, float scale2, float scale) {
    extern __shared__ __align__(16) unsigned char sa_raw[];
    SaBwdQSmem<DH> &sm = *reinterpret_cast<SaBwdQSmem<DH> *>(sa_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q0 = blockIdx.x * kSaQ, h = blockIdx.y, b = blockIdx.z, d = H * DH;
    const __nv_bfloat16 *qkb = qk + (size_t)b * L * 2 * d, *vb = v + (size_t)b * L * d;
    const __nv_bfloat16 *gb = d_o + (size_t)b * L * d, *ob = o + (size_t)b * L * d;
    sa_stage<kSaQ, DH>(sm.qs, qkb, 2 * d, q0, L, h * DH);
    sa_stage<kSaQ, DH>(sm.gs, gb, d, q0, L, h * DH);
    sa_stage<kSaK, DH>(sm.ks[0], qkb, 2 * d, 0, L, d + h * DH);
    sa_stage<kSaK, DH>(sm.vs[0], vb, d, 0, L, h * DH);
    sa_commit();
    // delta[q] = <dO[q], O[q]> over the head's channels: two threads per row
    {
        const int r = threadIdx.x >> 1, half = threadIdx.x & 1, q = q0 + r;
        float acc = 0.0f;
        if (q < L) {
            const __nv_bfloat16 *po = ob + (size_t)q * d + h * DH + half * (DH / 2), *pg = gb + (size_t)q * d + h * DH + half * (DH / 2);
#pragma unroll
            for (int c = 0; c < DH / 2; c += 2) {
                const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(po + c));
                const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(pg + c));
                acc = fmaf(a.x, g.x, fmaf(a.y, g.y, acc));
            }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (half == 0) sm.delta[r] = acc;
    }
    const int n_tiles = (L + kSaK - 1) / kSaK;
    uint32_t qa[DH / 16][4], ga[DH / 16][4];
    float dq[DH / 8][4];
#pragma unroll
    for (int j = 0; j < DH / 8; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.0f;
    const int q_lo = q0 + 16 * warp + (lane >> 2);
    float ls[2], dl[2];
    __nv_bfloat16 *pb = p_out + ((size_t)b * H + h) * Lp * Lp + (size_t)(q0 + 16 * warp) * Lp;
    __nv_bfloat16 *dsb = ds_out + ((size_t)b * H + h) * Lp * Lp + (size_t)(q0 + 16 * warp) * Lp;
    for (int t = 0; t < n_tiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < n_tiles) {
            sa_stage<kSaK, DH>(sm.ks[buf ^ 1], qkb, 2 * d, (t + 1) * kSaK, L, d + h * DH);
            sa_stage<kSaK, DH>(sm.vs[buf ^ 1], vb, d, (t + 1) * kSaK, L, h * DH);
            sa_commit();
            sa_wait<1>();
        } else {
            sa_wait<0>();
        }
        __syncthreads();
        if (t == 0) {
            const int ar = 16 * warp + (lane & 7) + (((lane >> 3) & 1) << 3), ac = (lane >> 4) << 3;
#pragma unroll
            for (int ksx = 0; ksx < DH / 16; ++ksx) {
                sa_ldsm4(qa[ksx], &sm.qs[ar][16 * ksx + ac]);
                sa_ldsm4(ga[ksx], &sm.gs[ar][16 * ksx + ac]);
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int q = q_lo + 8 * r;
                ls[r] = q < L ? lse2[((size_t)b * H + h) * L + q] : INFINITY;
                dl[r] = sm.delta[16 * warp + (lane >> 2) + 8 * r];
            }
        }
        float s[8][4], dp[8][4];
        sa_qkT<DH>(s, qa, sm.ks[buf], lane);
        sa_mask(s, sa_tile_bits(bits, n_tiles, L, q_lo, t), sa_tile_bits(bits, n_tiles, L, q_lo + 8, t), lane, scale2);
        sa_qkT<DH>(dp, ga, sm.vs[buf], lane);                        // dP = dO V^T
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float p = sa_ex2(s[j][e] - ls[e >> 1]);            // (-inf - lse -> 0; lse = +inf for dead rows -> 0)
                s[j][e] = p;
                dp[j][e] = p * (dp[j][e] - dl[e >> 1]) * scale;         // dS with respect to q k^T, times the softmax scale
            }
        }
        // P and dS tiles for the key-side kernel (rows q0 + 16 warp .. + 15 < Lp, columns of tile t)
        sa_store_tile(sm.st[warp], s, pb + t * kSaK, Lp, lane);
        sa_store_tile(sm.st[warp], dp, dsb + t * kSaK, Lp, lane);
        sa_pm<DH>(dq, dp, sm.ks[buf], lane);                         // dQ += dS K
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int q = q_lo + 8 * r;
        if (q >= L) continue;
        __nv_bfloat16 *dst = d_qk + ((size_t)b * L + q) * 2 * d + h * DH + 2 * (lane & 3);
#pragma unroll
        for (int j = 0; j < DH / 8; ++j) *reinterpret_cast<uint32_t *>(dst + 8 * j) = sa_pack(dq[j][2 * r], dq[j][2 * r + 1]);
    }
}

// acc (16 keys x DH) += T^T (T = tile [64 queries][64 keys] in shared memory, the warp's 16 key columns) x M ([64 queries][DH])
template <int DH>
__device__ __forceinline__ void sa_tTm(float (&acc)[DH / 8][4], const __nv_bfloat16 (*tt)[kSaK + 8], int key0,
                                       const __nv_bfloat16 (*mt)[DH + 8], int lane) {
    const int arow = (lane & 7) + ((lane >> 4) << 3), acol = ((lane >> 3) & 1) << 3;      // A = T^T through ldmatrix.trans
    const int brow = (lane & 7) + (((lane >> 3) & 1) << 3), bcol = (lane >> 4) << 3;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {                             // 16 queries per step
        uint32_t a[4];
        sa_ldsm4_t(a, &tt[16 * ks + arow][key0 + acol]);
#pragma unroll
        for (int np = 0; np < DH / 16; ++np) {
            uint32_t b[4];
            sa_ldsm4_t(b, &mt[16 * ks + brow][16 * np + bcol]);
            sa_mma(acc[2 * np], a, b[0], b[1]);
            sa_mma(acc[2 * np + 1], a, b[2], b[3]);
        }
    }
}

// backward 2: per key tile.  dV = P^T dO, dK = dS^T Q.
template <int DH> struct SaBwdKvSmem {
    __nv_bfloat16 qs[2][kSaQ][DH + 8];
    __nv_bfloat16 gs[2][kSaQ][DH + 8];
    __nv_bfloat16 ps[2][kSaQ][kSaK + 8];
    __nv_bfloat16 dss[2][kSaQ][kSaK + 8];
};

template <int DH>
__global__ void __launch_bounds__(kSaThreads)
selfattn_bwd_kv_kernel(const __nv_bfloat16 *__restrict__ qk, const __nv_bfloat16 *__restrict__ d_o,
                       const __nv_bfloat16 *__restrict__ p_in, const __nv_bfloat16 *__restrict__ ds_in,
                       __nv_bfloat16 *__restrict__ d_qk, __nv_bfloat16 *__restrict__ d_v, int L, int Lp, int H) {
    extern __shared__ __align__(16) unsigned char sa_raw[];
    SaBwdKvSmem<DH> &sm = *reinterpret_cast<SaBwdKvSmem<DH> *>(sa_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k0 = blockIdx.x * kSaK, h = blockIdx.y, b = blockIdx.z, d = H * DH;
    const __nv_bfloat16 *qkb = qk + (size_t)b * L * 2 * d, *gb = d_o + (size_t)b * L * d;
    const __nv_bfloat16 *pb = p_in + ((size_t)b * H + h) * Lp * Lp, *dsb = ds_in + ((size_t)b * H + h) * Lp * Lp;
    float dv[DH / 8][4], dk[DH / 8][4];
#pragma unroll
    for (int j = 0; j < DH / 8; ++j) {
        dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.0f;
        dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.0f;
    }
    auto prefetch = [&](int buf, int q0) {
        sa_stage<kSaQ, DH>(sm.qs[buf], qkb, 2 * d, q0, L, h * DH);
        sa_stage<kSaQ, DH>(sm.gs[buf], gb, d, q0, L, h * DH);
        sa_stage<kSaQ, kSaK>(sm.ps[buf], pb, Lp, q0, Lp, k0);        // rows q0 .. q0 + 63 exist (Lp is a multiple of 64)
        sa_stage<kSaQ, kSaK>(sm.dss[buf], dsb, Lp, q0, Lp, k0);
        sa_commit();
    };
    const int n_tiles = (L + kSaQ - 1) / kSaQ;
    prefetch(0, 0);
    for (int t = 0; t < n_tiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < n_tiles) { prefetch(buf ^ 1, (t + 1) * kSaQ); sa_wait<1>(); } else { sa_wait<0>(); }
        __syncthreads();
        sa_tTm<DH>(dv, sm.ps[buf], 16 * warp, sm.gs[buf], lane);
        sa_tTm<DH>(dk, sm.dss[buf], 16 * warp, sm.qs[buf], lane);
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int k = k0 + 16 * warp + (lane >> 2) + 8 * r;
        if (k >= L) continue;
        __nv_bfloat16 *dvp = d_v + ((size_t)b * L + k) * d + h * DH + 2 * (lane & 3);
        __nv_bfloat16 *dkp = d_qk + ((size_t)b * L + k) * 2 * d + d + h * DH + 2 * (lane & 3);
#pragma unroll
        for (int j = 0; j < DH / 8; ++j) {
            *reinterpret_cast<uint32_t *>(dvp + 8 * j) = sa_pack(dv[j][2 * r], dv[j][2 * r + 1]);
            *reinterpret_cast<uint32_t *>(dkp + 8 * j) = sa_pack(dk[j][2 * r], dk[j][2 * r + 1]);
        }
    }
}

}  // namespace tamtr

using namespace tamtr;

extern "C" int tamtr_self_attention_supported(int L, int H, int Dh) {
    return (L > 0 && H > 0 && H <= 65535 && (Dh == 32 || Dh == 64)) ? 1 : 0;
}
extern "C" int tamtr_self_attention_padded_len(int L) { return L > 0 ? (L + kSaQ - 1) / kSaQ * kSaQ : 0; }

static int sa_check(const void *a, const void *b, const void *c, int Bn, int L, int H, int Dh) {
    TAMTR_CHECK_ARG(a && b && c, TAMTR_E_BADARG, "self_attention: null pointer");
    TAMTR_CHECK_ARG(Bn > 0 && Bn <= 65535 && tamtr_self_attention_supported(L, H, Dh), TAMTR_E_UNSUPPORTED,
                    "self_attention: B = %d, L = %d, H = %d, head dim = %d (32 or 64)", Bn, L, H, Dh);
    TAMTR_CHECK_ARG(((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)c)) & 15) == 0, TAMTR_E_UNSUPPORTED,
                    "self_attention: pointers must be 16-byte aligned");
    return 0;
}

extern "C" int tamtr_self_attention_mask_words(int L) { return L > 0 ? L * ((L + kSaK - 1) / kSaK) : 0; }

extern "C" int tamtr_self_attention_pack_mask(const uint8_t *blocked, unsigned long long *bits, int L, void *stream) {
    TAMTR_CHECK_ARG(blocked && bits && L > 0, TAMTR_E_BADARG, "self_attention_pack_mask: bad argument");
    const int n = tamtr_self_attention_mask_words(L);
    selfattn_pack_mask_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(blocked, bits, L, (L + kSaK - 1) / kSaK);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_self_attention_forward(const void *qk, const void *v, const unsigned long long *mask_bits, void *o, float *lse2,
                                            int Bn, int L, int H, int Dh, void *stream) {
    int rc = sa_check(qk, v, o, Bn, L, H, Dh);
    if (rc) return rc;
    TAMTR_CHECK_ARG(lse2 != nullptr, TAMTR_E_BADARG, "self_attention_forward: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid((L + kSaQ - 1) / kSaQ, H, Bn);
    const float scale2 = kSaLog2e / sqrtf((float)Dh);
    if (Dh == 64)
        selfattn_fwd_kernel<64><<<grid, kSaThreads, 0, st>>>((const __nv_bfloat16 *)qk, (const __nv_bfloat16 *)v, mask_bits,
                                                            (__nv_bfloat16 *)o, lse2, L, H, scale2);
    else
        selfattn_fwd_kernel<32><<<grid, kSaThreads, 0, st>>>((const __nv_bfloat16 *)qk, (const __nv_bfloat16 *)v, mask_bits,
                                                            (__nv_bfloat16 *)o, lse2, L, H, scale2);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int DH>
static cudaError_t sa_backward_launch(dim3 grid, cudaStream_t st, const void *qk, const void *v, const unsigned long long *bits,
                                      const void *o, const void *d_o, const float *lse2, void *d_qk, void *d_v, __nv_bfloat16 *p,
                                      __nv_bfloat16 *ds, int L, int Lp, int H, float scale2, float scale) {
    static bool done[64] = {false};                              // > 48 KB of dynamic shared memory: opt in per device
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !done[dev]) {
        e = cudaFuncSetAttribute(selfattn_bwd_q_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SaBwdQSmem<DH>));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(selfattn_bwd_kv_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SaBwdKvSmem<DH>));
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) done[dev] = true;
    }
    selfattn_bwd_q_kernel<DH><<<grid, kSaThreads, sizeof(SaBwdQSmem<DH>), st>>>(
        (const __nv_bfloat16 *)qk, (const __nv_bfloat16 *)v, bits, (const __nv_bfloat16 *)o, (const __nv_bfloat16 *)d_o, lse2,
        (__nv_bfloat16 *)d_qk, p, ds, L, Lp, H, scale2, scale);
    selfattn_bwd_kv_kernel<DH><<<grid, kSaThreads, sizeof(SaBwdKvSmem<DH>), st>>>(
        (const __nv_bfloat16 *)qk, (const __nv_bfloat16 *)d_o, p, ds, (__nv_bfloat16 *)d_qk, (__nv_bfloat16 *)d_v, L, Lp, H);
    return cudaGetLastError();
}

extern "C" int tamtr_self_attention_backward(const void *qk, const void *v, const unsigned long long *mask_bits, const void *o,
                                             const void *d_o, const float *lse2, void *d_qk, void *d_v, void *scratch, int Bn,
                                             int L, int H, int Dh, void *stream) {
    int rc = sa_check(qk, v, o, Bn, L, H, Dh);
    if (rc) return rc;
    rc = sa_check(d_o, d_qk, d_v, Bn, L, H, Dh);
    if (rc) return rc;
    TAMTR_CHECK_ARG(lse2 && scratch && (((uintptr_t)scratch) & 15) == 0, TAMTR_E_BADARG, "self_attention_backward: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int Lp = tamtr_self_attention_padded_len(L);
    __nv_bfloat16 *p = (__nv_bfloat16 *)scratch, *ds = p + (size_t)Bn * H * Lp * Lp;
    const dim3 grid((L + kSaQ - 1) / kSaQ, H, Bn);
    const float scale = 1.0f / sqrtf((float)Dh), scale2 = kSaLog2e * scale;
    TAMTR_CUDA_OK(Dh == 64 ? sa_backward_launch<64>(grid, st, qk, v, mask_bits, o, d_o, lse2, d_qk, d_v, p, ds, L, Lp, H, scale2, scale)
                           : sa_backward_launch<32>(grid, st, qk, v, mask_bits, o, d_o, lse2, d_qk, d_v, p, ds, L, Lp, H, scale2, scale));
    count_launch(2);
    return 0;
}
