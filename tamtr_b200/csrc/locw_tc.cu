// Kernel 3, fused: the two query projections of MSDeformAttn AND their epilogue in one tcgen05 kernel (bf16 operands).
//
// Replaces /root/reference ultralytics/nn/modules/transformer.py:278-293 for 16-bit activations:
//   raw  = query [M, C] x W^T [C, 3*H*S]          sampling_offsets and attention_weights Linears (they share their input)
//   attn = softmax over the S = L*P logits of each (query, head);  loc = ref_xy + (off + b) / P * ref_wh * 0.5
// The unfused path (library GEMM -> [M, 3*H*S] fp32 in HBM -> locw_fwd_kernel, plus two torch.cat of the weights and
// biases) moves the GEMM result through memory twice.  Here one CTA owns 128 queries and Hc of the H heads (the heads of a
// query tile are split over up to 4 CTAs so that the launch covers the SMs: a CTA's TMA ingest bounds the main loop):
//   warp 0      streams the query tile and the heads' weight rows through a shared-memory ring with TMA (SWIZZLE_128B,
//               K-major both); sampling_offsets.weight and attention_weights.weight are read where they are (two maps)
//   warp 1      issues tcgen05.mma (M = 128, N = 3*Hc*S in chunks <= 256, K = 16) into TMEM
//   warps 2-17  epilogue, 4 warps per TMEM lane quarter, one head at a time each: tcgen05.ld of the query's row (one TMEM
//               lane = one query), bias, softmax, location arithmetic with the reference's op order and per-op rounding,
//               16-byte stores of loc / attn
// `raw` is written only when the caller needs it (reference-box gradients).
#include "tc_ptx.cuh"

namespace tamtr {

constexpr int kLwTileM = 128;
constexpr int kLwKB = 64;                         // K elements per stage = one 128-byte swizzled row
constexpr int kLwABytes = kLwTileM * kLwKB * 2;   // 16 KB
constexpr int kLwMaxN = 512;                      // TMEM columns
constexpr int kLwEpiGroups = 4;                  // epilogue warps per TMEM lane quarter: each takes every 4th head
constexpr int kLwThreads = 64 + kLwEpiGroups * 128;
constexpr int kLwSmemBudget = 200 * 1024;

struct LwBars {
    uint64_t full[8], empty[8], acc_full;
    uint32_t tmem_base;
};

__device__ __forceinline__ uint64_t lw_desc_k128(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, float *v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_x4(uint32_t taddr, float *v) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

// S = L*P samples per (query, head): 12 (3 levels x 4 points) or 16 (4 x 4)
template <int S>
__global__ void __launch_bounds__(kLwThreads, 1)
locw_tc_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_woff,
                   const __grid_constant__ CUtensorMap map_watt, const void *__restrict__ b_off,
                   const void *__restrict__ b_att, int bias_bf16, const float *__restrict__ ref, float *__restrict__ loc,
                   float *__restrict__ attn, float *__restrict__ raw, int M, int H, int Hc, int P, int n_kb, int n_chunks,
                   int chunk, int stages, int tmem_cols) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // blockIdx.y = which Hc of the H heads this CTA projects: the 128-query tile is shared by H / Hc CTAs so that the kernel
    // spreads over more SMs (a CTA's TMA ingest, ~50 B/clk, bounds the main loop: 416 KB per CTA with all 8 heads)
    const int NT = 3 * Hc * S, NTall = 3 * H * S, h0 = blockIdx.y * Hc;
    const int b_bytes = NT * kLwKB * 2;                      // this CTA's rows of the weight matrix, one K block
    const int stage_bytes = kLwABytes + b_bytes;
    LwBars &bars = *reinterpret_cast<LwBars *>(base + (size_t)stages * stage_bytes);
    float *s_bias = reinterpret_cast<float *>(base + (size_t)stages * stage_bytes + sizeof(LwBars));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kLwTileM;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&bars.full[s], 1); mbar_init(&bars.empty[s], 1); }
        mbar_init(&bars.acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars.tmem_base)),
                     "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = threadIdx.x; i < NTall; i += kLwThreads) {      // s_bias = sampling_offsets.bias ++ attention_weights.bias
        const int n_off = 2 * H * S;
        const void *src = i < n_off ? b_off : b_att;
        const int j = i < n_off ? i : i - n_off;
        s_bias[i] = bias_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(src)[j])
                              : __ldg(reinterpret_cast<const float *>(src) + j);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = bars.tmem_base;

    if (warp == 0) {
        // ===== TMA producer: per K block the 128-query tile and every row of W_cat
        if (lane == 0) {
            for (int kb = 0; kb < n_kb; ++kb) {
                const int s = kb % stages;
                mbar_wait(&bars.empty[s], ((kb / stages) & 1) ^ 1);
                mbar_expect_tx(&bars.full[s], (uint32_t)stage_bytes);
                uint8_t *a = base + (size_t)s * stage_bytes;
                tma_load_2d(a, &map_q, &bars.full[s], kb * kLwKB, m0);
                // rows of W_cat for heads [h0, h0 + Hc): 2*Hc*S offset rows (two boxes of Hc*S rows), then Hc*S logit rows
                uint8_t *bt = a + kLwABytes;
                const int box = Hc * S;
                tma_load_2d(bt, &map_woff, &bars.full[s], kb * kLwKB, h0 * 2 * S);
                tma_load_2d(bt + (size_t)box * 128, &map_woff, &bars.full[s], kb * kLwKB, h0 * 2 * S + box);
                tma_load_2d(bt + (size_t)2 * box * 128, &map_watt, &bars.full[s], kb * kLwKB, h0 * S);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer
        if (lane == 0) {
            // instruction descriptor: D = f32, A = B = bf16, both K-major, N = chunk, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(chunk >> 3) << 17) | ((128u >> 4) << 24);
            for (int kb = 0; kb < n_kb; ++kb) {
                const int s = kb % stages;
                mbar_wait(&bars.full[s], (kb / stages) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(base + (size_t)s * stage_bytes);
                const uint32_t b_addr = a_addr + kLwABytes;
#pragma unroll
                for (int k = 0; k < kLwKB / 16; ++k) {
                    const uint64_t ad = lw_desc_k128(a_addr + k * 32);
                    for (int j = 0; j < n_chunks; ++j)
                        umma_f16(tmem_base + j * chunk, ad, lw_desc_k128(b_addr + j * chunk * 128 + k * 32), idesc,
                                 (kb | k) ? 1u : 0u);
                }
                umma_commit(&bars.empty[s]);
            }
            umma_commit(&bars.acc_full);
        }
    } else {
        // ===== epilogue: 16 warps; warp % 4 = the TMEM lane quarter it may read, lane = query row, (warp - 2) / 4 = which
        // heads it takes.  (With four epilogue warps -- one per scheduler -- the kernel took 25 us, 21 of them in this
        // dependent exp / divide chain at one instruction per ~10 cycles: ncu, profiles/locw_tc_r1_ncu_summary.csv.)
        const int quarter = warp & 3, hg = (warp - 2) >> 2;
        const int m = m0 + quarter * 32 + lane;
        const bool live = m < M;
        float4 rb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) rb = __ldg(reinterpret_cast<const float4 *>(ref) + m);      // (cx, cy, w, h): RL = 1, RD = 4
        const float fP = (float)P;
        const bool p_pow2 = (P & (P - 1)) == 0;
        const float inv_p = 1.0f / fP;                                       // exact when P is a power of two
        mbar_wait(&bars.acc_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
        for (int hl = hg; hl < Hc; hl += kLwEpiGroups) {
            const int h = h0 + hl;
            float off[2 * S], lg[S];
#pragma unroll
            for (int i = 0; i < 2 * S / 8; ++i) tmem_ld_x8(trow + hl * 2 * S + 8 * i, off + 8 * i);
#pragma unroll
            for (int i = 0; i < S / 4; ++i) tmem_ld_x4(trow + 2 * Hc * S + hl * S + 4 * i, lg + 4 * i);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (live) {
                if (raw != nullptr) {
                    float4 *ro = reinterpret_cast<float4 *>(raw + (size_t)m * NTall + h * 2 * S);
#pragma unroll
                    for (int i = 0; i < 2 * S / 4; ++i) ro[i] = make_float4(off[4 * i], off[4 * i + 1], off[4 * i + 2], off[4 * i + 3]);
                    float4 *rl = reinterpret_cast<float4 *>(raw + (size_t)m * NTall + 2 * H * S + h * S);
#pragma unroll
                    for (int i = 0; i < S / 4; ++i) rl[i] = make_float4(lg[4 * i], lg[4 * i + 1], lg[4 * i + 2], lg[4 * i + 3]);
                }
                const float *boff = s_bias + h * 2 * S, *blg = s_bias + 2 * H * S + h * S;
                float zmax = -INFINITY;
#pragma unroll
                for (int s = 0; s < S; ++s) { lg[s] = lg[s] + blg[s]; zmax = fmaxf(zmax, lg[s]); }
                float sum = 0.0f;
#pragma unroll
                for (int s = 0; s < S; ++s) { lg[s] = expf(lg[s] - zmax); sum += lg[s]; }
#pragma unroll
                for (int s = 0; s < S; ++s) lg[s] = __fdiv_rn(lg[s], sum);
                float4 *ao = reinterpret_cast<float4 *>(attn + ((size_t)m * H + h) * S);
#pragma unroll
                for (int i = 0; i < S / 4; ++i) ao[i] = make_float4(lg[4 * i], lg[4 * i + 1], lg[4 * i + 2], lg[4 * i + 3]);
                // loc = xy + (off + b) / P * wh * 0.5, each op rounded on its own (transformer.py:292-293)
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    const float ox = __fadd_rn(off[2 * s], boff[2 * s]);
                    const float oy = __fadd_rn(off[2 * s + 1], boff[2 * s + 1]);
                    // x / P == x * (1/P) bit for bit when P is a power of two (TAM-TR: 4)
                    const float qx = p_pow2 ? __fmul_rn(ox, inv_p) : __fdiv_rn(ox, fP);
                    const float qy = p_pow2 ? __fmul_rn(oy, inv_p) : __fdiv_rn(oy, fP);
                    off[2 * s] = __fadd_rn(rb.x, __fmul_rn(__fmul_rn(qx, rb.z), 0.5f));
                    off[2 * s + 1] = __fadd_rn(rb.y, __fmul_rn(__fmul_rn(qy, rb.w), 0.5f));
                }
                float4 *lo = reinterpret_cast<float4 *>(loc + ((size_t)m * H + h) * 2 * S);
#pragma unroll
                for (int i = 0; i < 2 * S / 4; ++i) lo[i] = make_float4(off[4 * i], off[4 * i + 1], off[4 * i + 2], off[4 * i + 3]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
    }
}

}  // namespace tamtr

using namespace tamtr;

// heads per CTA: the largest split of the H heads (4, 2, 1 ways) whose 3*Hc*S columns form valid UMMA shapes and whose grid
// still fits one wave of SMs; 0 = unsupported
static int lw_heads_per_cta(int M, int H, int S) {
    const int tiles = (M + kLwTileM - 1) / kLwTileM;
    const int n_sm = ::tamtr::sm_count();
    int best = 0;
    for (int split = 1; split <= 4; split *= 2) {
        if (H % split) break;
        const int Hc = H / split, NT = 3 * Hc * S;
        if (NT > kLwMaxN || NT % 16 != 0 || Hc * S > 256) continue;
        const int n_chunks = (NT + 255) / 256;
        if (NT % n_chunks != 0 || (NT / n_chunks) % 16 != 0) continue;
        if (kLwABytes + NT * kLwKB * 2 > kLwSmemBudget / 2) continue;       // at least a double buffer
        if (best != 0 && tiles * split > n_sm) break;
        best = Hc;
    }
    return best;
}

// 1 when tamtr_locw_tc_forward supports the problem (the caller otherwise uses a library GEMM + tamtr_locw_forward)
extern "C" int tamtr_locw_tc_supported(int M, int C, int H, int L, int P, int RL, int RD) {
    const int S = L * P;
    if (M <= 0 || C <= 0 || H <= 0 || (S != 12 && S != 16) || RL != 1 || RD != 4) return 0;
    if (C % kLwKB != 0) return 0;
    return lw_heads_per_cta(M, H, S) > 0 ? 1 : 0;
}

extern "C" int tamtr_locw_tc_forward(const void *query_bf16, const void *w_off_bf16, const void *w_attn_bf16,
                                     const void *b_off, const void *b_attn, int bias_dtype, const float *ref, float *loc,
                                     float *attn, float *raw, int M, int C, int H, int L, int P, int RL, int RD,
                                     void *stream) {
    TAMTR_CHECK_ARG(query_bf16 && w_off_bf16 && w_attn_bf16 && b_off && b_attn && ref && loc && attn, TAMTR_E_BADARG,
                    "locw_tc_forward: null pointer");
    TAMTR_CHECK_ARG(bias_dtype == TAMTR_F32 || bias_dtype == TAMTR_BF16, TAMTR_E_UNSUPPORTED, "locw_tc_forward: bias dtype %d",
                    bias_dtype);
    TAMTR_CHECK_ARG(tamtr_locw_tc_supported(M, C, H, L, P, RL, RD), TAMTR_E_UNSUPPORTED,
                    "locw_tc_forward: unsupported problem (M=%d C=%d H=%d L=%d P=%d RL=%d RD=%d); need L*P in {12,16}, "
                    "C %% 64 == 0, 3*H*L*P <= 512, 4-d reference boxes shared by the levels", M, C, H, L, P, RL, RD);
    TAMTR_CHECK_ARG((((uintptr_t)query_bf16 | (uintptr_t)w_off_bf16 | (uintptr_t)w_attn_bf16 | (uintptr_t)ref |
                      (uintptr_t)loc | (uintptr_t)attn | (uintptr_t)raw) & 15) == 0, TAMTR_E_BADARG,
                    "locw_tc_forward: pointers must be 16-byte aligned");
    EncodeTiledFn encode = get_encode();
    TAMTR_CHECK_ARG(encode != nullptr, TAMTR_E_NODEVICE, "locw_tc_forward: cuTensorMapEncodeTiled unavailable");
    const int S = L * P, NTall = 3 * H * S;
    const int Hc = lw_heads_per_cta(M, H, S), NT = 3 * Hc * S;
    const int n_chunks = (NT + 255) / 256, chunk = NT / n_chunks;
    CUtensorMap map_q, map_woff, map_watt;
    const cuuint32_t estr[2] = {1, 1};
    {
        const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)M};
        const cuuint64_t strides[1] = {(cuuint64_t)C * 2};
        const cuuint32_t box[2] = {(cuuint32_t)kLwKB, (cuuint32_t)kLwTileM};
        const CUresult cr = encode(&map_q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(query_bf16), dims, strides,
                                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TAMTR_CHECK_ARG(cr == CUDA_SUCCESS, TAMTR_E_BADARG, "locw_tc_forward: query tensor map failed (%d)", (int)cr);
    }
    for (int which = 0; which < 2; ++which) {      // sampling_offsets.weight [2*H*S, C], attention_weights.weight [H*S, C]
        const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)((which == 0 ? 2 : 1) * H * S)};
        const cuuint64_t strides[1] = {(cuuint64_t)C * 2};
        const cuuint32_t box[2] = {(cuuint32_t)kLwKB, (cuuint32_t)(Hc * S)};
        const CUresult cr = encode(which == 0 ? &map_woff : &map_watt, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                                   const_cast<void *>(which == 0 ? w_off_bf16 : w_attn_bf16), dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TAMTR_CHECK_ARG(cr == CUDA_SUCCESS, TAMTR_E_BADARG, "locw_tc_forward: weight tensor map failed (%d)", (int)cr);
    }
    const int stage_bytes = kLwABytes + NT * kLwKB * 2;
    int stages = kLwSmemBudget / stage_bytes;
    if (stages > 8) stages = 8;
    const int n_kb = C / kLwKB;
    if (stages > n_kb) stages = n_kb;
    const int tmem_cols = NT <= 32 ? 32 : NT <= 64 ? 64 : NT <= 128 ? 128 : NT <= 256 ? 256 : 512;
    const size_t smem = (size_t)stages * stage_bytes + sizeof(LwBars) + (size_t)NTall * sizeof(float) + 1024;
    static bool attr_set[64] = {false};          // cudaFuncSetAttribute is per device
    int dev_id = 0;
    TAMTR_CUDA_OK(cudaGetDevice(&dev_id));
    if (dev_id < 0 || dev_id >= 64 || !attr_set[dev_id]) {
        TAMTR_CUDA_OK(cudaFuncSetAttribute(locw_tc_fwd_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        TAMTR_CUDA_OK(cudaFuncSetAttribute(locw_tc_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        if (dev_id >= 0 && dev_id < 64) attr_set[dev_id] = true;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid((M + kLwTileM - 1) / kLwTileM, H / Hc);
    {
        KernelTimer timer(K_LOCW_FWD, st);
        if (S == 12)
            locw_tc_fwd_kernel<12><<<grid, kLwThreads, smem, st>>>(map_q, map_woff, map_watt, b_off, b_attn,
                                                                   bias_dtype == TAMTR_BF16, ref, loc, attn, raw, M, H, Hc, P,
                                                                   n_kb, n_chunks, chunk, stages, tmem_cols);
        else
            locw_tc_fwd_kernel<16><<<grid, kLwThreads, smem, st>>>(map_q, map_woff, map_watt, b_off, b_attn,
                                                                   bias_dtype == TAMTR_BF16, ref, loc, attn, raw, M, H, Hc, P,
                                                                   n_kb, n_chunks, chunk, stages, tmem_cols);
    }
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}
