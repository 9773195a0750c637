// Query selection: the nq best-ranked tokens of every image (ultralytics/nn/modules/head.py:1240 / :437,
// `torch.topk(enc_outputs_scores.max(-1).values, self.num_queries, dim=1).indices`).
//
// The library runs a multi-block radix select + a sort as ~10 launches (~140 us of the 4.2 ms step for 16 x 33 600
// scores, mostly latency).  Here: one CTA per image.  The row's scores are turned into order-preserving 32-bit keys
// and kept in shared memory (134 KB for the 640^2 pyramid; rows that do not fit are re-read from global memory / L2).
// Shared-memory atomics cost ~2 cycles per lane, so a histogram over all n keys would be the whole budget; instead
// every thread keeps the maximum of its strided share, a radix select over those 1024 maxima gives a bound L with at
// least k keys >= L (each maximum is a key), and only the keys >= L -- a few hundred -- are compacted and go through
// the exact 12 + 12 + 8 bit radix select that finds the key of the k-th best token.  (More than 4096 candidates --
// heavily tied rows -- or k > 1024: the select runs over the whole row.)  One ordered pass then collects the k
// winners (every key above the threshold, plus the lowest-indexed tokens among those that equal it), and the winners
// are ranked by (score descending, index ascending) by counting.  Output order = torch.topk(sorted=True) whenever the
// scores are distinct; among equal scores torch leaves the order unspecified, here the lower index comes first.
#include "common.cuh"

namespace tamtr {

constexpr int kTopkThreads = 1024;
constexpr int kTopkWarps = kTopkThreads / 32;
constexpr int kTopkBins = 4096;
constexpr int kTopkCand = 4096;      // capacity of the candidate list

__device__ __forceinline__ uint32_t topk_key(float v) {
    uint32_t u = __float_as_uint(v);
    if (v != v) return 0xffffffffu;                      // NaN of either sign ranks first, as for torch.topk / torch.sort
    if (u == 0x80000000u) u = 0u;                        // -0.0 == +0.0
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // larger float <=> larger key
}

template <bool IN_SMEM>
__device__ __forceinline__ uint32_t topk_load(const uint32_t *keys, const float *row, int i) {
    if (IN_SMEM) return keys[i];
    return topk_key(__ldg(row + i));
}

// One radix step over the keys whose bits above `shift + bits` equal those of `prefix`: finds the digit in which the
// `need`-th largest of them lies.  Returns through s_sel = { digit, tokens needed inside that digit's bin }.
template <bool IN_SMEM>
__device__ __forceinline__ void topk_radix_step(const uint32_t *keys, const float *row, int n, uint32_t prefix,
                                                uint32_t high_mask, int shift, int bits, int need, uint32_t *hist,
                                                uint32_t *warp_tot, uint32_t *s_sel) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nbins = 1 << bits;
    for (int b = tid; b < kTopkBins; b += kTopkThreads) hist[b] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += kTopkThreads) {
        const uint32_t key = topk_load<IN_SMEM>(keys, row, i);
        if ((key & high_mask) == prefix) atomicAdd(&hist[(key >> shift) & (nbins - 1)], 1u);
    }
    __syncthreads();
    // suffix sums: thread t owns bins [4t, 4t + 4); above(t) = tokens in the bins of higher threads
    uint32_t c[4], own = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int b = 4 * tid + j;
        c[j] = b < nbins ? hist[b] : 0u;
        own += c[j];
    }
    uint32_t incl = own;                                   // inclusive suffix sum inside the warp (towards lane 31)
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_down_sync(0xffffffffu, incl, d);
        if (lane + d < 32) incl += o;
    }
    if (lane == 0) warp_tot[warp] = incl;
    __syncthreads();
    uint32_t above = incl - own;
    for (int w = warp + 1; w < kTopkWarps; ++w) above += warp_tot[w];
#pragma unroll
    for (int j = 3; j >= 0; --j) {
        if (above < (uint32_t)need && above + c[j] >= (uint32_t)need) {
            s_sel[0] = (uint32_t)(4 * tid + j);
            s_sel[1] = (uint32_t)need - above;
        }
        above += c[j];
    }
    __syncthreads();
}

template <bool IN_SMEM>
__global__ void __launch_bounds__(kTopkThreads, 1)
topk_rows_kernel(const float *__restrict__ scores, long long *__restrict__ out_idx, float *__restrict__ out_val, int n,
                 int k) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *list = reinterpret_cast<unsigned long long *>(smem_raw);            // [k] winners
    uint32_t *hist = reinterpret_cast<uint32_t *>(list + k);                                // [kTopkBins]
    uint32_t *warp_tot = hist + kTopkBins;                                                  // [2 * warps]
    uint32_t *s_sel = warp_tot + 2 * kTopkWarps;                                            // [4]
    uint32_t *tmax = s_sel + 4;                                                             // [threads]
    uint32_t *cand = tmax + kTopkThreads;                                                   // [kTopkCand]
    uint32_t *keys = cand + kTopkCand;                                                      // [n] (IN_SMEM)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *row = scores + (size_t)blockIdx.x * n;

    uint32_t my_max = 0;
    for (int i = tid; i < n; i += kTopkThreads) {
        const uint32_t key = topk_key(__ldg(row + i));
        if (IN_SMEM) keys[i] = key;
        my_max = max(my_max, key);
    }
    const int shifts[3] = {20, 8, 0}, widths[3] = {12, 12, 8};
    const uint32_t *sel_keys = keys;       // what the exact select runs over: the candidates, or the whole row
    const float *sel_row = row;
    int sel_n = n;
    bool sel_smem = IN_SMEM;
    if (n > kTopkCand && k <= kTopkThreads) {
        // bound L: top 24 bits of the k-th largest per-thread maximum (>= k keys are >= L)
        tmax[tid] = my_max;
        if (tid == 0) s_sel[2] = 0;
        uint32_t prefix = 0, high_mask = 0;
        int need = k;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            topk_radix_step<true>(tmax, nullptr, kTopkThreads, prefix, high_mask, shifts[p], widths[p], need, hist, warp_tot,
                                  s_sel);
            prefix |= s_sel[0] << shifts[p];
            high_mask |= ((1u << widths[p]) - 1u) << shifts[p];
            need = (int)s_sel[1];
            __syncthreads();
        }
        const uint32_t bound = prefix;
        for (int i0 = 0; i0 < n; i0 += kTopkThreads) {
            const int i = i0 + tid;
            const uint32_t key = i < n ? topk_load<IN_SMEM>(keys, row, i) : 0u;
            const bool hit = i < n && key >= bound;
            const uint32_t hb = __ballot_sync(0xffffffffu, hit);
            if (hb) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&s_sel[2], (uint32_t)__popc(hb));
                base = __shfl_sync(0xffffffffu, base, 0);
                const uint32_t at = base + __popc(hb & ((1u << lane) - 1u));
                if (hit && at < (uint32_t)kTopkCand) cand[at] = key;
            }
        }
        __syncthreads();
        const uint32_t n_cand = s_sel[2];
        if (n_cand <= (uint32_t)kTopkCand) {
            sel_keys = cand;
            sel_n = (int)n_cand;
            sel_smem = true;
        }
    }
    // exact radix select, most significant digit first: 12 + 12 + 8 bits
    uint32_t prefix = 0, high_mask = 0;
    int need = k;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        if (sel_smem)
            topk_radix_step<true>(sel_keys, sel_row, sel_n, prefix, high_mask, shifts[p], widths[p], need, hist, warp_tot, s_sel);
        else
            topk_radix_step<false>(sel_keys, sel_row, sel_n, prefix, high_mask, shifts[p], widths[p], need, hist, warp_tot,
                                   s_sel);
        prefix |= s_sel[0] << shifts[p];
        high_mask |= ((1u << widths[p]) - 1u) << shifts[p];
        need = (int)s_sel[1];
        __syncthreads();
    }
    const uint32_t thr = prefix;          // key of the k-th best token; `need` of the tokens that equal it are winners
    const int n_above = k - need;

    // ordered collection: warp w walks the contiguous region [w * per_warp, (w + 1) * per_warp) 32 tokens at a time
    const int per_warp = ((n + kTopkWarps - 1) / kTopkWarps + 31) & ~31;
    const int lo = warp * per_warp, hi = min(n, lo + per_warp);
    uint32_t g_cnt = 0, e_cnt = 0;
    for (int i = lo + lane; i < lo + per_warp; i += 32) {
        const uint32_t key = i < hi ? topk_load<IN_SMEM>(keys, row, i) : 0u;
        g_cnt += __popc(__ballot_sync(0xffffffffu, i < hi && key > thr));
        e_cnt += __popc(__ballot_sync(0xffffffffu, i < hi && key == thr));
    }
    if (lane == 0) {
        warp_tot[warp] = g_cnt;
        warp_tot[kTopkWarps + warp] = e_cnt;
    }
    __syncthreads();
    uint32_t g_pos = 0, e_pos = 0;
    for (int w = 0; w < warp; ++w) {
        g_pos += warp_tot[w];
        e_pos += warp_tot[kTopkWarps + w];
    }
    const uint32_t lt = (1u << lane) - 1u;
    for (int i = lo + lane; i < lo + per_warp; i += 32) {
        const uint32_t key = i < hi ? topk_load<IN_SMEM>(keys, row, i) : 0u;
        const bool g = i < hi && key > thr, e = i < hi && key == thr;
        const uint32_t gb = __ballot_sync(0xffffffffu, g), eb = __ballot_sync(0xffffffffu, e);
        const unsigned long long item = ((unsigned long long)key << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i);
        if (g) list[g_pos + __popc(gb & lt)] = item;
        if (e) {
            const uint32_t r = e_pos + __popc(eb & lt);
            if (r < (uint32_t)need) list[n_above + r] = item;
        }
        g_pos += __popc(gb);
        e_pos += __popc(eb);
    }
    __syncthreads();
    // rank by counting: the items are distinct (index in the low half), so rank = number of larger items
    for (int i = tid; i < k; i += kTopkThreads) {
        const unsigned long long mine = list[i];
        int rank = 0;
        for (int j = 0; j < k; ++j) rank += list[j] > mine;
        const uint32_t idx = 0xffffffffu - (uint32_t)(mine & 0xffffffffull);
        out_idx[(size_t)blockIdx.x * k + rank] = (long long)idx;
        if (out_val) out_val[(size_t)blockIdx.x * k + rank] = __ldg(row + idx);
    }
}

static const size_t kTopkSmemMax = 227 * 1024;
static size_t topk_smem(int n, int k, bool in_smem) {
    return (size_t)k * 8 + (size_t)(kTopkBins + 2 * kTopkWarps + 4 + kTopkThreads + kTopkCand) * 4 + (in_smem ? (size_t)n * 4 : 0);
}

}  // namespace tamtr

using namespace tamtr;

extern "C" int tamtr_topk_rows_supported(int n, int k) {
    if (!(n >= 1 && k >= 1 && k <= n && k <= 4096)) return 0;
    return topk_smem(n, k, true) <= kTopkSmemMax ? 2 : 1;     // 2: the row's keys stay in shared memory
}

extern "C" int tamtr_topk_rows(const float *scores, long long *out_idx, float *out_val, int rows, int n, int k,
                               void *stream) {
    TAMTR_CHECK_ARG(scores && out_idx, TAMTR_E_BADARG, "topk_rows: null pointer");
    TAMTR_CHECK_ARG(rows >= 1, TAMTR_E_BADARG, "topk_rows: rows = %d", rows);
    TAMTR_CHECK_ARG(tamtr_topk_rows_supported(n, k), TAMTR_E_UNSUPPORTED, "topk_rows: n = %d, k = %d (1 <= k <= min(n, 4096))",
                    n, k);
    cudaStream_t st = (cudaStream_t)stream;
    const bool in_smem = topk_smem(n, k, true) <= kTopkSmemMax;
    const size_t smem = topk_smem(n, k, in_smem);
    if (in_smem) {
        TAMTR_CUDA_OK(cudaFuncSetAttribute(topk_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTopkSmemMax));
        topk_rows_kernel<true><<<rows, kTopkThreads, smem, st>>>(scores, out_idx, out_val, n, k);
    } else {
        TAMTR_CUDA_OK(cudaFuncSetAttribute(topk_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTopkSmemMax));
        topk_rows_kernel<false><<<rows, kTopkThreads, smem, st>>>(scores, out_idx, out_val, n, k);
    }
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}
