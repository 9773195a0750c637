// Glue of the folded encoder side (tamtr_b200/fold.py) as a handful of kernels instead of ~250 tiny tensor-library launches
// per training step: everything between the token reductions / projections of csrc/tokgemm.cu and the small dense
// products (W Cov, W_value A) that stay library GEMMs.  All levels of the pyramid are handled by ONE launch each
// (blockIdx.y = level; per-level pointers travel in a by-value struct).
//
//   tamtr_fold_stats      partials of tok_reduce(X, X)  ->  mean(x) [C], Cov(x) [C, C]                       (per level)
//   tamtr_fold_bn         BatchNorm2d of the 1x1 conv output from (W Cov, W, mean): mu, var, running statistics
//                         (torch/nn/modules/batchnorm.py:155-193), s = gamma * rstd, t = beta - mu * s,
//                         A_ext = [diag(s) W | t]  in both layouts ([L, d, K] and [L, K, d], K = Cmax + 1)
//   tamtr_fold_pack       folded weights [L, N, K] fp32 -> per level a contiguous bf16 [N, C_l] operand + fp32 bias
//   tamtr_fold_unpack     partials of tok_reduce(grad_value, X) -> d(folded weights) [L, N0, K] (+ its [N0, L, K] copy),
//                         d(value_proj bias)
//   tamtr_fold_bn_bwd     d(A_ext) -> d(conv weight), d(gamma), d(beta)  (through s, t, mu and var)
//   tamtr_fold_gather     the rows picked by the query selection: X columns of (image, token) pairs -> [R, L * K] blocks
//                         (zero outside the pair's level, 1 in the level's bias column) so that feats rows = Xcat @ A_ext^T
#include "common.cuh"

namespace tamtr {

struct FoldPtrs {
    int L, d, Cm, K;                     // levels, hidden dim, widest level; K = Cm + 1 rounded up to 8 (column Cm = the
                                         // bias part, columns past it zero: keeps the library GEMMs on aligned operands)
    int C[kMaxLevels];
    int S[kMaxLevels];                   // split counts of the partial buffers
    float n_tok[kMaxLevels];
    const float *part_d[kMaxLevels];     // [S, M, C]
    const float *part_rs[kMaxLevels];    // [S, M]
    float *mean_x[kMaxLevels];           // [C]
    float *cov[kMaxLevels];              // [C, C]
    const float *wc[kMaxLevels];         // conv weight [d, C]
    const float *P[kMaxLevels];          // wc @ cov [d, C]
    const float *gamma[kMaxLevels];
    const float *beta[kMaxLevels];
    float *run_mean[kMaxLevels];
    float *run_var[kMaxLevels];
    long long *n_batches[kMaxLevels];
    float momentum[kMaxLevels];
    float eps[kMaxLevels];
    float *d_wc[kMaxLevels];             // gradients
    float *d_gamma[kMaxLevels];
    float *d_beta[kMaxLevels];
    void *w_out[kMaxLevels];             // bf16 [N, C]
};

__global__ void fold_stats_kernel(const FoldPtrs p) {
    // block = 256 consecutive entries e = r * C + c of one level's [C, C] matrix; the means of all C channels are first
    // formed in shared memory by the block (S * C coalesced loads), then every thread sums the S partials of its entry
    __shared__ float s_mean[512];
    const int l = blockIdx.y, C = p.C[l], S = p.S[l];
    if ((long)blockIdx.x * blockDim.x >= (long)C * C) return;
    const float *pd = p.part_d[l], *pr = p.part_rs[l];
    const float inv = 1.0f / p.n_tok[l];
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
        int s = 0;
        for (; s + 3 < S; s += 4) {
            m0 += __ldg(pr + (size_t)s * C + c);
            m1 += __ldg(pr + (size_t)(s + 1) * C + c);
            m2 += __ldg(pr + (size_t)(s + 2) * C + c);
            m3 += __ldg(pr + (size_t)(s + 3) * C + c);
        }
        for (; s < S; ++s) m0 += __ldg(pr + (size_t)s * C + c);
        s_mean[c] = ((m0 + m1) + (m2 + m3)) * inv;
    }
    __syncthreads();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= C * C) return;
    const int r = e / C, c = e - r * C;
    const size_t cc = (size_t)C * C;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
    int s = 0;
    for (; s + 3 < S; s += 4) {
        g0 += __ldg(pd + (size_t)s * cc + e);
        g1 += __ldg(pd + (size_t)(s + 1) * cc + e);
        g2 += __ldg(pd + (size_t)(s + 2) * cc + e);
        g3 += __ldg(pd + (size_t)(s + 3) * cc + e);
    }
    for (; s < S; ++s) g0 += __ldg(pd + (size_t)s * cc + e);
    p.cov[l][e] = ((g0 + g1) + (g2 + g3)) * inv - s_mean[r] * s_mean[c];
    if (c == 0) p.mean_x[l][r] = s_mean[r];
}

// one warp per (level, output channel j)
__global__ void fold_bn_kernel(const FoldPtrs p, float *__restrict__ a_ext, float *__restrict__ a_ext_t,
                               float *__restrict__ stats, int batch_stats, int update_running) {
    const int l = blockIdx.y, lane = threadIdx.x & 31;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= p.d) return;
    const int C = p.C[l], K = p.K;
    const float *w = p.wc[l] + (size_t)j * C;
    float mu, var;
    if (batch_stats) {
        const float *Pj = p.P[l] + (size_t)j * C, *mx = p.mean_x[l];
        float a = 0.f, b = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float wv = __ldg(w + c);
            a = fmaf(__ldg(Pj + c), wv, a);
            b = fmaf(wv, __ldg(mx + c), b);
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, m);
            b += __shfl_xor_sync(0xffffffffu, b, m);
        }
        var = fmaxf(a, 0.f);
        mu = b;
        if (update_running && lane == 0 && p.run_mean[l] != nullptr) {
            const float mom = p.momentum[l], n = p.n_tok[l];
            p.run_mean[l][j] = (1.f - mom) * p.run_mean[l][j] + mom * mu;
            p.run_var[l][j] = (1.f - mom) * p.run_var[l][j] + mom * var * (n / fmaxf(n - 1.f, 1.f));
            if (j == 0 && p.n_batches[l] != nullptr) *p.n_batches[l] += 1;
        }
    } else {
        mu = __ldg(p.run_mean[l] + j);
        var = __ldg(p.run_var[l] + j);
    }
    const float r = rsqrtf(var + p.eps[l]);
    const float s = __ldg(p.gamma[l] + j) * r;
    const float t = __ldg(p.beta[l] + j) - mu * s;
    if (lane == 0) {
        float *st = stats + ((size_t)l * p.d + j) * 4;
        st[0] = mu; st[1] = r; st[2] = s; st[3] = t;
    }
    float *row = a_ext + ((size_t)l * p.d + j) * K;
    float *col = a_ext_t + (size_t)l * K * p.d + j;
    for (int c = lane; c < K; c += 32) {
        const float v = c < C ? s * __ldg(w + c) : (c == p.Cm ? t : 0.f);
        row[c] = v;
        col[(size_t)c * p.d] = v;
    }
}

// F [L, N, K] fp32 (K = Cm + 1: folded weights, then the folded bias part) -> bf16 [N, C_l] per level, bias [L, N]
// (+ bv[n] for n < N0)
__global__ void fold_pack_kernel(const FoldPtrs p, const float *__restrict__ Fv, const float *__restrict__ Fe,
                                 const float *__restrict__ bv, float *__restrict__ bias, int N0, int NE) {
    const int l = blockIdx.y, n = blockIdx.x, C = p.C[l], K = p.K, N = N0 + NE;
    const float *src = n < N0 ? Fv + ((size_t)l * N0 + n) * K : Fe + ((size_t)l * NE + (n - N0)) * K;
    __nv_bfloat16 *dst = reinterpret_cast<__nv_bfloat16 *>(p.w_out[l]) + (size_t)n * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) dst[c] = __float2bfloat16_rn(__ldg(src + c));
    if (threadIdx.x == 0) bias[(size_t)l * N + n] = __ldg(src + p.Cm) + (n < N0 ? __ldg(bv + n) : 0.f);
}

// partials of the per-level weight-gradient reductions -> dF [L, N0, K] and its [N0, L, K] copy; d_bv [N0]
__global__ void fold_unpack_kernel(const FoldPtrs p, float *__restrict__ dF, float *__restrict__ dF_t,
                                   float *__restrict__ d_bv, int N0) {
    const int l = blockIdx.y, n = blockIdx.x, C = p.C[l], K = p.K, S = p.S[l];
    const float *pd = p.part_d[l], *pr = p.part_rs[l];
    for (int c = threadIdx.x; c < K; c += blockDim.x) {
        float v = 0.f;
        if (c < C) {
            for (int s = 0; s < S; ++s) v += __ldg(pd + ((size_t)s * N0 + n) * C + c);
        } else if (c == p.Cm) {
            for (int s = 0; s < S; ++s) v += __ldg(pr + (size_t)s * N0 + n);
            if (l == 0) {           // bias gradient of value_proj: the row sums of every level
                float tot = v;
                for (int l2 = 1; l2 < p.L; ++l2)
                    for (int s = 0; s < p.S[l2]; ++s) tot += __ldg(p.part_rs[l2] + (size_t)s * N0 + n);
                d_bv[n] = tot;
            }
        }
        dF[((size_t)l * N0 + n) * K + c] = v;
        dF_t[((size_t)n * p.L + l) * K + c] = v;
    }
}

// one warp per (level, channel j): dA [L, d, K] (+ the transposed contribution dAt [L, K, d] of the selected rows, or NULL)
__global__ void fold_bn_bwd_kernel(const FoldPtrs p, const float *__restrict__ dA, const float *__restrict__ dAt,
                                   const float *__restrict__ stats, int batch_stats, float *__restrict__ d_stat) {
    const int l = blockIdx.y, lane = threadIdx.x & 31;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= p.d) return;
    const int C = p.C[l], K = p.K;
    const float *w = p.wc[l] + (size_t)j * C;
    const float *g = dA + ((size_t)l * p.d + j) * K;
    const float *gt = dAt != nullptr ? dAt + (size_t)l * K * p.d + j : nullptr;
    const float *st = stats + ((size_t)l * p.d + j) * 4;
    const float mu = st[0], r = st[1], s = st[2];
    float dt = __ldg(g + p.Cm) + (gt != nullptr ? __ldg(gt + (size_t)p.Cm * p.d) : 0.f);
    float dsw = 0.f;                                         // sum_c dA[j, c] * w[j, c]
    for (int c = lane; c < C; c += 32) {
        const float ga = __ldg(g + c) + (gt != nullptr ? __ldg(gt + (size_t)c * p.d) : 0.f);
        dsw = fmaf(ga, __ldg(w + c), dsw);
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) dsw += __shfl_xor_sync(0xffffffffu, dsw, m);
    const float ds = dsw - dt * mu;                          // A = s w, t = beta - mu s
    const float gam = __ldg(p.gamma[l] + j);
    const float dmu = batch_stats ? -dt * s : 0.f;
    const float dvar = batch_stats ? -0.5f * r * r * r * (ds * gam) : 0.f;     // s = gamma * (var + eps)^-1/2
    if (lane == 0) {
        p.d_gamma[l][j] = ds * r;
        p.d_beta[l][j] = dt;
        if (d_stat != nullptr) {            // d(loss)/d(mu_j), d(loss)/d(var_j): the caller carries them on to the feature maps
            d_stat[((size_t)l * p.d + j) * 2] = dmu;
            d_stat[((size_t)l * p.d + j) * 2 + 1] = dvar;
        }
    }
    float *dw = p.d_wc[l] + (size_t)j * C;
    const float *Pj = batch_stats ? p.P[l] + (size_t)j * C : nullptr, *mx = batch_stats ? p.mean_x[l] : nullptr;
    for (int c = lane; c < C; c += 32) {
        const float ga = __ldg(g + c) + (gt != nullptr ? __ldg(gt + (size_t)c * p.d) : 0.f);
        float v = s * ga;
        if (batch_stats) v += dmu * __ldg(mx + c) + 2.f * dvar * __ldg(Pj + c);     // mu = w . mean, var = w^T Cov w
        dw[c] = v;
    }
}

struct GatherPtrs {
    int L, Cm, K, Lv;
    int C[kMaxLevels], start[kMaxLevels], hw[kMaxLevels];
    const __nv_bfloat16 *x[kMaxLevels];          // [B, C, HW]
};

// xcat[r, l * K + c] = x_l[b, c, t] for the level l that holds token `tok` of pair r = (b, tok), 1 in column l * K + Cm
__global__ void fold_gather_kernel(const GatherPtrs p, const long long *__restrict__ flat_idx, float *__restrict__ xcat,
                                   int R) {
    const int r = blockIdx.x;
    if (r >= R) return;
    const long long fi = flat_idx[r];
    const int b = (int)(fi / p.Lv), tok = (int)(fi - (long long)b * p.Lv);
    const int K = p.K;
    float *row = xcat + (size_t)r * p.L * K;
    for (int l = 0; l < p.L; ++l) {
        const int rel = tok - p.start[l];
        const bool in = rel >= 0 && rel < p.hw[l];
        const __nv_bfloat16 *src = p.x[l] + (size_t)b * p.C[l] * p.hw[l] + (in ? rel : 0);
        for (int c = threadIdx.x; c < K; c += blockDim.x) {
            float v = 0.f;
            if (in) v = c < p.C[l] ? __bfloat162float(src[(size_t)c * p.hw[l]]) : (c == p.Cm ? 1.f : 0.f);
            row[(size_t)l * K + c] = v;
        }
    }
}

__device__ __forceinline__ float fold_ld(const void *p, size_t i, int bf16) {
    return bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(p)[i]) : reinterpret_cast<const float *>(p)[i];
}

// Constants of the ranking (tamtr_tok_project_rank) from the parameters, one launch:
//   we_all [d + NT, d] f32: rows [0, d) = enc_output.0.weight (converted); rows d + k = the COEFFICIENT rows of the tail:
//   class k < nc: Wp[k] = score_w[k] * ln_w; last row (fused mode): enc_bias; other rows 0.  The caller multiplies the tail
//   rows by We in place (tail := coef @ We, one small library GEMM).
//   consts = { sum eb, sum eb^2, bw[NT], sw[NT], ck[NT] }: bw[k] = Wp[k] . eb, sw[k] = sum Wp[k], ck[k] = score_w[k] . ln_b +
//   score_b[k].  Block r < d copies a row of We; block d + k builds coefficient row k and its constants.
__global__ void fold_rank_consts_kernel(const void *__restrict__ We, const void *__restrict__ eb, const void *__restrict__ sw_,
                                        const void *__restrict__ sb, const float *__restrict__ ln_w,
                                        const float *__restrict__ ln_b, float *__restrict__ we_all,
                                        float *__restrict__ consts, int d, int nc, int NT, int fused, int lin_bf16) {
    __shared__ float red[3][32];
    const int r = blockIdx.x, tid = threadIdx.x;
    if (r < d) {
        for (int i = tid; i < d; i += blockDim.x) we_all[(size_t)r * d + i] = fold_ld(We, (size_t)r * d + i, lin_bf16);
        return;
    }
    const int k = r - d;
    float *out = we_all + (size_t)r * d;
    const bool cls = k < nc, dot = fused && k == NT - 1;
    float a = 0.f, b = 0.f, c = 0.f;
    for (int j = tid; j < d; j += blockDim.x) {
        const float e = fold_ld(eb, j, lin_bf16);
        float coef = 0.f;
        if (cls) {
            const float sw = fold_ld(sw_, (size_t)k * d + j, lin_bf16);
            coef = sw * __ldg(ln_w + j);
            a = fmaf(coef, e, a);
            b += coef;
            c = fmaf(sw, __ldg(ln_b + j), c);
        } else if (dot) {
            coef = e;
            a += e;
            b = fmaf(e, e, b);
        }
        out[j] = coef;
    }
    if (!fused) return;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, m);
        b += __shfl_xor_sync(0xffffffffu, b, m);
        c += __shfl_xor_sync(0xffffffffu, c, m);
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = a; red[1][tid >> 5] = b; red[2][tid >> 5] = c; }
    __syncthreads();
    if (tid == 0) {
        a = b = c = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += red[0][w]; b += red[1][w]; c += red[2][w]; }
        if (cls) {
            consts[2 + k] = a;
            consts[2 + NT + k] = b;
            consts[2 + 2 * NT + k] = c + fold_ld(sb, k, lin_bf16);
        } else {
            if (dot) { consts[0] = a; consts[1] = b; }
            consts[2 + k] = 0.f; consts[2 + NT + k] = 0.f; consts[2 + 2 * NT + k] = 0.f;
        }
    }
}

}  // namespace tamtr

using namespace tamtr;

extern "C" int tamtr_fold_rank_consts(const void *We, const void *eb, const void *score_w, const void *score_b,
                                      const float *ln_w, const float *ln_b, float *we_all, float *consts, int d, int nc,
                                      int NT, int fused, int lin_dtype, void *stream) {
    TAMTR_CHECK_ARG(We && eb && score_w && score_b && ln_w && ln_b && we_all && consts, TAMTR_E_BADARG,
                    "fold_rank_consts: null pointer");
    TAMTR_CHECK_ARG(d > 0 && nc > 0 && NT >= nc + (fused ? 1 : 0), TAMTR_E_BADARG, "fold_rank_consts: bad sizes");
    TAMTR_CHECK_ARG(lin_dtype == TAMTR_F32 || lin_dtype == TAMTR_BF16, TAMTR_E_UNSUPPORTED, "fold_rank_consts: dtype %d",
                    lin_dtype);
    fold_rank_consts_kernel<<<d + NT, 256, 0, (cudaStream_t)stream>>>(We, eb, score_w, score_b, ln_w, ln_b, we_all, consts, d,
                                                                      nc, NT, fused, lin_dtype == TAMTR_BF16);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

// The level tables cross the C ABI as arrays of kMaxLevels entries (unused entries ignored).
static int fold_fill(FoldPtrs &p, int L, int d, const int *C) {
    if (L < 1 || L > kMaxLevels || d < 1) return TAMTR_E_BADARG;
    p.L = L; p.d = d; p.Cm = 0;
    for (int l = 0; l < kMaxLevels; ++l) {
        p.C[l] = l < L ? C[l] : 0;
        if (l < L && C[l] < 1) return TAMTR_E_BADARG;
        if (p.C[l] > p.Cm) p.Cm = p.C[l];
    }
    p.K = (p.Cm + 1 + 7) / 8 * 8;
    return 0;
}

extern "C" int tamtr_fold_stats(int L, const int *C, const int *S, const float *n_tok, const float *const *part_d,
                                const float *const *part_rs, float *const *mean_x, float *const *cov, void *stream) {
    TAMTR_CHECK_ARG(C && S && n_tok && part_d && part_rs && mean_x && cov, TAMTR_E_BADARG, "fold_stats: null pointer");
    FoldPtrs p = {};
    TAMTR_CHECK_ARG(fold_fill(p, L, 1, C) == 0, TAMTR_E_BADARG, "fold_stats: bad level table");
    for (int l = 0; l < L; ++l) {
        p.S[l] = S[l]; p.n_tok[l] = n_tok[l]; p.part_d[l] = part_d[l]; p.part_rs[l] = part_rs[l];
        p.mean_x[l] = mean_x[l]; p.cov[l] = cov[l];
        TAMTR_CHECK_ARG(S[l] > 0 && n_tok[l] > 0 && part_d[l] && part_rs[l] && mean_x[l] && cov[l], TAMTR_E_BADARG,
                        "fold_stats: bad level %d", l);
    }
    const dim3 grid((p.Cm * p.Cm + 255) / 256, L);
    fold_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_fold_bn(int L, int d, const int *C, const float *n_tok, const float *const *wc, const float *const *P,
                             const float *const *mean_x, const float *const *gamma, const float *const *beta,
                             float *const *run_mean, float *const *run_var, long long *const *n_batches,
                             const float *momentum, const float *eps, int batch_stats, int update_running, float *a_ext,
                             float *a_ext_t, float *stats, void *stream) {
    TAMTR_CHECK_ARG(C && n_tok && wc && gamma && beta && momentum && eps && a_ext && a_ext_t && stats, TAMTR_E_BADARG,
                    "fold_bn: null pointer");
    TAMTR_CHECK_ARG(!batch_stats || (P && mean_x), TAMTR_E_BADARG, "fold_bn: batch statistics need P and mean_x");
    TAMTR_CHECK_ARG(batch_stats || (run_mean && run_var), TAMTR_E_BADARG, "fold_bn: running statistics missing");
    FoldPtrs p = {};
    TAMTR_CHECK_ARG(fold_fill(p, L, d, C) == 0, TAMTR_E_BADARG, "fold_bn: bad level table");
    for (int l = 0; l < L; ++l) {
        p.n_tok[l] = n_tok[l]; p.wc[l] = wc[l]; p.gamma[l] = gamma[l]; p.beta[l] = beta[l];
        p.P[l] = P ? P[l] : nullptr; p.mean_x[l] = mean_x ? const_cast<float *>(mean_x[l]) : nullptr;
        p.run_mean[l] = run_mean ? run_mean[l] : nullptr; p.run_var[l] = run_var ? run_var[l] : nullptr;
        p.n_batches[l] = n_batches ? n_batches[l] : nullptr;
        p.momentum[l] = momentum[l]; p.eps[l] = eps[l];
        TAMTR_CHECK_ARG(wc[l] && gamma[l] && beta[l], TAMTR_E_BADARG, "fold_bn: bad level %d", l);
        TAMTR_CHECK_ARG(batch_stats || (p.run_mean[l] && p.run_var[l]), TAMTR_E_BADARG, "fold_bn: level %d has no statistics", l);
    }
    const dim3 grid((d + 7) / 8, L);
    fold_bn_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, a_ext, a_ext_t, stats, batch_stats, update_running);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_fold_pack(int L, const int *C, const float *Fv, const float *Fe, const float *bv, void *const *w_out,
                               float *bias, int N0, int NE, void *stream) {
    TAMTR_CHECK_ARG(C && Fv && Fe && bv && w_out && bias && N0 > 0 && NE >= 0, TAMTR_E_BADARG, "fold_pack: bad argument");
    FoldPtrs p = {};
    TAMTR_CHECK_ARG(fold_fill(p, L, 1, C) == 0, TAMTR_E_BADARG, "fold_pack: bad level table");
    for (int l = 0; l < L; ++l) {
        p.w_out[l] = w_out[l];
        TAMTR_CHECK_ARG(w_out[l] != nullptr, TAMTR_E_BADARG, "fold_pack: bad level %d", l);
    }
    fold_pack_kernel<<<dim3(N0 + NE, L), 128, 0, (cudaStream_t)stream>>>(p, Fv, Fe, bv, bias, N0, NE);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_fold_unpack(int L, const int *C, const int *S, const float *const *part_d, const float *const *part_rs,
                                 float *dF, float *dF_t, float *d_bv, int N0, void *stream) {
    TAMTR_CHECK_ARG(C && S && part_d && part_rs && dF && dF_t && d_bv && N0 > 0, TAMTR_E_BADARG, "fold_unpack: bad argument");
    FoldPtrs p = {};
    TAMTR_CHECK_ARG(fold_fill(p, L, 1, C) == 0, TAMTR_E_BADARG, "fold_unpack: bad level table");
    for (int l = 0; l < L; ++l) {
        p.S[l] = S[l]; p.part_d[l] = part_d[l]; p.part_rs[l] = part_rs[l];
        TAMTR_CHECK_ARG(S[l] > 0 && part_d[l] && part_rs[l], TAMTR_E_BADARG, "fold_unpack: bad level %d", l);
    }
    fold_unpack_kernel<<<dim3(N0, L), 256, 0, (cudaStream_t)stream>>>(p, dF, dF_t, d_bv, N0);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_fold_bn_bwd(int L, int d, const int *C, const float *const *wc, const float *const *P,
                                 const float *const *mean_x, const float *const *gamma, const float *dA, const float *dAt,
                                 const float *stats, int batch_stats, float *const *d_wc, float *const *d_gamma,
                                 float *const *d_beta, float *d_stat, void *stream) {
    TAMTR_CHECK_ARG(C && wc && gamma && dA && stats && d_wc && d_gamma && d_beta, TAMTR_E_BADARG, "fold_bn_bwd: null pointer");
    TAMTR_CHECK_ARG(!batch_stats || (P && mean_x), TAMTR_E_BADARG, "fold_bn_bwd: batch statistics need P and mean_x");
    FoldPtrs p = {};
    TAMTR_CHECK_ARG(fold_fill(p, L, d, C) == 0, TAMTR_E_BADARG, "fold_bn_bwd: bad level table");
    for (int l = 0; l < L; ++l) {
        p.wc[l] = wc[l]; p.gamma[l] = gamma[l];
        p.P[l] = P ? P[l] : nullptr; p.mean_x[l] = mean_x ? const_cast<float *>(mean_x[l]) : nullptr;
        p.d_wc[l] = d_wc[l]; p.d_gamma[l] = d_gamma[l]; p.d_beta[l] = d_beta[l];
        TAMTR_CHECK_ARG(wc[l] && gamma[l] && d_wc[l] && d_gamma[l] && d_beta[l], TAMTR_E_BADARG, "fold_bn_bwd: bad level %d", l);
    }
    fold_bn_bwd_kernel<<<dim3((d + 7) / 8, L), 256, 0, (cudaStream_t)stream>>>(p, dA, dAt, stats, batch_stats, d_stat);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_fold_gather(int L, int Lv, const int *C, const int *start, const int *hw, const void *const *x_bf16,
                                 const long long *flat_idx, float *xcat, int R, void *stream) {
    TAMTR_CHECK_ARG(C && start && hw && x_bf16 && flat_idx && xcat && R > 0 && Lv > 0, TAMTR_E_BADARG,
                    "fold_gather: bad argument");
    TAMTR_CHECK_ARG(L >= 1 && L <= kMaxLevels, TAMTR_E_UNSUPPORTED, "fold_gather: %d levels", L);
    GatherPtrs p = {};
    p.L = L; p.Lv = Lv; p.Cm = 0;
    for (int l = 0; l < L; ++l) {
        p.C[l] = C[l]; p.start[l] = start[l]; p.hw[l] = hw[l];
        p.x[l] = reinterpret_cast<const __nv_bfloat16 *>(x_bf16[l]);
        TAMTR_CHECK_ARG(C[l] > 0 && hw[l] > 0 && x_bf16[l], TAMTR_E_BADARG, "fold_gather: bad level %d", l);
        if (C[l] > p.Cm) p.Cm = C[l];
    }
    p.K = (p.Cm + 1 + 7) / 8 * 8;
    fold_gather_kernel<<<R, 128, 0, (cudaStream_t)stream>>>(p, flat_idx, xcat, R);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}
