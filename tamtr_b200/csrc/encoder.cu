// Encoder-side token kernels of the detection heads (sm_100a): everything that touches the [B, Lv, d] token tensor
// outside the GEMMs.  All three are single-pass, HBM-bound streaming kernels (16-byte accesses, fp32 accumulate).
//
// * tamtr_col_reduce2 / tamtr_affine_rows -- BatchNorm of `input_proj` in TOKEN-MAJOR layout.  The reference computes
//   conv1x1 + BatchNorm2d in NCHW and then flatten(2).permute(0,2,1) + cat (ultralytics/nn/modules/head.py:1202-1218):
//   four passes over the level plus a strided transpose copy, and their mirror images in the backward.  Here the
//   1x1 conv is a GEMM that writes [B, HW_l, d] straight into the level's slice of the token tensor, the batch
//   statistics are column reductions (sum a, sum a*b) and normalisation / its backward are one affine pass
//   out = A*a + Bc*b + Cc with per-(level, channel) coefficients (ops.py::_InputProjFn has the algebra).
// * tamtr_rank_tokens -- query-selection ranking (head.py:1229-1237): per token LayerNorm statistics of
//   enc_output.0's GEMM row, folded with the skinny score GEMM, max over classes.  Replaces valid-mask multiply,
//   LayerNorm (fp32 materialisation of [B, Lv, d]), cast, Linear(d -> nc) epilogue and max.
#include "common.cuh"

namespace tamtr {

constexpr int kEncThreads = 256;

template <typename T> struct Pack;                 // 16-byte pack of T <-> fp32
template <> struct Pack<float> {
    static constexpr int N = 4;
    __device__ static __forceinline__ void load(const float *p, float (&f)[4]) {
        const float4 v = *reinterpret_cast<const float4 *>(p);
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    }
    __device__ static __forceinline__ void store(float *p, const float (&f)[4]) {
        *reinterpret_cast<float4 *>(p) = make_float4(f[0], f[1], f[2], f[3]);
    }
};
template <> struct Pack<__nv_bfloat16> {
    static constexpr int N = 8;
    __device__ static __forceinline__ void load(const __nv_bfloat16 *p, float (&f)[8]) {
        const uint4 u = *reinterpret_cast<const uint4 *>(p);
        f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
        f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
        f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
        f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
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
};

// partial[cta][0][c] = sum over the CTA's rows of a[r][c], partial[cta][1][c] = sum of a[r][c] * b[r][c].
// Rows are the tokens [tok0, tok0 + ntok) of every image of a [B, Lv, d] tensor.  Thread = one 16-byte column pack;
// blockDim.x = d / N packs, blockDim.y row lanes; the y lanes are combined through shared memory.
template <typename T>
__global__ void col_reduce2_kernel(const T *__restrict__ a, const T *__restrict__ b, float *__restrict__ partial,
                                   int Lv, int d, int tok0, int ntok, long rows, int rows_per_cta) {
    constexpr int N = Pack<T>::N;
    extern __shared__ float s_red[];   // [blockDim.y][2][d]
    const int col = threadIdx.x * N;
    float sa[N], sab[N];
#pragma unroll
    for (int i = 0; i < N; ++i) sa[i] = sab[i] = 0.f;
    const long r0 = (long)blockIdx.x * rows_per_cta;
    const long r1 = min(rows, r0 + rows_per_cta);
    // (image, token) of the row are carried along instead of being divided out of r for every row (a 64-bit division per
    // 32 bytes loaded made the kernel instruction-bound: 61 % of HBM peak); two rows per iteration keep four 16-byte
    // loads in flight per thread
    long r = r0 + threadIdx.y;
    int img = (int)(r / ntok), t = (int)(r - (long)img * ntok);
    const int step = blockDim.y;
    auto advance = [&](int &im, int &tk) { tk += step; while (tk >= ntok) { tk -= ntok; ++im; } };
    for (; r + step < r1; r += 2 * step) {
        int img2 = img, t2 = t;
        advance(img2, t2);
        const size_t off = ((size_t)img * Lv + tok0 + t) * d + col, off2 = ((size_t)img2 * Lv + tok0 + t2) * d + col;
        float fa[N], fb[N], ga[N], gb[N];
        Pack<T>::load(a + off, fa);
        Pack<T>::load(b + off, fb);
        Pack<T>::load(a + off2, ga);
        Pack<T>::load(b + off2, gb);
#pragma unroll
        for (int i = 0; i < N; ++i) { sa[i] += fa[i]; sab[i] = fmaf(fa[i], fb[i], sab[i]); }
#pragma unroll
        for (int i = 0; i < N; ++i) { sa[i] += ga[i]; sab[i] = fmaf(ga[i], gb[i], sab[i]); }
        img = img2; t = t2;
        advance(img, t);
    }
    if (r < r1) {
        const size_t off = ((size_t)img * Lv + tok0 + t) * d + col;
        float fa[N], fb[N];
        Pack<T>::load(a + off, fa);
        Pack<T>::load(b + off, fb);
#pragma unroll
        for (int i = 0; i < N; ++i) { sa[i] += fa[i]; sab[i] = fmaf(fa[i], fb[i], sab[i]); }
    }
    float *mine = s_red + (size_t)threadIdx.y * 2 * d;
#pragma unroll
    for (int i = 0; i < N; ++i) { mine[col + i] = sa[i]; mine[d + col + i] = sab[i]; }
    __syncthreads();
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    for (int j = tid; j < 2 * d; j += nthr) {
        float s = 0.f;
        for (int y = 0; y < blockDim.y; ++y) s += s_red[(size_t)y * 2 * d + j];
        partial[(size_t)blockIdx.x * 2 * d + j] = s;
    }
}

// BatchNorm bookkeeping of one pyramid level, one thread per channel: folds the column-reduction partials (in fp64:
// var = E[x^2] - mu^2 cancels) and emits everything the affine pass needs.  Replaces ~15 single-digit-microsecond
// library launches per level and direction (fp64 sum over the partials + scalar algebra on [d] vectors).
//   forward : mu = S1/M, var = max(S2/M - mu^2, 0), rstd = 1/sqrt(var + eps),
//             scale = gamma*rstd, shift = beta - mu*scale; running stats <- (1-mom)*old + mom*(mu, var*M/(M-1))
// fold of the partials for one channel: blockDim = (32 channels, 32 lanes); lane y sums partials y, y+32, ... in fp64,
// shared memory combines the lanes.  Returns the totals to the threads with threadIdx.y == 0.
__device__ __forceinline__ void bn_fold_partials(const float *__restrict__ partial, int n_cta, int d, int c, double &t1,
                                                 double &t2) {
    __shared__ double s_fold[2][32][33];
    double s1 = 0.0, s2 = 0.0;
    if (c < d) {
        for (int i = threadIdx.y; i < n_cta; i += 32) {
            s1 += (double)partial[(size_t)i * 2 * d + c];
            s2 += (double)partial[(size_t)i * 2 * d + d + c];
        }
    }
    s_fold[0][threadIdx.y][threadIdx.x] = s1;
    s_fold[1][threadIdx.y][threadIdx.x] = s2;
    __syncthreads();
    t1 = t2 = 0.0;
    if (threadIdx.y == 0) {
        for (int y = 0; y < 32; ++y) { t1 += s_fold[0][y][threadIdx.x]; t2 += s_fold[1][y][threadIdx.x]; }
    }
}

__global__ void __launch_bounds__(1024)
bn_forward_coeffs_kernel(const float *__restrict__ partial, int n_cta, double M, const float *__restrict__ gamma,
                         const float *__restrict__ beta, double eps, int use_batch_stats, double momentum,
                         float *__restrict__ running_mean, float *__restrict__ running_var, float *__restrict__ scale,
                         float *__restrict__ shift, double *__restrict__ mu_out, double *__restrict__ rstd_out, int d) {
    const int c = blockIdx.x * 32 + threadIdx.x;
    double s1 = 0.0, s2 = 0.0;
    if (use_batch_stats) bn_fold_partials(partial, n_cta, d, c, s1, s2);
    if (c >= d || threadIdx.y != 0) return;
    double mu, var;
    if (use_batch_stats) {
        mu = s1 / M;
        var = s2 / M - mu * mu;
        if (var < 0.0) var = 0.0;
        if (running_mean != nullptr) {
            const double unbiased = var * (M / (M > 1.0 ? M - 1.0 : 1.0));
            running_mean[c] = (float)((double)running_mean[c] * (1.0 - momentum)) + (float)mu * (float)momentum;
            running_var[c] = (float)((double)running_var[c] * (1.0 - momentum)) + (float)unbiased * (float)momentum;
        }
    } else {
        mu = (double)running_mean[c];
        var = (double)running_var[c];
    }
    const double rstd = 1.0 / sqrt(var + eps);
    const double sc = (double)gamma[c] * rstd;
    scale[c] = (float)sc;
    shift[c] = (float)((double)beta[c] - mu * sc);
    mu_out[c] = mu;
    rstd_out[c] = rstd;
}

//   backward: SG = sum G, SGP = sum G*pre;  d_beta = SG, d_gamma = rstd*(SGP - mu*SG), a = gamma*rstd,
//             d_pre = A*G + Bc*pre + Cc with A = a, Bc = -a*rstd*d_gamma/M, Cc = -a*SG/M + a*rstd*mu*d_gamma/M
//             (eval-mode statistics: Bc = Cc = 0)
__global__ void __launch_bounds__(1024)
bn_backward_coeffs_kernel(const float *__restrict__ partial, int n_cta, double M, const float *__restrict__ gamma,
                          const double *__restrict__ mu, const double *__restrict__ rstd, int batch_stats,
                          float *__restrict__ A, float *__restrict__ Bc, float *__restrict__ Cc,
                          float *__restrict__ d_gamma, float *__restrict__ d_beta, int d) {
    const int c = blockIdx.x * 32 + threadIdx.x;
    double sg, sgp;
    bn_fold_partials(partial, n_cta, d, c, sg, sgp);
    if (c >= d || threadIdx.y != 0) return;
    const double dgam = rstd[c] * (sgp - mu[c] * sg);
    const double a = (double)gamma[c] * rstd[c];
    A[c] = (float)a;
    Bc[c] = batch_stats ? (float)(-a * rstd[c] * dgam / M) : 0.f;
    Cc[c] = batch_stats ? (float)(-a * sg / M + a * rstd[c] * mu[c] * dgam / M) : 0.f;
    d_gamma[c] = (float)dgam;
    d_beta[c] = (float)sg;
}

// out[c] += sum over this CTA's rows of g[r][c]  (bias gradient of a Linear layer: g is [rows, n], out is zeroed by the
// launcher).  Thread = one 16-byte column pack, blockDim.y row lanes, blockIdx.y row chunks; one fp32 reduction per
// column and CTA.  The library's generic reduction takes 16 us for [4800, 512] bf16, 45 times per training step.
template <typename T>
__global__ void __launch_bounds__(256) col_sum_kernel(const T *__restrict__ g, float *__restrict__ out, int rows, int n,
                                                      int rows_per_cta) {
    constexpr int N = Pack<T>::N;
    __shared__ float s_part[8][32 * 8 + 8];
    const int pack = blockIdx.x * 32 + threadIdx.x;          // column pack handled by this thread
    const int col = pack * N;
    float acc[N];
#pragma unroll
    for (int i = 0; i < N; ++i) acc[i] = 0.f;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    if (col < n) {
        int r = r0 + threadIdx.y;
        for (; r + 8 < r1; r += 16) {                         // two independent 16-byte loads in flight
            float f0[N], f1[N];
            Pack<T>::load(g + (size_t)r * n + col, f0);
            Pack<T>::load(g + (size_t)(r + 8) * n + col, f1);
#pragma unroll
            for (int i = 0; i < N; ++i) acc[i] += f0[i] + f1[i];
        }
        for (; r < r1; r += 8) {
            float f0[N];
            Pack<T>::load(g + (size_t)r * n + col, f0);
#pragma unroll
            for (int i = 0; i < N; ++i) acc[i] += f0[i];
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) s_part[threadIdx.y][threadIdx.x * N + i] = acc[i];
    __syncthreads();
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int j = tid; j < 32 * N; j += 256) {
        const int cc = blockIdx.x * 32 * N + j;
        if (cc < n) {
            float t = 0.f;
#pragma unroll
            for (int y = 0; y < 8; ++y) t += s_part[y][j];
            atomicAdd(out + cc, t);
        }
    }
}

struct RowLevels {
    int n;
    int start[kMaxLevels + 1];   // token starts, start[n] = Lv
};

// out[b,t,c] = A[l][c] * a[b,t,c] + Bc[l][c] * b[b,t,c] + Cc[l][c]    (l = level of token t; b may be null)
// blockDim.x = d / N column packs (each thread keeps its per-level coefficients in registers while it walks down
// the rows of one level), blockDim.y row lanes; blockIdx.y = level, blockIdx.x strides over that level's rows.
template <typename T, bool HAS_B>
__global__ void affine_rows_kernel(T *__restrict__ out, const T *__restrict__ a, const T *__restrict__ b,
                                   const float *__restrict__ A, const float *__restrict__ Bc,
                                   const float *__restrict__ Cc, const RowLevels lv, int B, int Lv, int d) {
    constexpr int N = Pack<T>::N;
    const int l = blockIdx.y;
    const int col = threadIdx.x * N;
    const int tok0 = lv.start[l], ntok = lv.start[l + 1] - tok0;
    float ca[N], cb[N], cc[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        ca[i] = __ldg(A + (size_t)l * d + col + i);
        cc[i] = __ldg(Cc + (size_t)l * d + col + i);
        cb[i] = HAS_B ? __ldg(Bc + (size_t)l * d + col + i) : 0.f;
    }
    const int rows = B * ntok;
    for (int r = blockIdx.x * blockDim.y + threadIdx.y; r < rows; r += gridDim.x * blockDim.y) {
        const int img = r / ntok, t = tok0 + r - img * ntok;
        const size_t off = ((size_t)img * Lv + t) * d + col;
        float fa[N], fb[N], fo[N];
        Pack<T>::load(a + off, fa);
        if (HAS_B) Pack<T>::load(b + off, fb);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            float v = fmaf(ca[i], fa[i], cc[i]);
            if (HAS_B) v = fmaf(cb[i], fb[i], v);
            fo[i] = v;
        }
        Pack<T>::store(out + off, fo);
    }
}

// One warp per token.  v = valid ? E[row] + enc_bias : enc_bias;  (mean, rstd) = LayerNorm statistics of v;
// score_k = rstd * (raw[row][k] + bw[k] - mean * sw[k]) + ck[k];  out[row] = max_k score_k
//   raw = E @ (score_w * ln_w)^T  (library GEMM on the un-biased rows),  bw = enc_bias . W', sw = sum_c W'[k][c],
//   ck = ln_b . score_w[k] + score_b[k].  Invalid tokens (anchor too close to the border, head.py:1195-1199) have the
//   constant row enc_bias: raw is ignored for them (raw_invalid = 0).
template <typename T, int PP>   // PP = 16-byte packs per lane = ceil(d / (32 * N))
__global__ void __launch_bounds__(kEncThreads)
rank_tokens_kernel(const T *__restrict__ E, const float *__restrict__ raw, const float *__restrict__ enc_bias,
                   const uint8_t *__restrict__ valid, const float *__restrict__ bw, const float *__restrict__ sw,
                   const float *__restrict__ ck, float *__restrict__ out, long rows, int Lv, int d, int nc, int raw_stride,
                   float eps) {
    constexpr int N = Pack<T>::N;
    constexpr int TOK = 4;            // tokens per warp: 4 * PP independent 16-byte loads in flight per lane
    const int lane = threadIdx.x & 31;
    const int nrows = (int)rows;      // B*Lv < 2^31 (checked by the launcher): 32-bit index math
    const int row0 = (blockIdx.x * (kEncThreads / 32) + (threadIdx.x >> 5)) * TOK;
    if (row0 >= nrows) return;
    float v[TOK][PP][N];
    bool ok[TOK];
    float rawv[TOK];
    uint8_t vld[TOK];
    const int t0 = row0 % Lv;
    // every global load of the warp's TOK tokens is issued up front and independently -- the validity bytes, the embedding
    // rows (loaded whether valid or not, zeroed afterwards) and the score-GEMM row -- so that a warp pays ONE DRAM round trip
    // (the first version chained three: valid -> E -> statistics -> raw)
#pragma unroll
    for (int t = 0; t < TOK; ++t) {
        const int row = row0 + t;
        int tok = t0 + t;
        tok = tok >= Lv ? tok - Lv : tok;
        vld[t] = row < nrows ? __ldg(valid + tok) : (uint8_t)0;
        rawv[t] = (row < nrows && nc <= 32 && lane < nc) ? __ldg(raw + (size_t)row * raw_stride + lane) : 0.f;
#pragma unroll
        for (int p = 0; p < PP; ++p) {
            const int c = (p * 32 + lane) * N;
#pragma unroll
            for (int i = 0; i < N; ++i) v[t][p][i] = 0.f;
            if (row < nrows && c < d) Pack<T>::load(E + (size_t)row * d + c, v[t][p]);
        }
    }
#pragma unroll
    for (int t = 0; t < TOK; ++t) {
        ok[t] = vld[t] != 0;
        if (!ok[t]) {
#pragma unroll
            for (int p = 0; p < PP; ++p)
#pragma unroll
                for (int i = 0; i < N; ++i) v[t][p][i] = 0.f;
        }
    }
    float eb[PP][N];
#pragma unroll
    for (int p = 0; p < PP; ++p) {
        const int c = (p * 32 + lane) * N;
#pragma unroll
        for (int i = 0; i < N; i += 4) {
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < d) q = __ldg(reinterpret_cast<const float4 *>(enc_bias + c + i));
            eb[p][i] = q.x; eb[p][i + 1] = q.y; eb[p][i + 2] = q.z; eb[p][i + 3] = q.w;
        }
    }
    const float inv_d = 1.0f / (float)d;
    // the TOK tokens are reduced in lockstep so that their shuffle / load chains overlap (ILP across tokens)
    float s[TOK], ss[TOK], best[TOK];
#pragma unroll
    for (int t = 0; t < TOK; ++t) {
        s[t] = 0.f;
#pragma unroll
        for (int p = 0; p < PP; ++p)
#pragma unroll
            for (int i = 0; i < N; ++i) { v[t][p][i] += eb[p][i]; s[t] += v[t][p][i]; }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1)
#pragma unroll
        for (int t = 0; t < TOK; ++t) s[t] += __shfl_xor_sync(0xffffffffu, s[t], m);
#pragma unroll
    for (int t = 0; t < TOK; ++t) {
        const float mean = s[t] * inv_d;
        s[t] = mean;
        ss[t] = 0.f;
#pragma unroll
        for (int p = 0; p < PP; ++p) {
            const int c = (p * 32 + lane) * N;
            if (c < d) {
#pragma unroll
                for (int i = 0; i < N; ++i) { const float dlt = v[t][p][i] - mean; ss[t] = fmaf(dlt, dlt, ss[t]); }
            }
        }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1)
#pragma unroll
        for (int t = 0; t < TOK; ++t) ss[t] += __shfl_xor_sync(0xffffffffu, ss[t], m);
    const float cbw = lane < nc ? __ldg(bw + lane) : 0.f, csw = lane < nc ? __ldg(sw + lane) : 0.f,
                cck = lane < nc ? __ldg(ck + lane) : 0.f;
#pragma unroll
    for (int t = 0; t < TOK; ++t) {
        const int row = row0 + t;
        const float rstd = rsqrtf(ss[t] * inv_d + eps);
        best[t] = -INFINITY;
        if (row < nrows) {
            if (nc <= 32) {
                if (lane < nc) {
                    const float r = ok[t] ? rawv[t] : 0.f;
                    best[t] = rstd * (r + cbw - s[t] * csw) + cck;
                }
            } else {
                for (int k = lane; k < nc; k += 32) {
                    const float r = ok[t] ? __ldg(raw + (size_t)row * raw_stride + k) : 0.f;
                    best[t] = fmaxf(best[t], rstd * (r + __ldg(bw + k) - s[t] * __ldg(sw + k)) + __ldg(ck + k));
                }
            }
        }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1)
#pragma unroll
        for (int t = 0; t < TOK; ++t) best[t] = fmaxf(best[t], __shfl_xor_sync(0xffffffffu, best[t], m));
    if (lane < TOK && row0 + lane < nrows) {
        float o = best[0];
#pragma unroll
        for (int t = 1; t < TOK; ++t) o = lane == t ? best[t] : o;
        out[row0 + lane] = o;
    }
}

static int enc_check(int dtype, int d) {
    TAMTR_CHECK_ARG(dtype == TAMTR_F32 || dtype == TAMTR_BF16, TAMTR_E_UNSUPPORTED, "encoder: dtype %d", dtype);
    const int n = dtype == TAMTR_F32 ? 4 : 8;
    TAMTR_CHECK_ARG(d > 0 && d % n == 0 && d <= 1024, TAMTR_E_UNSUPPORTED,
                    "encoder: channel dim %d unsupported (multiple of %d, <= 1024)", d, n);
    return 0;
}

}  // namespace tamtr

using namespace tamtr;

extern "C" int tamtr_col_reduce2_ctas(int B, int ntok) {
    const long rows = (long)B * ntok;
    long n = (rows + 127) / 128;
    if (n > 592) n = 592;   // 4 CTAs per SM
    return (int)(n < 1 ? 1 : n);
}

extern "C" int tamtr_col_reduce2(const void *a, const void *b, float *partial, int dtype, int B, int Lv, int d, int tok0,
                                 int ntok, void *stream) {
    TAMTR_CHECK_ARG(a && b && partial, TAMTR_E_BADARG, "col_reduce2: null pointer");
    TAMTR_CHECK_ARG(B > 0 && ntok > 0 && tok0 >= 0 && tok0 + ntok <= Lv, TAMTR_E_BADARG, "col_reduce2: bad token range");
    const int rc = enc_check(dtype, d);
    if (rc) return rc;
    const long rows = (long)B * ntok;
    const int ctas = tamtr_col_reduce2_ctas(B, ntok);
    const int rpc = (int)((rows + ctas - 1) / ctas);
    const int n = dtype == TAMTR_F32 ? 4 : 8;
    const int bx = d / n;
    int by = kEncThreads / bx;
    if (by < 1) by = 1;
    const size_t smem = sizeof(float) * (size_t)by * 2 * d;
    TAMTR_CHECK_ARG(smem <= 48 * 1024, TAMTR_E_UNSUPPORTED, "col_reduce2: d = %d needs too much shared memory", d);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == TAMTR_F32)
        col_reduce2_kernel<float><<<ctas, dim3(bx, by), smem, st>>>((const float *)a, (const float *)b, partial, Lv, d,
                                                                    tok0, ntok, rows, rpc);
    else
        col_reduce2_kernel<__nv_bfloat16><<<ctas, dim3(bx, by), smem, st>>>(
            (const __nv_bfloat16 *)a, (const __nv_bfloat16 *)b, partial, Lv, d, tok0, ntok, rows, rpc);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_affine_rows(void *out, const void *a, const void *b, const float *A, const float *Bc, const float *Cc,
                                 int dtype, int B, int Lv, int d, int L, const int32_t *level_starts_host, void *stream) {
    TAMTR_CHECK_ARG(out && a && A && Cc && level_starts_host, TAMTR_E_BADARG, "affine_rows: null pointer");
    TAMTR_CHECK_ARG(!b || Bc, TAMTR_E_BADARG, "affine_rows: b given without Bc");
    TAMTR_CHECK_ARG(B > 0 && Lv > 0 && L >= 1 && L <= kMaxLevels, TAMTR_E_BADARG, "affine_rows: bad sizes");
    const int rc = enc_check(dtype, d);
    if (rc) return rc;
    RowLevels lv;
    lv.n = L;
    for (int i = 0; i <= kMaxLevels; ++i) lv.start[i] = i < L ? level_starts_host[i] : Lv;
    const int n = dtype == TAMTR_F32 ? 4 : 8;
    const int bx = d / n;
    int by = kEncThreads / bx;
    if (by < 1) by = 1;
    const dim3 block(bx, by), grid(148 * 4, L);
    cudaStream_t st = (cudaStream_t)stream;
#define AFF(T, HB) affine_rows_kernel<T, HB><<<grid, block, 0, st>>>((T *)out, (const T *)a, (const T *)b, A, Bc, Cc, lv, B, Lv, d)
    if (dtype == TAMTR_F32) { if (b) AFF(float, true); else AFF(float, false); }
    else { if (b) AFF(__nv_bfloat16, true); else AFF(__nv_bfloat16, false); }
#undef AFF
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_rank_tokens(const void *E, const float *raw, const float *enc_bias, const uint8_t *valid,
                                 const float *bw, const float *sw, const float *ck, float *out, int dtype, int B, int Lv,
                                 int d, int nc, int raw_stride, float eps, void *stream) {
    TAMTR_CHECK_ARG(E && raw && enc_bias && valid && bw && sw && ck && out, TAMTR_E_BADARG, "rank_tokens: null pointer");
    TAMTR_CHECK_ARG(B > 0 && Lv > 0 && nc > 0 && raw_stride >= nc, TAMTR_E_BADARG, "rank_tokens: bad sizes");
    const int rc = enc_check(dtype, d);
    if (rc) return rc;
    const int n = dtype == TAMTR_F32 ? 4 : 8;
    TAMTR_CHECK_ARG(d <= 4 * 32 * n, TAMTR_E_UNSUPPORTED, "rank_tokens: d = %d too large", d);
    const long rows = (long)B * Lv;
    TAMTR_CHECK_ARG(rows < (1L << 31) - 64, TAMTR_E_UNSUPPORTED, "rank_tokens: too many tokens");
    const long per_cta = (kEncThreads / 32) * 4;
    const unsigned blocks = (unsigned)((rows + per_cta - 1) / per_cta);
    const int pp = (d + 32 * n - 1) / (32 * n);
    cudaStream_t st = (cudaStream_t)stream;
#define RANK(T, PPV) rank_tokens_kernel<T, PPV><<<blocks, kEncThreads, 0, st>>>((const T *)E, raw, enc_bias, valid, bw, sw, ck, out, rows, Lv, d, nc, raw_stride, eps)
    if (dtype == TAMTR_F32) { if (pp <= 1) RANK(float, 1); else if (pp <= 2) RANK(float, 2); else RANK(float, 4); }
    else { if (pp <= 1) RANK(__nv_bfloat16, 1); else if (pp <= 2) RANK(__nv_bfloat16, 2); else RANK(__nv_bfloat16, 4); }
#undef RANK
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_bn_forward_coeffs(const float *partial, int n_cta, double M, const float *gamma, const float *beta,
                                       double eps, int use_batch_stats, double momentum, float *running_mean,
                                       float *running_var, float *scale, float *shift, double *mu, double *rstd, int d,
                                       void *stream) {
    TAMTR_CHECK_ARG(gamma && beta && scale && shift && mu && rstd, TAMTR_E_BADARG, "bn_forward_coeffs: null pointer");
    TAMTR_CHECK_ARG(d > 0 && M > 0, TAMTR_E_BADARG, "bn_forward_coeffs: bad sizes");
    TAMTR_CHECK_ARG(use_batch_stats ? (partial != nullptr && n_cta > 0) : (running_mean && running_var), TAMTR_E_BADARG,
                    "bn_forward_coeffs: statistics source missing");
    TAMTR_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), TAMTR_E_BADARG,
                    "bn_forward_coeffs: running_mean / running_var must come together");
    bn_forward_coeffs_kernel<<<(d + 31) / 32, dim3(32, 32), 0, (cudaStream_t)stream>>>(
        partial, n_cta, M, gamma, beta, eps, use_batch_stats, momentum, running_mean, running_var, scale, shift, mu, rstd, d);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_bn_backward_coeffs(const float *partial, int n_cta, double M, const float *gamma, const double *mu,
                                        const double *rstd, int batch_stats, float *A, float *Bc, float *Cc,
                                        float *d_gamma, float *d_beta, int d, void *stream) {
    TAMTR_CHECK_ARG(partial && gamma && mu && rstd && A && Bc && Cc && d_gamma && d_beta, TAMTR_E_BADARG,
                    "bn_backward_coeffs: null pointer");
    TAMTR_CHECK_ARG(d > 0 && M > 0 && n_cta > 0, TAMTR_E_BADARG, "bn_backward_coeffs: bad sizes");
    bn_backward_coeffs_kernel<<<(d + 31) / 32, dim3(32, 32), 0, (cudaStream_t)stream>>>(partial, n_cta, M, gamma, mu, rstd,
                                                                                 batch_stats, A, Bc, Cc, d_gamma, d_beta, d);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_col_sum(const void *g, float *out, int dtype, int rows, int n, void *stream) {
    TAMTR_CHECK_ARG(g && out, TAMTR_E_BADARG, "col_sum: null pointer");
    TAMTR_CHECK_ARG(rows > 0 && n > 0, TAMTR_E_BADARG, "col_sum: bad sizes");
    TAMTR_CHECK_ARG(dtype == TAMTR_F32 || dtype == TAMTR_BF16, TAMTR_E_UNSUPPORTED, "col_sum: dtype %d", dtype);
    const int np = dtype == TAMTR_F32 ? 4 : 8;
    TAMTR_CHECK_ARG(n % np == 0 && ((uintptr_t)g & 15) == 0, TAMTR_E_UNSUPPORTED,
                    "col_sum: n = %d must be a multiple of %d and g 16-byte aligned", n, np);
    cudaStream_t st = (cudaStream_t)stream;
    TAMTR_CUDA_OK(cudaMemsetAsync(out, 0, (size_t)n * sizeof(float), st));
    const int gx = (n / np + 31) / 32;
    int n_sm = 148;
    n_sm = ::tamtr::sm_count();
    int gy = (2 * n_sm + gx - 1) / gx;                       // ~2 CTAs per SM in total
    const int max_gy = (rows + 15) / 16;                     // at least 16 rows per CTA
    if (gy > max_gy) gy = max_gy;
    if (gy < 1) gy = 1;
    const int rpc = (rows + gy - 1) / gy;
    if (dtype == TAMTR_F32)
        col_sum_kernel<float><<<dim3(gx, gy), dim3(32, 8), 0, st>>>((const float *)g, out, rows, n, rpc);
    else
        col_sum_kernel<__nv_bfloat16><<<dim3(gx, gy), dim3(32, 8), 0, st>>>((const __nv_bfloat16 *)g, out, rows, n, rpc);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}
