// Bipartite matching of queries to ground-truth boxes on the device (the reference: models/utils/ops.py:116-117 moves the
// cost matrix to the host and calls scipy.optimize.linear_sum_assignment per image, a synchronisation in the middle of
// every training step and four times per step with the auxiliary losses).
//
// Same algorithm as SciPy's rectangular_lsap (Crouse's shortest-augmenting-path variant of Jonker-Volgenant), same
// fp64 arithmetic on the fp32 costs, same scan order and tie rules (the `remaining` list starts reversed, a strictly
// smaller path cost wins, an equal one wins only if its column is still unassigned), so the assignment is the one the
// reference computes, not just one of equal cost.  One warp per (layer, image) problem: the lanes stride over the
// remaining columns, the arg-min is a shuffle reduction with SciPy's tie order as the comparator.  The smaller side of
// the matrix plays the rows (SciPy transposes when there are more rows than columns); the sub-matrix of the image is
// staged in shared memory in that orientation when it fits.
#include <math_constants.h>

#include "common.cuh"

namespace tamtr {

constexpr int kAsgThreads = 128;

struct AsgKey {
    double val;
    int free_pos;   // position `it` in the remaining list if the column is unassigned, else -1
    int pos;        // position `it`
};

// does b come before a in SciPy's sequential scan result?  (lowest value; among equal values the LAST unassigned
// column encountered if there is one, otherwise the FIRST column encountered)
__device__ __forceinline__ bool asg_better(const AsgKey &b, const AsgKey &a) {
    if (b.pos < 0) return false;
    if (a.pos < 0) return true;
    if (b.val != a.val) return b.val < a.val;
    if (b.free_pos >= 0 || a.free_pos >= 0) return b.free_pos > a.free_pos;
    return b.pos < a.pos;
}

__global__ void __launch_bounds__(kAsgThreads)
lsap_kernel(const float *__restrict__ C, const int *__restrict__ gt_start, const int *__restrict__ out_start,
            long long *__restrict__ out_q, long long *__restrict__ out_g, int bs, int nq, int c_cols, int padded,
            int max_cols, int stage_elems, long long out_layer_stride, const int *__restrict__ count,
            int *__restrict__ match) {
    extern __shared__ unsigned char smem[];
    const int b = blockIdx.x, layer = blockIdx.y;
    // two ways to describe the images' ground truths: prefix sums (pair lists out) or per-image counts of a padded batch
    // (fixed-shape output: match[layer, b, q] = ground-truth index of the image or -1)
    const int g0 = count ? 0 : gt_start[b];
    const int ng = count ? min(count[b], c_cols) : gt_start[b + 1] - g0;
    if (match)
        for (int q = threadIdx.x; q < nq; q += kAsgThreads) match[((size_t)layer * bs + b) * nq + q] = -1;
    if (ng <= 0) return;
    const int total_gt = c_cols;                                                   // row pitch of C
    const float *Cb = C + ((size_t)layer * bs + b) * (size_t)nq * total_gt + (padded ? 0 : g0);   // Cb[q * pitch + g]
    const bool rows_are_gt = ng < nq;       // SciPy transposes only when there are MORE rows (queries) than columns
    const int nr = rows_are_gt ? ng : nq, nc = rows_are_gt ? nq : ng;

    // shared-memory carve-up (sizes fixed by max_cols = max(nq, max ng))
    double *spc = reinterpret_cast<double *>(smem);                 // shortest path cost per column
    double *v = spc + max_cols;
    double *u = v + max_cols;                                       // rows <= max_cols
    int *path = reinterpret_cast<int *>(u + max_cols);
    int *row4col = path + max_cols;
    int *col4row = row4col + max_cols;
    int *remaining = col4row + max_cols;
    unsigned char *SC = reinterpret_cast<unsigned char *>(remaining + max_cols);
    unsigned char *SR = SC + max_cols;
    float *stage = reinterpret_cast<float *>(SR + max_cols + ((8 - (2 * max_cols) % 8) % 8));   // 4-byte aligned
    const bool staged = (long)nr * nc <= stage_elems;

    for (int i = threadIdx.x; i < nc; i += kAsgThreads) { v[i] = 0.0; row4col[i] = -1; }
    for (int i = threadIdx.x; i < nr; i += kAsgThreads) { u[i] = 0.0; col4row[i] = -1; }
    if (staged) {                                                    // stage[i * nc + j] = cost(row i, column j)
        for (int e = threadIdx.x; e < nq * ng; e += kAsgThreads) {
            const int q = e / ng, g = e - q * ng;
            const float c = Cb[(size_t)q * total_gt + g];
            if (rows_are_gt) stage[g * nc + q] = c; else stage[q * nc + g] = c;
        }
    }
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;

    for (int cur = 0; cur < nr; ++cur) {
        for (int j = lane; j < nc; j += 32) { spc[j] = CUDART_INF; SC[j] = 0; remaining[j] = nc - j - 1; }
        for (int i = lane; i < nr; i += 32) SR[i] = 0;
        __syncwarp();
        int num_remaining = nc, sink = -1, i = cur;
        double min_val = 0.0;
        while (sink < 0) {
            if (lane == 0) SR[i] = 1;
            const double ui = u[i];
            AsgKey best{CUDART_INF, -1, -1};
            for (int it = lane; it < num_remaining; it += 32) {
                const int j = remaining[it];
                const float cf = staged ? stage[i * nc + j]
                                        : (rows_are_gt ? Cb[(size_t)j * total_gt + i] : Cb[(size_t)i * total_gt + j]);
                const double r = min_val + (double)cf - ui - v[j];
                double s = spc[j];
                if (r < s) { path[j] = i; spc[j] = r; s = r; }
                const AsgKey k{s, row4col[j] < 0 ? it : -1, it};
                if (asg_better(k, best)) best = k;
            }
            // warp arg-min with SciPy's tie order, as integer reductions (redux.sync): the fp64 value through its
            // order-preserving 64-bit image (high word, then low word among the lanes that tie on the high word), then
            // among the lanes at the minimum the largest unassigned position, else the smallest position
            unsigned long long ord = ~0ull;
            if (best.pos >= 0) {
                const long long bits = __double_as_longlong(best.val + 0.0);          // (-0.0 -> +0.0)
                ord = bits < 0 ? ~(unsigned long long)bits : (unsigned long long)bits ^ 0x8000000000000000ull;
            }
            const unsigned hi = (unsigned)(ord >> 32), lo = (unsigned)ord;
            const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
            const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
            const bool at_min = best.pos >= 0 && hi == mhi && lo == mlo;
            const int mfree = __reduce_max_sync(0xffffffffu, at_min ? best.free_pos : -1);
            const int mpos = mfree >= 0 ? mfree : __reduce_min_sync(0xffffffffu, at_min ? best.pos : 0x7fffffff);
            const unsigned long long mord = ((unsigned long long)mhi << 32) | mlo;
            best.pos = mord == ~0ull ? -1 : mpos;
            best.val = __longlong_as_double((long long)((mord >> 63) ? mord ^ 0x8000000000000000ull : ~mord));
            min_val = best.val;
            if (best.pos < 0 || min_val == CUDART_INF) { sink = -2; break; }    // infeasible (cannot happen: costs are finite)
            const int j = remaining[best.pos];
            __syncwarp();
            if (row4col[j] < 0) sink = j; else i = row4col[j];
            if (lane == 0) {
                SC[j] = 1;
                remaining[best.pos] = remaining[--num_remaining];
            } else {
                --num_remaining;
            }
            __syncwarp();
        }
        if (sink < 0) break;
        // dual update
        for (int r = lane; r < nr; r += 32) {
            if (r == cur) u[r] += min_val;
            else if (SR[r]) u[r] += min_val - spc[col4row[r]];
        }
        for (int j = lane; j < nc; j += 32)
            if (SC[j]) v[j] -= min_val - spc[j];
        __syncwarp();
        // augment along the path (sequential, short)
        if (lane == 0) {
            int j = sink;
            while (true) {
                const int r = path[j];
                row4col[j] = r;
                const int prev = col4row[r];
                col4row[r] = j;
                j = prev;
                if (r == cur) break;
            }
        }
        __syncwarp();
    }

    if (match) {
        __syncwarp();
        for (int q = lane; q < nq; q += 32) {
            const int g = rows_are_gt ? row4col[q] : col4row[q];
            if (g >= 0) match[((size_t)layer * bs + b) * nq + q] = g;
        }
        return;
    }
    // pairs in ascending query order (what linear_sum_assignment returns), gt indices made global like ops.py:120
    long long *oq = out_q + layer * out_layer_stride + out_start[b];
    long long *og = out_g + layer * out_layer_stride + out_start[b];
    int base = 0;
    for (int q0 = 0; q0 < nq; q0 += 32) {
        const int q = q0 + lane;
        int g = -1;
        if (q < nq) g = rows_are_gt ? row4col[q] : col4row[q];
        const unsigned m = __ballot_sync(0xffffffffu, g >= 0);
        if (g >= 0) {
            const int slot = base + __popc(m & ((1u << lane) - 1));
            oq[slot] = q;
            og[slot] = g0 + g;
        }
        base += __popc(m);
    }
}

}  // namespace tamtr

using namespace tamtr;

extern "C" int tamtr_linear_sum_assignment(const float *C, const int *gt_start_dev, const int *out_start_dev,
                                           long long *out_q, long long *out_g, int n_layers, int bs, int nq,
                                           int c_cols, int padded, int max_gt, long long out_layer_stride,
                                           void *stream) {
    TAMTR_CHECK_ARG(C && gt_start_dev && out_start_dev && out_q && out_g, TAMTR_E_BADARG, "linear_sum_assignment: null pointer");
    TAMTR_CHECK_ARG(n_layers > 0 && bs > 0 && nq > 0 && c_cols > 0 && max_gt > 0 && max_gt <= c_cols, TAMTR_E_BADARG,
                    "linear_sum_assignment: bad sizes");
    TAMTR_CHECK_ARG(bs <= 65535 && n_layers <= 65535, TAMTR_E_UNSUPPORTED, "linear_sum_assignment: grid too large");
    const int max_cols = nq > max_gt ? nq : max_gt;
    const size_t fixed = (size_t)max_cols * (3 * sizeof(double) + 4 * sizeof(int) + 2) + 16;
    int dev = 0, max_smem = 0;
    TAMTR_CUDA_OK(cudaGetDevice(&dev));
    TAMTR_CUDA_OK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    TAMTR_CHECK_ARG(fixed + 1024 <= (size_t)max_smem, TAMTR_E_UNSUPPORTED,
                    "linear_sum_assignment: %d columns need more shared memory than the device has", max_cols);
    size_t want = (size_t)nq * max_gt * sizeof(float);               // the whole sub-matrix of the largest image
    if (fixed + want > (size_t)max_smem) want = 0;                   // too big: read the costs from global memory (L2)
    const size_t smem = fixed + want;
    TAMTR_CUDA_OK(cudaFuncSetAttribute(lsap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lsap_kernel<<<dim3(bs, n_layers), kAsgThreads, smem, (cudaStream_t)stream>>>(
        C, gt_start_dev, out_start_dev, out_q, out_g, bs, nq, c_cols, padded, max_cols, (int)(want / sizeof(float)),
        out_layer_stride, nullptr, nullptr);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_linear_sum_assignment_padded(const float *C, const int *count_dev, int *match, int n_layers, int bs,
                                                  int nq, int max_gt, void *stream) {
    TAMTR_CHECK_ARG(C && count_dev && match, TAMTR_E_BADARG, "linear_sum_assignment_padded: null pointer");
    TAMTR_CHECK_ARG(n_layers > 0 && bs > 0 && nq > 0 && max_gt > 0, TAMTR_E_BADARG, "linear_sum_assignment_padded: bad sizes");
    TAMTR_CHECK_ARG(bs <= 65535 && n_layers <= 65535, TAMTR_E_UNSUPPORTED, "linear_sum_assignment_padded: grid too large");
    const int max_cols = nq > max_gt ? nq : max_gt;
    const size_t fixed = (size_t)max_cols * (3 * sizeof(double) + 4 * sizeof(int) + 2) + 16;
    int dev = 0, max_smem = 0;
    TAMTR_CUDA_OK(cudaGetDevice(&dev));
    TAMTR_CUDA_OK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    TAMTR_CHECK_ARG(fixed + 1024 <= (size_t)max_smem, TAMTR_E_UNSUPPORTED,
                    "linear_sum_assignment_padded: %d columns need more shared memory than the device has", max_cols);
    size_t want = (size_t)nq * max_gt * sizeof(float);
    if (fixed + want > (size_t)max_smem) want = 0;
    const size_t smem = fixed + want;
    TAMTR_CUDA_OK(cudaFuncSetAttribute(lsap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lsap_kernel<<<dim3(bs, n_layers), kAsgThreads, smem, (cudaStream_t)stream>>>(
        C, nullptr, nullptr, nullptr, nullptr, bs, nq, max_gt, 1, max_cols, (int)(want / sizeof(float)), 0, count_dev, match);
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}
