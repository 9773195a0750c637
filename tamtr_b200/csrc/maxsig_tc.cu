// BTA-PAN text-image attention gate on the 5th-gen tensor cores: tcgen05.mma + TMEM accumulators + TMA-staged tiles.
//
// Replaces /root/reference ultralytics/nn/extra_modules/block.py:216-220 for bf16 activations:
//   logits[pix, n] = sum_c embed[b, m*32 + c, pix] * guide[b, n, m, c]        (einsum "bmchw,bnmc->bmhwn")
//   aw[b, m, pix]  = sigmoid( max_n logits[pix, n] / sqrt(32) + bias[m] )
// as a warp-specialised persistent-per-(b,m) kernel:
//   warp 0   TMA producer : cp.async.bulk.tensor.2d loads the 32-channel x 128-pixel tile of `embed` (NCHW: pixels are
//                           contiguous, so the tile is an MN-major A operand) as two 64-pixel boxes, 128B-swizzled,
//                           into an 8-stage shared-memory ring (mbarrier complete_tx)
//   warp 1   MMA issuer   : one elected lane issues 2 x tcgen05.mma.cta_group::1.kind::f16 (M=128, N=Npad, K=16 each;
//                           K = hc = 32) per tile into one of two TMEM accumulator stages; tcgen05.commit releases the
//                           smem stage and publishes the accumulator
//   warps 2-5 epilogue    : tcgen05.ld (32 lanes x 32b x 16 columns) -> running max / arg-max over the N real text
//                           tokens -> scale, bias, sigmoid -> coalesced fp32 store (+ uint8 arg-max for the backward)
// The guide block of (b, m) ([N, 32], K-major B operand) is converted to bf16 and laid out as UMMA core matrices
// (no swizzle) in shared memory once per CTA.
//
// Roofline: the op is HBM-bound (reads B*C*HW bf16 once; AI = 2*N/2 = N flop/B <= 80 << ridge 255), so tensor-pipe
// utilisation is structurally low -- see DESIGN.md section 3 "Kernel 2"; the fused 3x3-conv variant is the tensor-bound one.
#include "tc_ptx.cuh"

namespace tamtr {

constexpr int kTcTileM = 128;                       // pixels per tile = UMMA M = TMEM lanes
constexpr int kTcHc = 32;                           // channels per head = K
constexpr int kTcStages = 8;                        // smem ring depth: 2 CTAs/SM x 8 x 8 KB = 128 KB in flight per SM
constexpr int kTcAccStages = 2;                     // TMEM accumulator double buffer
constexpr int kTcAccCols = 128;                     // columns reserved per accumulator stage (Npad <= 128)
constexpr int kTcABytes = kTcHc * kTcTileM * 2;     // 8192: two 64-pixel boxes of 32 rows x 128 B
constexpr int kTcMaxNpad = 128;
constexpr int kTcBBytes = (kTcMaxNpad / 8) * 512;   // core-matrix layout: 512 B per 8 rows (4 K-chunks x 128 B)
constexpr int kTcThreads = 192;

struct TcSmem {
    alignas(1024) uint8_t a[kTcStages][kTcABytes];
    alignas(128) uint8_t b[kTcBBytes];
    alignas(8) uint64_t full[kTcStages];
    uint64_t empty[kTcStages];
    uint64_t acc_full[kTcAccStages];
    uint64_t acc_empty[kTcAccStages];
    uint32_t tmem_base;
};

// A: MN-major, SWIZZLE_128B.  64 pixels (128 B) x 8 channel rows = one 1024 B atom; atoms stack along K every 1024 B
// (SBO), the second 64-pixel half of the tile sits 4096 B further (LBO).  Field layout: cute/arch/mma_sm100_desc.hpp.
__device__ __forceinline__ uint64_t make_desc_a(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}
// B: K-major, no swizzle.  Core matrix = 8 rows x 16 B contiguous (128 B); next K chunk +128 B (LBO), next 8 rows
// +512 B (SBO).
__device__ __forceinline__ uint64_t make_desc_b(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46);
}

__global__ void __launch_bounds__(kTcThreads, 2)
gate_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ guide,
                   const float *__restrict__ bias, float *__restrict__ aw, uint8_t *__restrict__ amax, int nh, int HW,
                   int N, int Npad) {
    extern __shared__ uint8_t smem_raw[];
    TcSmem &sm = *reinterpret_cast<TcSmem *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bm = blockIdx.y;          // b * nh + m
    const int m = bm % nh, b = bm / nh;
    const int n_tiles = (HW + kTcTileM - 1) / kTcTileM;

    // ---- one-time setup
    if (threadIdx.x == 0) {
        for (int s = 0; s < kTcStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
        for (int a = 0; a < kTcAccStages; ++a) { mbar_init(&sm.acc_full[a], 1); mbar_init(&sm.acc_empty[a], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: 2 accumulator stages x 128 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)),
                     "n"(kTcAccStages * kTcAccCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // guide block of (b, m) -> bf16 core matrices; rows n >= N are zero
    for (int i = threadIdx.x; i < Npad * kTcHc; i += kTcThreads) {
        const int n = i / kTcHc, c = i % kTcHc;
        const float v = n < N ? __ldg(guide + (((size_t)b * N + n) * nh + m) * kTcHc + c) : 0.0f;
        const int off = (n >> 3) * 512 + (c >> 3) * 128 + (n & 7) * 16 + (c & 7) * 2;
        *reinterpret_cast<__nv_bfloat16 *>(sm.b + off) = __float2bfloat16_rn(v);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // st.shared -> visible to the tensor-core proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = sm.tmem_base;

    if (warp == 0) {
        // ===== TMA producer
        if (lane == 0) {
            int it = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
                const int s = it % kTcStages;
                const uint32_t ph = (it / kTcStages) & 1;
                mbar_wait(&sm.empty[s], ph ^ 1);
                mbar_expect_tx(&sm.full[s], kTcABytes);
                const int row = bm * kTcHc;                       // first channel row of this head in [B*C, HW]
                tma_load_2d(sm.a[s], &tmap, &sm.full[s], t * kTcTileM, row);
                tma_load_2d(sm.a[s] + 4096, &tmap, &sm.full[s], t * kTcTileM + 64, row);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer
        if (lane == 0) {
            // instruction descriptor: D=f32, A=B=bf16, A MN-major, B K-major, N = Npad, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(Npad >> 3) << 17) |
                                   ((uint32_t)(kTcTileM >> 4) << 24);
            const uint32_t b_addr = smem_u32(sm.b);
            int it = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
                const int s = it % kTcStages, a = it % kTcAccStages;
                mbar_wait(&sm.acc_empty[a], ((it / kTcAccStages) & 1) ^ 1);
                mbar_wait(&sm.full[s], (it / kTcStages) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(sm.a[s]);
                const uint32_t d = tmem_base + a * kTcAccCols;
#pragma unroll
                for (int k = 0; k < kTcHc / 16; ++k)
                    umma_f16(d, make_desc_a(a_addr + k * 2048), make_desc_b(b_addr + k * 256), idesc, k);
                umma_commit(&sm.empty[s]);       // smem stage free once the MMAs have read it
                umma_commit(&sm.acc_full[a]);    // accumulator ready for the epilogue
            }
        }
    } else {
        // ===== epilogue: warps 2..5 own TMEM lane quarters (warp % 4)
        const int quarter = warp & 3;
        const float inv = 1.0f / sqrtf((float)kTcHc);
        const float bs = __ldg(bias + m);
        int it = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const int a = it % kTcAccStages;
            mbar_wait(&sm.acc_full[a], (it / kTcAccStages) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + a * kTcAccCols;
            float best = -INFINITY;
            int arg = 0;
            for (int c0 = 0; c0 < Npad; c0 += 16) {
                uint32_t r[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr + c0));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float v = __uint_as_float(r[j]);
                    if (c0 + j < N && v > best) { best = v; arg = c0 + j; }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&sm.acc_empty[a]);
            const int pix = t * kTcTileM + quarter * 32 + lane;
            if (pix < HW) {
                const float z = best * inv + bs;
                const size_t o = (size_t)bm * HW + pix;
                aw[o] = 1.0f / (1.0f + expf(-z));
                amax[o] = (uint8_t)arg;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "n"(kTcAccStages * kTcAccCols));
    }
}

}  // namespace tamtr

using namespace tamtr;

extern "C" int tamtr_max_sigmoid_tc_forward(const void *embed_bf16, const float *guide, const float *bias, float *aw,
                                            uint8_t *amax, int B, int nh, int hc, int HW, int N, void *stream) {
    TAMTR_CHECK_ARG(embed_bf16 && guide && bias && aw && amax, TAMTR_E_BADARG, "max_sigmoid_tc_forward: null pointer");
    TAMTR_CHECK_ARG(B > 0 && nh > 0 && HW > 0 && N > 0, TAMTR_E_BADARG, "max_sigmoid_tc: non-positive size");
    TAMTR_CHECK_ARG(hc == kTcHc, TAMTR_E_UNSUPPORTED, "max_sigmoid_tc: head channels %d unsupported (32)", hc);
    TAMTR_CHECK_ARG(N <= kTcMaxNpad, TAMTR_E_UNSUPPORTED, "max_sigmoid_tc: %d text tokens > %d", N, kTcMaxNpad);
    TAMTR_CHECK_ARG(HW % 8 == 0, TAMTR_E_UNSUPPORTED, "max_sigmoid_tc: H*W = %d must be a multiple of 8 (TMA row stride)",
                    HW);
    TAMTR_CHECK_ARG(((uintptr_t)embed_bf16 & 15) == 0, TAMTR_E_BADARG, "max_sigmoid_tc: embed must be 16-byte aligned");
    TAMTR_CHECK_ARG((long)B * nh <= 65535, TAMTR_E_UNSUPPORTED, "max_sigmoid_tc: grid too large");
    EncodeTiledFn encode = get_encode();
    TAMTR_CHECK_ARG(encode != nullptr, TAMTR_E_NODEVICE, "max_sigmoid_tc: cuTensorMapEncodeTiled unavailable");

    // embed viewed as a [B*C, HW] row-major bf16 matrix; box = 64 pixels x 32 channel rows, 128B swizzle, zero fill
    CUtensorMap tmap;
    const cuuint64_t dims[2] = {(cuuint64_t)HW, (cuuint64_t)B * nh * hc};
    const cuuint64_t strides[1] = {(cuuint64_t)HW * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)kTcHc};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(embed_bf16), dims, strides,
                               box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TAMTR_CHECK_ARG(cr == CUDA_SUCCESS, TAMTR_E_BADARG, "max_sigmoid_tc: cuTensorMapEncodeTiled failed (%d)", (int)cr);

    const int Npad = ((N + 15) / 16) * 16;
    const int n_tiles = (HW + kTcTileM - 1) / kTcTileM;
    // one wave of (at most) 2 CTAs per SM: each CTA walks the tiles of its (b, m) with stride gridDim.x
    int n_sm = 148;
    n_sm = ::tamtr::sm_count();
    int chunks = (2 * n_sm) / (B * nh);
    if (chunks < 1) chunks = 1;
    if (chunks > n_tiles) chunks = n_tiles;
    const size_t smem = sizeof(TcSmem) + 1024;
    static bool attr_set[64] = {false};          // cudaFuncSetAttribute is per device
    int dev_id = 0;
    TAMTR_CUDA_OK(cudaGetDevice(&dev_id));
    if (dev_id < 0 || dev_id >= 64 || !attr_set[dev_id]) {
        TAMTR_CUDA_OK(cudaFuncSetAttribute(gate_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev_id >= 0 && dev_id < 64) attr_set[dev_id] = true;
    }
    cudaStream_t st = (cudaStream_t)stream;
    {
        KernelTimer timer(K_GATE_TC_FWD, st);
        gate_tc_fwd_kernel<<<dim3(chunks, B * nh), kTcThreads, smem, st>>>(tmap, guide, bias, aw, amax, nh, HW, N, Npad);
    }
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}
