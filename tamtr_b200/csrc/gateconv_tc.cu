// BTA-PAN text-guided projection: the 3x3 `proj_conv` of MaxSigmoidAttnBlock as an implicit GEMM on the 5th-gen
// tensor cores with BatchNorm (folded affine) and the max-sigmoid text gate fused into the epilogue.
//
// Replaces /root/reference ultralytics/nn/extra_modules/block.py:222-225 for bf16 activations:
//   y[b, co, h, w] = ( sum_{dy,dx,ci} W[co, ci, dy, dx] * x[b, ci, h+dy-1, w+dx-1] * s[co] + t[co] ) * aw[b, co/hc, h, w]
// (s, t = BatchNorm2d affine of `proj_conv.bn`; aw = the gate of tamtr_max_sigmoid_*_forward).
//
// This is the tensor-bound half of "kernel 2" (DESIGN.md section 3): AI = 2*9*Cin*Cout / (2*(Cin+Cout)) ~ 1150 flop/B
// at C = 256, far above the ridge.  Layout: activations channels-last ([B, H, W, C] bf16), weights [Cout, 3, 3, Cin]
// bf16, so both UMMA operands are K-major and every smem row is one pixel's (one filter's) 64-channel slice = 128 B.
//
// CTA tile = 256 output pixels (a TW x 256/TW patch of one image) x all Cout channels, as two M=128 halves whose
// fp32 accumulators fill TMEM (2 x 256 columns).  K loop = 9 taps x Cin/64 blocks.  Per K block:
//   warp 0   TMA producer : ONE cp.async.bulk.tensor.4d box (64 ch, TW, 256/TW, 1) at (c0, x0+dx-1, y0+dy-1, b) -- the
//                           halo is the tensor map's out-of-bounds zero fill, there is no im2col buffer and no padding
//                           branch -- plus one 2-D box (64, Cout) of the weights; 128B swizzle; 3-stage 64 KB ring
//   warp 1   MMA issuer   : 2 halves x 4 x tcgen05.mma.cta_group::1.kind::f16 (M=128, N=Cout, K=16), the weight tile in
//                           shared memory is read by both halves (so smem fill traffic per flop is that of a 256x256 tile)
//   warps 2-9 epilogue    : tcgen05.ld 32 lanes x 32 columns, y = (acc*s + t) * gate, bf16 pack, 16-byte stores
//                           (each thread owns one pixel = one contiguous Cout*2-byte row of the channels-last output)
#include "tc_ptx.cuh"

namespace tamtr {

constexpr int kCvStages = 3;
constexpr int kCvKBlk = 64;                          // channels per K block: 64 bf16 = one 128-byte swizzled row
constexpr int kCvTilePix = 256;                      // output pixels per CTA tile (two UMMA M=128 halves)
constexpr int kCvABytes = kCvTilePix * kCvKBlk * 2;  // 32 KB
constexpr int kCvMaxCout = 256;
constexpr int kCvBBytes = kCvMaxCout * kCvKBlk * 2;  // 32 KB
constexpr int kCvThreads = 320;                      // producer warp, MMA warp, 8 epilogue warps
constexpr int kCvTmemCols = 512;

struct CvSmem {
    alignas(1024) uint8_t a[kCvStages][kCvABytes];
    alignas(1024) uint8_t b[kCvStages][kCvBBytes];
    alignas(16) float scale[kCvMaxCout];
    alignas(16) float shift[kCvMaxCout];
    alignas(8) uint64_t full[kCvStages];
    uint64_t empty[kCvStages];
    uint64_t acc_full;
    uint64_t acc_empty;
    uint32_t tmem_base;
};

// K-major operand, SWIZZLE_128B: rows of 128 B, 8-row swizzle atoms of 1024 B stacked every 1024 B (SBO); LBO unused.
__device__ __forceinline__ uint64_t make_desc_k128(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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

struct CvGeom {
    int B, H, W, Cin, Cout, nh;
    int TW, TH;            // tile = TW x TH pixels, TW * TH = 256, TW a power of two
    int tiles_x, tiles_y;  // per image
};

__global__ void __launch_bounds__(kCvThreads, 1)
gate_conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                       const float *__restrict__ bn_scale, const float *__restrict__ bn_shift,
                       const float *__restrict__ gate, __nv_bfloat16 *__restrict__ y, const CvGeom g) {
    extern __shared__ uint8_t smem_raw[];
    CvSmem &sm = *reinterpret_cast<CvSmem *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_img = g.tiles_x * g.tiles_y;
    const int n_tiles = g.B * tiles_img;
    const int kc_per_tap = g.Cin / kCvKBlk;
    const int n_kb = 9 * kc_per_tap;
    const uint32_t stage_bytes = (uint32_t)kCvABytes + (uint32_t)g.Cout * kCvKBlk * 2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kCvStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
        mbar_init(&sm.acc_full, 1);
        mbar_init(&sm.acc_empty, kCvThreads - 64);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)),
                     "n"(kCvTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int c = threadIdx.x; c < g.Cout; c += kCvThreads) {
        sm.scale[c] = __ldg(bn_scale + c);
        sm.shift[c] = __ldg(bn_shift + c);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = sm.tmem_base;

    if (warp == 0) {
        // ===== TMA producer
        if (lane == 0) {
            uint32_t kbg = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int b = t / tiles_img, r = t - b * tiles_img;
                const int x0 = (r % g.tiles_x) * g.TW, y0 = (r / g.tiles_x) * g.TH;
                for (int tap = 0; tap < 9; ++tap) {
                    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                    for (int kc = 0; kc < kc_per_tap; ++kc, ++kbg) {
                        const int s = kbg % kCvStages;
                        mbar_wait(&sm.empty[s], ((kbg / kCvStages) & 1) ^ 1);
                        mbar_expect_tx(&sm.full[s], stage_bytes);
                        tma_load_4d(sm.a[s], &tmap_x, &sm.full[s], kc * kCvKBlk, x0 + dx, y0 + dy, b);
                        tma_load_2d(sm.b[s], &tmap_w, &sm.full[s], tap * g.Cin + kc * kCvKBlk, 0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer
        if (lane == 0) {
            // instruction descriptor: D = f32, A = B = bf16, both K-major, N = Cout, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(g.Cout >> 3) << 17) | ((128u >> 4) << 24);
            uint32_t kbg = 0;
            int it = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
                mbar_wait(&sm.acc_empty, (it & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = 0; kb < n_kb; ++kb, ++kbg) {
                    const int s = kbg % kCvStages;
                    mbar_wait(&sm.full[s], (kbg / kCvStages) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = smem_u32(sm.a[s]), b_addr = smem_u32(sm.b[s]);
#pragma unroll
                    for (int k = 0; k < kCvKBlk / 16; ++k) {
                        const uint64_t bd = make_desc_k128(b_addr + k * 32);
                        const uint32_t acc = (kb | k) ? 1u : 0u;
                        umma_f16(tmem_base, make_desc_k128(a_addr + k * 32), bd, idesc, acc);
                        umma_f16(tmem_base + kCvMaxCout, make_desc_k128(a_addr + kCvABytes / 2 + k * 32), bd, idesc, acc);
                    }
                    umma_commit(&sm.empty[s]);
                }
                umma_commit(&sm.acc_full);
            }
        }
    } else {
        // ===== epilogue: warp w owns TMEM lanes 32*(w%4).. of half (w-2)/4
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const int p = half * 128 + quarter * 32 + lane;          // pixel slot inside the 256-pixel tile
        const int px_in = p & (g.TW - 1), py_in = p / g.TW;
        const int hc = g.Cout / g.nh;
        const size_t HW = (size_t)g.H * g.W;
        int it = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const int b = t / tiles_img, r = t - b * tiles_img;
            const int px = (r % g.tiles_x) * g.TW + px_in, py = (r / g.tiles_x) * g.TH + py_in;
            const bool live = px < g.W && py < g.H;
            const size_t pix = (size_t)py * g.W + px;
            mbar_wait(&sm.acc_full, it & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + half * kCvMaxCout;
            __nv_bfloat16 *dst = y + ((size_t)b * HW + pix) * g.Cout;
            for (int c0 = 0; c0 < g.Cout; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
                float gt = 1.0f;
                if (gate != nullptr && live) gt = __ldg(gate + ((size_t)b * g.nh + c0 / hc) * HW + pix);
                uint4 o[4];
                uint32_t *ow = reinterpret_cast<uint32_t *>(o);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 s4 = *reinterpret_cast<const float4 *>(&sm.scale[c0 + j]);
                    const float4 t4 = *reinterpret_cast<const float4 *>(&sm.shift[c0 + j]);
                    const float f0 = fmaf(__uint_as_float(v[j + 0]), s4.x, t4.x) * gt;
                    const float f1 = fmaf(__uint_as_float(v[j + 1]), s4.y, t4.y) * gt;
                    const float f2 = fmaf(__uint_as_float(v[j + 2]), s4.z, t4.z) * gt;
                    const float f3 = fmaf(__uint_as_float(v[j + 3]), s4.w, t4.w) * gt;
                    __nv_bfloat162 lo = __floats2bfloat162_rn(f0, f1), hi = __floats2bfloat162_rn(f2, f3);
                    ow[j / 2] = *reinterpret_cast<uint32_t *>(&lo);
                    ow[j / 2 + 1] = *reinterpret_cast<uint32_t *>(&hi);
                }
                if (live) {
                    uint4 *d4 = reinterpret_cast<uint4 *>(dst + c0);
#pragma unroll
                    for (int q = 0; q < 4; ++q) d4[q] = o[q];
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&sm.acc_empty);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kCvTmemCols));
    }
}

// ---- layout change NCHW -> channels-last (one pass, HBM-bound): 64 channels x 64 pixels per CTA through shared memory.
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const T *__restrict__ x, T *__restrict__ y, int C, int HW) {
    __shared__ T tile[64][66];
    const int b = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const T *xb = x + (size_t)b * C * HW;
    T *yb = y + (size_t)b * C * HW;
    for (int i = ty; i < 64; i += 8) {
        const int c = c0 + i;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int pp = p0 + tx + 32 * j;
            if (c < C && pp < HW) tile[i][tx + 32 * j] = xb[(size_t)c * HW + pp];
        }
    }
    __syncthreads();
    for (int i = ty; i < 64; i += 8) {
        const int pp = p0 + i;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = c0 + tx + 32 * j;
            if (c < C && pp < HW) yb[(size_t)pp * C + c] = tile[tx + 32 * j][i];
        }
    }
}

}  // namespace tamtr

using namespace tamtr;

extern "C" int tamtr_nchw_to_nhwc(const void *x, void *y, int dtype, int B, int C, int HW, void *stream) {
    TAMTR_CHECK_ARG(x && y, TAMTR_E_BADARG, "nchw_to_nhwc: null pointer");
    TAMTR_CHECK_ARG(B > 0 && C > 0 && HW > 0 && B <= 65535, TAMTR_E_BADARG, "nchw_to_nhwc: bad sizes");
    TAMTR_CHECK_ARG(dtype == TAMTR_F32 || dtype == TAMTR_BF16, TAMTR_E_UNSUPPORTED, "nchw_to_nhwc: dtype %d", dtype);
    const dim3 grid((HW + 63) / 64, (C + 63) / 64, B);
    cudaStream_t st = (cudaStream_t)stream;
    {
        KernelTimer timer(K_NHWC, st);
        if (dtype == TAMTR_BF16)
            nchw_to_nhwc_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x, (__nv_bfloat16 *)y, C, HW);
        else
            nchw_to_nhwc_kernel<<<grid, 256, 0, st>>>((const float *)x, (float *)y, C, HW);
    }
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int tamtr_gate_conv3x3_tc_forward(const void *x_nhwc, const void *w_ohwi, const float *bn_scale,
                                             const float *bn_shift, const float *gate, void *y_nhwc, int B, int H, int W,
                                             int Cin, int Cout, int nh, void *stream) {
    TAMTR_CHECK_ARG(x_nhwc && w_ohwi && bn_scale && bn_shift && y_nhwc, TAMTR_E_BADARG, "gate_conv3x3_tc: null pointer");
    TAMTR_CHECK_ARG(B > 0 && H > 0 && W > 0 && nh > 0, TAMTR_E_BADARG, "gate_conv3x3_tc: non-positive size");
    TAMTR_CHECK_ARG(Cin > 0 && Cin % kCvKBlk == 0, TAMTR_E_UNSUPPORTED, "gate_conv3x3_tc: Cin = %d must be a multiple of 64",
                    Cin);
    TAMTR_CHECK_ARG(Cout >= 32 && Cout <= kCvMaxCout && Cout % 32 == 0, TAMTR_E_UNSUPPORTED,
                    "gate_conv3x3_tc: Cout = %d must be a multiple of 32 in [32, 256]", Cout);
    TAMTR_CHECK_ARG(Cout % nh == 0 && (Cout / nh) % 32 == 0, TAMTR_E_UNSUPPORTED,
                    "gate_conv3x3_tc: channels per head (%d / %d) must be a multiple of 32", Cout, nh);
    TAMTR_CHECK_ARG((((uintptr_t)x_nhwc | (uintptr_t)w_ohwi | (uintptr_t)y_nhwc) & 15) == 0, TAMTR_E_BADARG,
                    "gate_conv3x3_tc: pointers must be 16-byte aligned");
    EncodeTiledFn encode = get_encode();
    TAMTR_CHECK_ARG(encode != nullptr, TAMTR_E_NODEVICE, "gate_conv3x3_tc: cuTensorMapEncodeTiled unavailable");

    // tile shape: the power-of-two TW x 256/TW patch that covers the map with the fewest tiles
    CvGeom g{B, H, W, Cin, Cout, nh, 16, 16, 0, 0};
    long best = -1;
    for (int tw = 8; tw <= 32; tw *= 2) {
        const int th = kCvTilePix / tw;
        const long n = (long)((W + tw - 1) / tw) * ((H + th - 1) / th);
        if (best < 0 || n < best || (n == best && tw == 16)) {
            best = n;
            g.TW = tw;
            g.TH = th;
        }
    }
    g.tiles_x = (W + g.TW - 1) / g.TW;
    g.tiles_y = (H + g.TH - 1) / g.TH;

    CUtensorMap tmap_x, tmap_w;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        const cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
        const cuuint32_t box[4] = {(cuuint32_t)kCvKBlk, (cuuint32_t)g.TW, (cuuint32_t)g.TH, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        const CUresult cr = encode(&tmap_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(x_nhwc), dims, strides,
                                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TAMTR_CHECK_ARG(cr == CUDA_SUCCESS, TAMTR_E_BADARG, "gate_conv3x3_tc: tensor map (x) failed (%d)", (int)cr);
    }
    {
        const cuuint64_t dims[2] = {(cuuint64_t)9 * Cin, (cuuint64_t)Cout};
        const cuuint64_t strides[1] = {(cuuint64_t)9 * Cin * 2};
        const cuuint32_t box[2] = {(cuuint32_t)kCvKBlk, (cuuint32_t)Cout};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult cr = encode(&tmap_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(w_ohwi), dims, strides,
                                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TAMTR_CHECK_ARG(cr == CUDA_SUCCESS, TAMTR_E_BADARG, "gate_conv3x3_tc: tensor map (w) failed (%d)", (int)cr);
    }

    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    const long n_tiles = (long)B * g.tiles_x * g.tiles_y;
    // persistent CTAs, one per SM; equalise the number of tiles per CTA so the last wave is not ragged
    const long waves = (n_tiles + n_sm - 1) / n_sm;
    const int grid = (int)((n_tiles + waves - 1) / waves);
    const size_t smem = sizeof(CvSmem) + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        TAMTR_CUDA_OK(cudaFuncSetAttribute(gate_conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    cudaStream_t st = (cudaStream_t)stream;
    {
        KernelTimer timer(K_GATE_CONV_TC, st);
        gate_conv3x3_tc_kernel<<<grid, kCvThreads, smem, st>>>(tmap_x, tmap_w, bn_scale, bn_shift, gate,
                                                               (__nv_bfloat16 *)y_nhwc, g);
    }
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}
