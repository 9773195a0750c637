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
// CTA tile = 256 output pixels (a TW x TH patch of one image, TW*TH = 256) x all Cout channels, as two M=128 halves
// whose fp32 accumulators fill TMEM (2 x 256 columns).  K loop = Cin/64 channel blocks x 3 column taps x 3 row taps.
//   warp 0   TMA producer : per (channel block, dx) ONE cp.async.bulk.tensor.4d box (64 ch, TW, TH+2, 1) at
//                           (c0, x0+dx-1, y0-1, b): the patch with its two halo rows.  The halo and the image border
//                           are the tensor map's out-of-bounds zero fill -- no im2col buffer, no padding branch.  The
//                           three row taps dy of that dx are the SAME shared-memory buffer read at a start address
//                           shifted by dy*TW rows (a multiple of the 1024-byte swizzle atom), so x leaves L2 3.4 times
//                           instead of 9 (the kernel is L2->SM bound otherwise: first version, 64 KB per K block,
//                           ran at 82 % of the measured ~6300 B/clk L2 cap with the tensor pipe 54 % busy).
//                           Per K block one 2-D box (64, Cout) of the weights.  128B swizzle; 3 A slots + 3 B stages.
//   warp 1   MMA issuer   : 2 halves x 4 x tcgen05.mma.cta_group::1.kind::f16 (M=128, N=Cout, K=16); the weight tile in
//                           shared memory is read by both halves
//   warps 2-9 epilogue    : tcgen05.ld 32 lanes x 32 columns, y = (acc*s + t) * gate, bf16 pack, 16-byte stores
//                           (each thread owns one pixel = one contiguous Cout*2-byte row of the channels-last output)
#include <stdlib.h>

#include "tc_ptx.cuh"

namespace tamtr {

constexpr int kCvStages = 3;                         // weight (B) stages
constexpr int kCvASlots = 3;                         // activation (A) slots: one (channel block, dx) patch each
constexpr int kCvKBlk = 64;                          // channels per K block: 64 bf16 = one 128-byte swizzled row
constexpr int kCvTilePix = 256;                      // output pixels per CTA tile (two UMMA M=128 halves)
constexpr int kCvABytes = 40 * 1024;                 // TW*(TH+2) rows of 128 B: 272 / 288 / 320 rows for TW = 8/16/32
constexpr int kCvMaxCout = 256;
constexpr int kCvBBytes = kCvMaxCout * kCvKBlk * 2;  // 32 KB
constexpr int kCvThreads = 320;                      // producer warp, MMA warp, 8 epilogue warps
constexpr int kCvTmemCols = 512;

struct CvSmem {
    alignas(1024) uint8_t a[kCvASlots][kCvABytes];
    alignas(1024) uint8_t b[kCvStages][kCvBBytes];
    alignas(16) float scale[kCvMaxCout];
    alignas(16) float shift[kCvMaxCout];
    alignas(8) uint64_t full[kCvStages];
    uint64_t empty[kCvStages];
    uint64_t a_full[kCvASlots];
    uint64_t a_empty[kCvASlots];
    uint64_t acc_full;
    uint64_t acc_empty;
    uint32_t tmem_base;
};

// K-major operand, SWIZZLE_128B: rows of 128 B, 8-row swizzle atoms of 1024 B stacked every 1024 B (SBO); LBO unused.
__device__ __forceinline__ uint64_t make_desc_k128(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// The same descriptor split into its 32-bit halves: only the low word (start address >> 4, LBO) changes between MMAs,
// so the issue loop adds small constants to it instead of rebuilding 64-bit values (the single issuing thread's
// instruction count is what bounds the MMA rate once the operands are in place).
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr) { return (addr >> 4) | (1u << 16); }
__device__ __forceinline__ uint64_t desc_of(uint32_t lo) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(kDescHi));
    return d;
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
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

// One accumulator row (= one output pixel, Cout channels) from TMEM: y = (acc * s + t) * gate[head], packed to bf16 and
// written as 16-byte vectors into the pixel's contiguous channels-last row.
__device__ __forceinline__ void epilogue_row(uint32_t taddr, __nv_bfloat16 *dst, const float *gate_px, const float *scale,
                                             const float *shift, int Cout, int hc, size_t HW, bool live) {
    for (int c0 = 0; c0 < Cout; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        float gt = 1.0f;
        if (gate_px != nullptr && live) gt = __ldg(gate_px + (size_t)(c0 / hc) * HW);
        uint4 o[4];
        uint32_t *ow = reinterpret_cast<uint32_t *>(o);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 s4 = *reinterpret_cast<const float4 *>(&scale[c0 + j]);
            const float4 t4 = *reinterpret_cast<const float4 *>(&shift[c0 + j]);
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
}

struct CvGeom {
    int B, H, W, Cin, Cout, nh;
    int TW, TH;            // tile = TW x TH pixels, TW * TH = 256, TW a power of two
    int tiles_x, tiles_y;  // per image
    int debug;             // timing experiments only (TAMTR_GATECONV_DEBUG): bit 0 = epilogue does not store
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
    const uint32_t a_bytes = (uint32_t)g.TW * (g.TH + 2) * kCvKBlk * 2;
    const uint32_t b_bytes = (uint32_t)g.Cout * kCvKBlk * 2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kCvStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
        for (int s = 0; s < kCvASlots; ++s) { mbar_init(&sm.a_full[s], 1); mbar_init(&sm.a_empty[s], 1); }
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
            uint32_t kbg = 0, an = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int b = t / tiles_img, r = t - b * tiles_img;
                const int x0 = (r % g.tiles_x) * g.TW, y0 = (r / g.tiles_x) * g.TH;
                for (int kc = 0; kc < kc_per_tap; ++kc) {
                    for (int dxi = 0; dxi < 3; ++dxi, ++an) {
                        const int slot = an % kCvASlots;
                        mbar_wait(&sm.a_empty[slot], ((an / kCvASlots) & 1) ^ 1);
                        mbar_expect_tx(&sm.a_full[slot], a_bytes);
                        tma_load_4d(sm.a[slot], &tmap_x, &sm.a_full[slot], kc * kCvKBlk, x0 + dxi - 1, y0 - 1, b);
                        for (int dyi = 0; dyi < 3; ++dyi, ++kbg) {
                            const int s = kbg % kCvStages;
                            mbar_wait(&sm.empty[s], ((kbg / kCvStages) & 1) ^ 1);
                            mbar_expect_tx(&sm.full[s], b_bytes);
                            tma_load_2d(sm.b[s], &tmap_w, &sm.full[s], (dyi * 3 + dxi) * g.Cin + kc * kCvKBlk, 0);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp walks the loop (uniform control flow and registers), one elected lane issues
        {
            // instruction descriptor: D = f32, A = B = bf16, both K-major, N = Cout, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(g.Cout >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t a_lo0 = desc_lo(smem_u32(sm.a[0])), b_lo0 = desc_lo(smem_u32(sm.b[0]));
            const uint32_t row_shift = (uint32_t)g.TW * 128 >> 4;     // one image row of the patch, in descriptor units
            uint32_t kbg = 0, an = 0;
            int it = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
                mbar_wait(&sm.acc_empty, (it & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int ks = 0; ks < 3 * kc_per_tap; ++ks, ++an) {   // (channel block, dx) patches
                    const int slot = an % kCvASlots;
                    mbar_wait(&sm.a_full[slot], (an / kCvASlots) & 1);
                    uint32_t a_lo = a_lo0 + slot * (kCvABytes >> 4);
                    for (int dyi = 0; dyi < 3; ++dyi, ++kbg, a_lo += row_shift) {
                        const int s = kbg % kCvStages;
                        mbar_wait(&sm.full[s], (kbg / kCvStages) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t b_lo = b_lo0 + s * (kCvBBytes >> 4);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < kCvKBlk / 16; ++k) {
                                const uint64_t bd = desc_of(b_lo + 2 * k);
                                const uint32_t acc = (ks | dyi | k) ? 1u : 0u;
                                umma_f16(tmem_base, desc_of(a_lo + 2 * k), bd, idesc, acc);
                                umma_f16(tmem_base + kCvMaxCout, desc_of(a_lo + (128 * 128 >> 4) + 2 * k), bd, idesc, acc);
                            }
                            umma_commit(&sm.empty[s]);
                            if (dyi == 2) umma_commit(&sm.a_empty[slot]);
                        }
                        __syncwarp();
                    }
                }
                if (elect_one()) umma_commit(&sm.acc_full);
                __syncwarp();
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
            const float *gate_px = gate == nullptr ? nullptr : gate + (size_t)b * g.nh * HW + pix;
            epilogue_row(taddr, dst, gate_px, sm.scale, sm.shift, g.Cout, hc, HW, live);
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


// ===================================================================================================================
// CTA-pair version (cta_group::2).  ncu on the single-CTA kernel above: every tcgen05.mma (M=128, N=256, K=16) takes
// ~198 clk instead of 128 -- it fetches 12 KB of operands from shared memory per instruction, and the operand fetch
// runs at ~64 B/clk/SM (tensor pipe 64.7 % of active cycles, exactly that ratio).  With two SMs on one M=256 x N=256
// tile each CTA supplies its own 128 pixel rows of A and only HALF of the weight tile (N/2 rows), i.e. 8 KB per
// instruction -> the fetch matches the 128-clk MMA.  Each CTA's TMEM holds its 128 rows x 256 columns, double-buffered
// (2 x 256 columns), so the epilogue of tile i overlaps the main loop of tile i+1.
//   per CTA: warp 0 TMA producer (its own pixel rows + its half of the weights, completion signalled on the LEADER's
//   mbarriers), warp 1 TMEM allocation (+ MMA issue in the leader CTA only; tcgen05.commit multicast to both CTAs),
//   warps 2-5 epilogue of the CTA's own 128 pixels.
constexpr int kPrASlots = 3;
constexpr int kPrABytes = 32 * 1024;                 // TW*(THc+2) <= 256 rows of 128 B
constexpr int kPrBStages = 5;
constexpr int kPrBBytes = (kCvMaxCout / 2) * kCvKBlk * 2;   // 16 KB: this CTA's half of the weight tile
constexpr int kPrThreads = 192;
constexpr int kPrAccStages = 2;
constexpr int kPrOutCh = 64;                         // channels per epilogue chunk = one 128-byte swizzled row per pixel
constexpr int kPrOutBytes = 128 * kPrOutCh * 2;      // 16 KB

struct PrSmem {
    alignas(1024) uint8_t a[kPrASlots][kPrABytes];
    alignas(1024) uint8_t b[kPrBStages][kPrBBytes];
    alignas(1024) uint8_t out[2][kPrOutBytes];      // epilogue staging for the TMA store (128 pixels x 64 channels)
    alignas(16) float scale[kCvMaxCout];
    alignas(16) float shift[kCvMaxCout];
    alignas(8) uint64_t b_full[kPrBStages];
    uint64_t b_empty[kPrBStages];
    uint64_t a_full[kPrASlots];
    uint64_t a_empty[kPrASlots];
    uint64_t acc_full[kPrAccStages];
    uint64_t acc_empty[kPrAccStages];
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta pointer of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(const void *p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma2_load_4d(void *dst, const CUtensorMap *map, uint32_t bar_cluster, int c0, int c1,
                                             int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void *dst, const CUtensorMap *map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// arrive on the mbarrier at this shared-memory offset in BOTH CTAs of the pair once all prior MMAs have completed
__device__ __forceinline__ void umma2_commit_both(uint64_t *bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"((uint16_t)3)
        : "memory");
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, const void *src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void epilogue_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPrThreads, 1)
gate_conv3x3_pair_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                         const __grid_constant__ CUtensorMap tmap_y, const float *__restrict__ bn_scale,
                         const float *__restrict__ bn_shift, const float *__restrict__ gate, const CvGeom g) {
    extern __shared__ uint8_t smem_raw[];
    PrSmem &sm = *reinterpret_cast<PrSmem *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();          // 0 = leader (issues the MMAs), 1 = peer
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int tiles_img = g.tiles_x * g.tiles_y;
    const int n_tiles = g.B * tiles_img;
    const int kc_per_tap = g.Cin / kCvKBlk;
    const int THc = g.TH / 2;                         // image rows of the pair tile owned by one CTA (TW * THc <= 128)
    const uint32_t a_bytes = (uint32_t)g.TW * (THc + 2) * kCvKBlk * 2;
    const uint32_t b_bytes = (uint32_t)(g.Cout / 2) * kCvKBlk * 2;

    if (threadIdx.x == 0) {
        // full barriers live in the leader: ONE arrival (the leader's expect_tx of both CTAs' bytes); the peer's TMA only
        // contributes complete_tx (a remote arrive per stage costs the peer producer a MEMBAR: measured 3.4x slower)
        for (int s = 0; s < kPrBStages; ++s) { mbar_init(&sm.b_full[s], 1); mbar_init(&sm.b_empty[s], 1); }
        for (int s = 0; s < kPrASlots; ++s) { mbar_init(&sm.a_full[s], 1); mbar_init(&sm.a_empty[s], 1); }
        for (int s = 0; s < kPrAccStages; ++s) { mbar_init(&sm.acc_full[s], 1); mbar_init(&sm.acc_empty[s], 2 * 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)),
                     "n"(kCvTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    for (int c = threadIdx.x; c < g.Cout; c += kPrThreads) {
        sm.scale[c] = __ldg(bn_scale + c);
        sm.shift[c] = __ldg(bn_shift + c);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                               // the peer's mbarriers exist before anything signals them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = sm.tmem_base;

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own pixel rows, own half of the weight rows; completion -> leader's barriers
        if (lane == 0) {
            uint32_t kbg = 0, an = 0;
            for (int t = pair; t < n_tiles; t += n_pairs) {
                const int b = t / tiles_img, r = t - b * tiles_img;
                const int x0 = (r % g.tiles_x) * g.TW, y0 = (r / g.tiles_x) * g.TH + (int)rank * THc;
                for (int kc = 0; kc < kc_per_tap; ++kc) {
                    for (int dxi = 0; dxi < 3; ++dxi, ++an) {
                        const int slot = an % kPrASlots;
                        mbar_wait(&sm.a_empty[slot], ((an / kPrASlots) & 1) ^ 1);
                        const uint32_t abar = map_to_cta(&sm.a_full[slot], 0);
                        if (rank == 0) mbar_expect_tx(&sm.a_full[slot], 2 * a_bytes);
                        tma2_load_4d(sm.a[slot], &tmap_x, abar, kc * kCvKBlk, x0 + dxi - 1, y0 - 1, b);
                        for (int dyi = 0; dyi < 3; ++dyi, ++kbg) {
                            const int s = kbg % kPrBStages;
                            mbar_wait(&sm.b_empty[s], ((kbg / kPrBStages) & 1) ^ 1);
                            const uint32_t bbar = map_to_cta(&sm.b_full[s], 0);
                            if (rank == 0) mbar_expect_tx(&sm.b_full[s], 2 * b_bytes);
                            tma2_load_2d(sm.b[s], &tmap_w, bbar, (dyi * 3 + dxi) * g.Cin + kc * kCvKBlk,
                                         (int)rank * (g.Cout / 2));
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: leader CTA only; the whole warp walks the loop, one elected lane issues
        if (rank == 0) {
            // instruction descriptor: D = f32, A = B = bf16, both K-major, N = Cout, M = 256 (128 rows per CTA)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(g.Cout >> 3) << 17) | ((256u >> 4) << 24);
            const uint32_t a_lo0 = desc_lo(smem_u32(sm.a[0])), b_lo0 = desc_lo(smem_u32(sm.b[0]));
            const uint32_t row_shift = (uint32_t)g.TW * 128 >> 4;
            uint32_t kbg = 0, an = 0;
            int it = 0;
            for (int t = pair; t < n_tiles; t += n_pairs, ++it) {
                const int as = it % kPrAccStages;
                mbar_wait(&sm.acc_empty[as], ((it / kPrAccStages) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem_base + as * kCvMaxCout;
                for (int ks = 0; ks < 3 * kc_per_tap; ++ks, ++an) {
                    const int slot = an % kPrASlots;
                    mbar_wait(&sm.a_full[slot], (an / kPrASlots) & 1);
                    uint32_t a_lo = a_lo0 + slot * (kPrABytes >> 4);
                    for (int dyi = 0; dyi < 3; ++dyi, ++kbg, a_lo += row_shift) {
                        const int s = kbg % kPrBStages;
                        mbar_wait(&sm.b_full[s], (kbg / kPrBStages) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t b_lo = b_lo0 + s * (kPrBBytes >> 4);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < kCvKBlk / 16; ++k)
                                umma2_f16(d, desc_of(a_lo + 2 * k), desc_of(b_lo + 2 * k), idesc, (ks | dyi | k) ? 1u : 0u);
                            umma2_commit_both(&sm.b_empty[s]);
                            if (dyi == 2) umma2_commit_both(&sm.a_empty[slot]);
                        }
                        __syncwarp();
                    }
                }
                if (elect_one()) umma2_commit_both(&sm.acc_full[as]);
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue (both CTAs): warp w owns TMEM lanes 32*(w%4).. = pixels of this CTA's half of the pair tile.
        // 64 channels at a time: TMEM -> registers -> (acc*s + t)*gate -> bf16 -> the pixel's 128-byte row of a swizzled
        // staging tile -> ONE TMA store of the (64 ch, TW, THc) box; the tensor map clips the ragged image border.
        // (Direct 16-byte global stores from the 32 pixel-owning lanes touch 32 lines per instruction: measured 8 us of
        // the 90 us kernel, they contend with the tensor cores' operand reads in the shared L1/smem array.)
        const int quarter = warp & 3;
        const int p = quarter * 32 + lane;
        const int et = (int)threadIdx.x - 64;
        const int px_in = p % g.TW, py_in = p / g.TW + (int)rank * THc;   // rows p >= TW*THc of the M=128 block are idle
        const int hc = g.Cout / g.nh;
        const size_t HW = (size_t)g.H * g.W;
        const uint32_t row_off = (uint32_t)p * 128, sw = (uint32_t)(p & 7);
        uint32_t chunk = 0;
        int it = 0;
        for (int t = pair; t < n_tiles; t += n_pairs, ++it) {
            const int as = it % kPrAccStages;
            const int b = t / tiles_img, r = t - b * tiles_img;
            const int x0 = (r % g.tiles_x) * g.TW, y0 = (r / g.tiles_x) * g.TH;
            const int px = x0 + px_in, py = y0 + py_in;
            const bool live = p < g.TW * THc && px < g.W && py < g.H;
            const float *gate_px = (gate == nullptr || !live) ? nullptr : gate + (size_t)b * g.nh * HW + (size_t)py * g.W + px;
            mbar_wait(&sm.acc_full[as], (it / kPrAccStages) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * kCvMaxCout;
            for (int c0 = 0; c0 < g.Cout; c0 += kPrOutCh, ++chunk) {
                uint8_t *stage = sm.out[chunk & 1];
                if (et == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // store of chunk-2 has read it
                epilogue_bar();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int cc = c0 + 32 * h;
                    uint32_t v[32];
                    tmem_ld32(taddr + cc, v);
                    const float gt = gate_px != nullptr ? __ldg(gate_px + (size_t)(cc / hc) * HW) : 1.0f;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t o[4];
#pragma unroll
                        for (int j = 0; j < 8; j += 4) {
                            const float4 s4 = *reinterpret_cast<const float4 *>(&sm.scale[cc + q * 8 + j]);
                            const float4 t4 = *reinterpret_cast<const float4 *>(&sm.shift[cc + q * 8 + j]);
                            const float f0 = fmaf(__uint_as_float(v[q * 8 + j + 0]), s4.x, t4.x) * gt;
                            const float f1 = fmaf(__uint_as_float(v[q * 8 + j + 1]), s4.y, t4.y) * gt;
                            const float f2 = fmaf(__uint_as_float(v[q * 8 + j + 2]), s4.z, t4.z) * gt;
                            const float f3 = fmaf(__uint_as_float(v[q * 8 + j + 3]), s4.w, t4.w) * gt;
                            __nv_bfloat162 lo = __floats2bfloat162_rn(f0, f1), hi = __floats2bfloat162_rn(f2, f3);
                            o[j / 2] = *reinterpret_cast<uint32_t *>(&lo);
                            o[j / 2 + 1] = *reinterpret_cast<uint32_t *>(&hi);
                        }
                        // 16-byte piece (h*4 + q) of this pixel's row, at its 128B-swizzled position
                        const uint32_t piece = (uint32_t)(h * 4 + q) ^ sw;
                        *reinterpret_cast<uint4 *>(stage + row_off + piece * 16) = make_uint4(o[0], o[1], o[2], o[3]);
                    }
                }
                if (c0 + kPrOutCh >= g.Cout) {                        // all TMEM reads of this tile done: hand it back
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(map_to_cta(&sm.acc_empty[as], 0));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                epilogue_bar();
                if (et == 0 && !(g.debug & 1)) tma_store_4d(&tmap_y, stage, c0, x0, y0 + (int)rank * THc, b);
            }
        }
        if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                               // nobody leaves while the pair's MMAs / multicast arrives can touch it
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kCvTmemCols));
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

// bf16 fast path: 64 channels x 128 pixels per CTA, 16-byte global accesses on both sides.  Shared memory holds 32-bit
// words = (pixel 2q, pixel 2q+1) of one channel; the output side reads 8 channels of one pixel pair and splits the
// halves with byte permutes, so each thread stores two full 16-byte pieces (8 channels of one pixel each).
__global__ void __launch_bounds__(256) nchw_to_nhwc_bf16_kernel(const uint16_t *__restrict__ x, uint16_t *__restrict__ y,
                                                                int C, int HW) {
    __shared__ uint32_t tile[64][65];
    const int b = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 128;
    const uint16_t *xb = x + (size_t)b * C * HW;
    uint16_t *yb = y + (size_t)b * C * HW;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int piece = threadIdx.x + 256 * i;          // 64 channels x 16 pieces of 8 pixels
        const int c = piece >> 4, k = piece & 15;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (c0 + c < C && p0 + 8 * k < HW) v = *reinterpret_cast<const uint4 *>(xb + (size_t)(c0 + c) * HW + p0 + 8 * k);
        tile[c][4 * k + 0] = v.x;
        tile[c][4 * k + 1] = v.y;
        tile[c][4 * k + 2] = v.z;
        tile[c][4 * k + 3] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int item = threadIdx.x + 256 * i;           // 64 pixel pairs x 8 channel groups of 8
        const int cg = item & 7, pp = item >> 3;
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = tile[8 * cg + j][pp];
        uint4 lo, hi;
        lo.x = __byte_perm(w[0], w[1], 0x5410); hi.x = __byte_perm(w[0], w[1], 0x7632);
        lo.y = __byte_perm(w[2], w[3], 0x5410); hi.y = __byte_perm(w[2], w[3], 0x7632);
        lo.z = __byte_perm(w[4], w[5], 0x5410); hi.z = __byte_perm(w[4], w[5], 0x7632);
        lo.w = __byte_perm(w[6], w[7], 0x5410); hi.w = __byte_perm(w[6], w[7], 0x7632);
        const int pix = p0 + 2 * pp, c = c0 + 8 * cg;
        if (c < C) {
            if (pix < HW) *reinterpret_cast<uint4 *>(yb + (size_t)pix * C + c) = lo;
            if (pix + 1 < HW) *reinterpret_cast<uint4 *>(yb + (size_t)(pix + 1) * C + c) = hi;
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
        if (dtype == TAMTR_BF16 && HW % 8 == 0 && C % 8 == 0 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0)
            nchw_to_nhwc_bf16_kernel<<<dim3((HW + 127) / 128, (C + 63) / 64, B), 256, 0, st>>>(
                (const uint16_t *)x, (uint16_t *)y, C, HW);
        else if (dtype == TAMTR_BF16)
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

    // Two SMs per tile (cta_group::2) unless the debug override TAMTR_GATECONV_SINGLE_CTA=1 asks for the single-CTA kernel.
    static const bool single = [] { const char *e = getenv("TAMTR_GATECONV_SINGLE_CTA"); return e && e[0] == '1'; }();
    CvGeom g{B, H, W, Cin, Cout, nh, 16, 16, 0, 0, 0};
    { const char *e = getenv("TAMTR_GATECONV_DEBUG"); if (e) g.debug = atoi(e); }
    if (single) {
        // the power-of-two TW x 256/TW patch that covers the map with the fewest tiles
        long best = -1;
        for (int tw = 8; tw <= 32; tw *= 2) {
            const int th = kCvTilePix / tw;
            const long n = (long)((W + tw - 1) / tw) * ((H + th - 1) / th);
            if (best < 0 || n < best || (n == best && tw == 16)) { best = n; g.TW = tw; g.TH = th; }
        }
    } else {
        // per CTA a TW x THc patch (TW a multiple of 8 so that a row tap is a whole number of swizzle atoms,
        // TW*THc <= 128 = UMMA rows, TW*(THc+2) <= 256 patch rows); the pair stacks two of them.  Fewest tiles wins, then the
        // fuller M block, then the smaller patch.
        long best_n = -1;
        int best_fill = 0, best_rows = 0;
        for (int tw = 8; tw <= 80; tw += 8) {
            const int thc = 128 / tw, rows = tw * (thc + 2);
            if (thc < 1 || rows > 256) continue;
            const long n = (long)((W + tw - 1) / tw) * ((H + 2 * thc - 1) / (2 * thc));
            const int fill = tw * thc;
            if (best_n < 0 || n < best_n || (n == best_n && (fill > best_fill || (fill == best_fill && rows < best_rows)))) {
                best_n = n; best_fill = fill; best_rows = rows;
                g.TW = tw; g.TH = 2 * thc;
            }
        }
    }
    g.tiles_x = (W + g.TW - 1) / g.TW;
    g.tiles_y = (H + g.TH - 1) / g.TH;

    const int box_rows = single ? g.TH + 2 : g.TH / 2 + 2;     // image rows per activation box (with the two halo rows)
    const int box_cout = single ? Cout : Cout / 2;             // weight rows per CTA
    CUtensorMap tmap_x, tmap_w;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        const cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
        const cuuint32_t box[4] = {(cuuint32_t)kCvKBlk, (cuuint32_t)g.TW, (cuuint32_t)box_rows, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        const CUresult cr = encode(&tmap_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(x_nhwc), dims, strides,
                                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TAMTR_CHECK_ARG(cr == CUDA_SUCCESS, TAMTR_E_BADARG, "gate_conv3x3_tc: tensor map (x) failed (%d)", (int)cr);
    }
    {
        const cuuint64_t dims[2] = {(cuuint64_t)9 * Cin, (cuuint64_t)Cout};
        const cuuint64_t strides[1] = {(cuuint64_t)9 * Cin * 2};
        const cuuint32_t box[2] = {(cuuint32_t)kCvKBlk, (cuuint32_t)box_cout};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult cr = encode(&tmap_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(w_ohwi), dims, strides,
                                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TAMTR_CHECK_ARG(cr == CUDA_SUCCESS, TAMTR_E_BADARG, "gate_conv3x3_tc: tensor map (w) failed (%d)", (int)cr);
    }

    CUtensorMap tmap_y;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        const cuuint64_t strides[3] = {(cuuint64_t)Cout * 2, (cuuint64_t)W * Cout * 2, (cuuint64_t)H * W * Cout * 2};
        const cuuint32_t box[4] = {(cuuint32_t)kPrOutCh, (cuuint32_t)g.TW, (cuuint32_t)g.TH / 2, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        const CUresult cr = encode(&tmap_y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, y_nhwc, dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TAMTR_CHECK_ARG(cr == CUDA_SUCCESS, TAMTR_E_BADARG, "gate_conv3x3_tc: tensor map (y) failed (%d)", (int)cr);
    }

    int n_sm = 148;
    n_sm = ::tamtr::sm_count();
    const long n_tiles = (long)B * g.tiles_x * g.tiles_y;
    // persistent CTAs (single) / CTA pairs, one per SM; equalise the number of tiles each walks so the last wave is
    // not ragged
    const long workers = single ? n_sm : n_sm / 2;
    const long waves = (n_tiles + workers - 1) / workers;
    const int n_work = (int)((n_tiles + waves - 1) / waves);
    cudaStream_t st = (cudaStream_t)stream;
    static bool attr_set[64] = {false};          // cudaFuncSetAttribute is per device
    int dev_id = 0;
    TAMTR_CUDA_OK(cudaGetDevice(&dev_id));
    if (dev_id < 0 || dev_id >= 64 || !attr_set[dev_id]) {
        TAMTR_CUDA_OK(cudaFuncSetAttribute(gate_conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(sizeof(CvSmem) + 1024)));
        TAMTR_CUDA_OK(cudaFuncSetAttribute(gate_conv3x3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(sizeof(PrSmem) + 1024)));
        if (dev_id >= 0 && dev_id < 64) attr_set[dev_id] = true;
    }
    {
        KernelTimer timer(K_GATE_CONV_TC, st);
        if (single)
            gate_conv3x3_tc_kernel<<<n_work, kCvThreads, sizeof(CvSmem) + 1024, st>>>(
                tmap_x, tmap_w, bn_scale, bn_shift, gate, (__nv_bfloat16 *)y_nhwc, g);
        else
            gate_conv3x3_pair_kernel<<<2 * n_work, kPrThreads, sizeof(PrSmem) + 1024, st>>>(
                tmap_x, tmap_w, tmap_y, bn_scale, bn_shift, gate, g);
    }
    count_launch();
    TAMTR_CUDA_OK(cudaGetLastError());
    return 0;
}
