// Shared host/device helpers for the tamtr_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/tamtr_b200.h"

namespace tamtr {

void set_error(const char *fmt, ...);
void count_launch(unsigned n = 1);
int sm_count();   // SMs of the CURRENT device (148 on B200)

enum KernelId { K_MSDA_FWD = 0, K_MSDA_BWD, K_LOCW_FWD, K_LOCW_BWD, K_CTR_FWD, K_CTR_BWD, K_GATE_FWD, K_GATE_BWD,
                K_GATE_TC_FWD, K_GATE_CONV_TC, K_NHWC, K_LN_FWD, K_LN_BWD, K_SSCAN_FWD, K_SSCAN_BWD, K_TOK_PROJECT, K_TOK_REDUCE, K_COUNT };

// RAII pair of CUDA events around one kernel launch (no-op unless tamtr_profile_enable(1)).
struct KernelTimer {
    KernelTimer(int id, cudaStream_t st);
    ~KernelTimer();
    int id_;
    cudaStream_t st_;
    bool live_;
    cudaEvent_t a_, b_;
};

#define TAMTR_CHECK_ARG(cond, code, ...)           \
    do {                                            \
        if (!(cond)) {                              \
            ::tamtr::set_error(__VA_ARGS__);        \
            return (code);                          \
        }                                           \
    } while (0)

#define TAMTR_CUDA_OK(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            ::tamtr::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));         \
            return (int)_e;                                                              \
        }                                                                                \
    } while (0)

constexpr int kMaxLevels = 8;
constexpr int kMaxSamples = 32;  // L*P

// Pyramid geometry, passed by value (lives in the kernel-parameter constant bank).
struct Levels {
    int n;                  // L
    int P;                  // points per level (0 when the levels have different point counts)
    int S;                  // samples per (query, head) = sum of the points of all levels
    int h[kMaxLevels];      // H_l
    int w[kMaxLevels];      // W_l
    int start[kMaxLevels];  // first token of level l
    unsigned char level_of[kMaxSamples];  // sample s = l*P + p -> l (saves an integer division per tap)
};

// points_host == nullptr: P points on every level (sample s = l*P + p); otherwise points_host[l] points on level l,
// samples ordered level by level (the torch.split(..., [2, 4, 6], dim=-2) of utils.py:108 / [6, 4, 2] of utils.py:159).
inline int fill_levels(Levels &lv, int L, int P, const int32_t *shapes_host, int expect_Lv,
                       const int32_t *points_host = nullptr) {
    if (L < 1 || L > kMaxLevels) return TAMTR_E_UNSUPPORTED;
    int S = 0;
    for (int l = 0; l < L; ++l) {
        const int pl = points_host ? points_host[l] : P;
        if (pl < 1) return TAMTR_E_BADARG;
        S += pl;
        if (S > kMaxSamples) return TAMTR_E_UNSUPPORTED;
    }
    lv.n = L;
    lv.P = points_host ? 0 : P;
    lv.S = S;
    long s = 0;
    for (int l = 0; l < kMaxLevels; ++l) {
        if (l < L) {
            lv.h[l] = shapes_host[2 * l];
            lv.w[l] = shapes_host[2 * l + 1];
            if (lv.h[l] < 1 || lv.w[l] < 1) return TAMTR_E_BADARG;
            lv.start[l] = (int)s;
            s += (long)lv.h[l] * lv.w[l];
        } else {
            lv.h[l] = lv.w[l] = 1;
            lv.start[l] = 0;
        }
    }
    if (expect_Lv >= 0 && s != expect_Lv) return TAMTR_E_BADARG;
    for (int i = 0; i < kMaxSamples; ++i) lv.level_of[i] = 0;
    for (int l = 0, i = 0; l < L; ++l)
        for (int p = 0; p < (points_host ? points_host[l] : P); ++p) lv.level_of[i++] = (unsigned char)l;
    return 0;
}

}  // namespace tamtr
