/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the multi-scale deformable attention core op.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Restates, with explicit indexing (no grid_sample), what the reference computes in
 *   /root/reference/ultralytics/nn/modules/utils.py:42-89   (multi_scale_deformable_attn_pytorch)
 * whose arithmetic lives in torch ATen's grid_sampler_2d (bilinear, padding_mode='zeros', align_corners=False):
 *   torch/include/ATen/native/GridSampler.h:27-36   (grid_sampler_unnormalize)
 *   torch/include/ATen/native/GridSampler.h:205-207 (within_bounds_2d)
 * Parity pinning: the reference ships no tests for this path ("parity unpinned by the reference"); this file is
 * pinned instead against outputs of the reference itself run in the build container (tests/golden/*.pt,
 * produced by oracle/make_goldens.py) -- see tests/test_oracle.py.
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC   (contraction off: the ONE fused multiply-add of the
 * contract is written as an explicit fmaf below).
 *
 * Layouts (all contiguous, fp32):
 *   value [B, Lv, H, Dh]      level l occupies tokens [start_l, start_l + H_l*W_l), row-major (y*W_l + x)
 *   shapes[L][2] = (H_l, W_l) (utils.py:56,60)
 *   loc   [B, Lq, H, L, P, 2] (x, y) in [0,1] image-normalised, may leave [0,1]
 *   attn  [B, Lq, H, L, P]
 *   out   [B, Lq, H*Dh]
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* utils.py:58  g = 2*loc - 1 (two separately rounded fp32 ops), then GridSampler.h:34 ((g+1)*size-1)/2 which
 * executes FMA-contracted on both the CPU vectorised kernel and the CUDA kernel:
 *   ix = fmaf(g + 1, size, -1) * 0.5                                                                        */
static inline float unnormalize(float loc, int size) {
    volatile float two_loc = 2.0f * loc;
    volatile float g = two_loc - 1.0f;
    volatile float gp = g + 1.0f;
    return fmaf(gp, (float)size, -1.0f) * 0.5f;
}

typedef struct {
    int x0, y0;        /* north-west corner */
    float w[4];        /* nw, ne, sw, se bilinear weights (GridSampler formulas: (x1-ix)*(y1-iy) ...) */
    int inb[4];        /* corner inside [0,W)x[0,H) (zeros padding) */
    float ix, iy;
} tap_t;

static inline void make_tap(float lx, float ly, int Hl, int Wl, tap_t *t) {
    float ix = unnormalize(lx, Wl), iy = unnormalize(ly, Hl);
    float fx = floorf(ix), fy = floorf(iy);
    /* NaN / huge coordinates: keep the int conversion defined; such taps are fully out of bounds */
    int x0 = (fx >= -2.0f && fx <= (float)Wl + 1.0f) ? (int)fx : -2;
    int y0 = (fy >= -2.0f && fy <= (float)Hl + 1.0f) ? (int)fy : -2;
    float x1f = fx + 1.0f, y1f = fy + 1.0f;
    t->ix = ix; t->iy = iy; t->x0 = x0; t->y0 = y0;
    t->w[0] = (x1f - ix) * (y1f - iy);
    t->w[1] = (ix - fx) * (y1f - iy);
    t->w[2] = (x1f - ix) * (iy - fy);
    t->w[3] = (ix - fx) * (iy - fy);
    int x1 = x0 + 1, y1 = y0 + 1;
    int vx0 = (x0 >= 0 && x0 < Wl), vx1 = (x1 >= 0 && x1 < Wl);
    int vy0 = (y0 >= 0 && y0 < Hl), vy1 = (y1 >= 0 && y1 < Hl);
    int finite = (ix == ix) && (iy == iy) && fx >= -2.0f && fx <= (float)Wl + 1.0f && fy >= -2.0f && fy <= (float)Hl + 1.0f;
    t->inb[0] = finite && vx0 && vy0; t->inb[1] = finite && vx1 && vy0;
    t->inb[2] = finite && vx0 && vy1; t->inb[3] = finite && vx1 && vy1;
}

static void level_starts(const int *shapes, int L, int *start) {
    int s = 0;
    for (int l = 0; l < L; ++l) { start[l] = s; s += shapes[2 * l] * shapes[2 * l + 1]; }
}

/* Integer corner indices + in-bounds flags: the "bit-exact" object of the parity contract.
 * x0,y0: [B,Lq,H,L,P] int32;  inb: [B,Lq,H,L,P,4] uint8 (nw,ne,sw,se). */
int oracle_msda_corners(const float *loc, const int *shapes, int B, int Lq, int H, int L, int P,
                        int32_t *x0, int32_t *y0, uint8_t *inb) {
    long n = (long)B * Lq * H;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i)
        for (int l = 0; l < L; ++l)
            for (int p = 0; p < P; ++p) {
                long s = (i * L + l) * P + p;
                tap_t t;
                make_tap(loc[2 * s], loc[2 * s + 1], shapes[2 * l], shapes[2 * l + 1], &t);
                x0[s] = t.x0; y0[s] = t.y0;
                for (int k = 0; k < 4; ++k) inb[4 * s + k] = (uint8_t)t.inb[k];
            }
    return 0;
}

int oracle_msda_forward(const float *value, const int *shapes, const float *loc, const float *attn,
                        int B, int Lv, int H, int Dh, int Lq, int L, int P, float *out) {
    int start[64];
    if (L > 64 || Dh > 1024) return -1;
    level_starts(shapes, L, start);
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int q = 0; q < Lq; ++q) {
            double acc[1024];
            for (int h = 0; h < H; ++h) {
                for (int c = 0; c < Dh; ++c) acc[c] = 0.0;
                long qh = ((long)b * Lq + q) * H + h;
                for (int l = 0; l < L; ++l) {
                    int Hl = shapes[2 * l], Wl = shapes[2 * l + 1];
                    for (int p = 0; p < P; ++p) {
                        long s = (qh * L + l) * P + p;
                        tap_t t;
                        make_tap(loc[2 * s], loc[2 * s + 1], Hl, Wl, &t);
                        float a = attn[s];
                        for (int k = 0; k < 4; ++k) {
                            if (!t.inb[k]) continue;
                            int x = t.x0 + (k & 1), y = t.y0 + (k >> 1);
                            const float *v = value + (((long)b * Lv + start[l] + (long)y * Wl + x) * H + h) * Dh;
                            double wk = (double)a * (double)t.w[k];
                            for (int c = 0; c < Dh; ++c) acc[c] += wk * (double)v[c];
                        }
                    }
                }
                float *o = out + qh * Dh;   /* out[b,q,h*Dh+c] */
                for (int c = 0; c < Dh; ++c) o[c] = (float)acc[c];
            }
        }
    return 0;
}

/* grad_value must be zero-initialised by the caller? No: this function zeroes it (same as the product op). */
int oracle_msda_backward(const float *grad_out, const float *value, const int *shapes, const float *loc,
                         const float *attn, int B, int Lv, int H, int Dh, int Lq, int L, int P,
                         float *grad_value, float *grad_loc, float *grad_attn) {
    int start[64];
    if (L > 64) return -1;
    level_starts(shapes, L, start);
    memset(grad_value, 0, sizeof(float) * (size_t)B * Lv * H * Dh);
    /* (b,h) pairs own disjoint slices of grad_value -> race free */
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int h = 0; h < H; ++h)
            for (int q = 0; q < Lq; ++q) {
                long qh = ((long)b * Lq + q) * H + h;
                const float *g = grad_out + qh * Dh;
                for (int l = 0; l < L; ++l) {
                    int Hl = shapes[2 * l], Wl = shapes[2 * l + 1];
                    for (int p = 0; p < P; ++p) {
                        long s = (qh * L + l) * P + p;
                        tap_t t;
                        make_tap(loc[2 * s], loc[2 * s + 1], Hl, Wl, &t);
                        float a = attn[s];
                        double dot[4] = {0, 0, 0, 0};
                        for (int k = 0; k < 4; ++k) {
                            if (!t.inb[k]) continue;
                            int x = t.x0 + (k & 1), y = t.y0 + (k >> 1);
                            long off = (((long)b * Lv + start[l] + (long)y * Wl + x) * H + h) * Dh;
                            const float *v = value + off;
                            float *gv = grad_value + off;
                            float awk = a * t.w[k];
                            double d = 0.0;
                            for (int c = 0; c < Dh; ++c) { d += (double)g[c] * (double)v[c]; gv[c] += awk * g[c]; }
                            dot[k] = d;
                        }
                        float fx = floorf(t.ix), fy = floorf(t.iy);
                        double tx = (double)(t.ix - fx), ty = (double)(t.iy - fy);
                        double ux = (double)((fx + 1.0f) - t.ix), uy = (double)((fy + 1.0f) - t.iy);
                        grad_attn[s] = (float)((double)t.w[0] * dot[0] + (double)t.w[1] * dot[1] +
                                               (double)t.w[2] * dot[2] + (double)t.w[3] * dot[3]);
                        double gix = (double)a * (-uy * dot[0] + uy * dot[1] - ty * dot[2] + ty * dot[3]);
                        double giy = (double)a * (-ux * dot[0] - tx * dot[1] + ux * dot[2] + tx * dot[3]);
                        /* d ix / d loc_x = W_l (GridSampler.h:51 gives W/2, times d(2*loc-1)/d loc = 2) */
                        grad_loc[2 * s] = (float)(gix * (double)Wl);
                        grad_loc[2 * s + 1] = (float)(giy * (double)Hl);
                    }
                }
            }
    return 0;
}
