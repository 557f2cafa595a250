/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see raster_oracle_impl.inc for the full header).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * PARITY UNPINNED: upstream diff-gaussian-rasterization is neither vendored in /root/reference nor pinned.
 *
 * Exports two variants of the same restatement:
 *   oracle32_*  REAL=float  — bit-exact target for integer outputs, image reference
 *   oracle64_*  REAL=double — gradient reference
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -fopenmp)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

/* ---- float variant ---- */
#define REAL float
#define FN(x) CAT(f32_, x)
#define FMA(a, b, c) fmaf((a), (b), (c))
#define FMIN(a, b) fminf((a), (b))
#define FMAX(a, b) fmaxf((a), (b))
#define SQRT(a) sqrtf(a)
#define CEIL(a) ceilf(a)
#define EXP(a) expf(a)
#include "raster_oracle_impl.inc"
#undef REAL
#undef FN
#undef FMA
#undef FMIN
#undef FMAX
#undef SQRT
#undef CEIL
#undef EXP

/* ---- double variant ---- */
#define REAL double
#define FN(x) CAT(f64_, x)
#define FMA(a, b, c) fma((a), (b), (c))
#define FMIN(a, b) fmin((a), (b))
#define FMAX(a, b) fmax((a), (b))
#define SQRT(a) sqrt(a)
#define CEIL(a) ceil(a)
#define EXP(a) exp(a)
#include "raster_oracle_impl.inc"
#undef REAL
#undef FN

enum { Q_DEPTH = 0, Q_RADII, Q_XY, Q_CONIC_O, Q_COV3D, Q_TILES, Q_RECT, Q_OFFSETS, Q_KEYS_UNSORTED, Q_VALS_UNSORTED,
       Q_KEYS, Q_VALS, Q_RANGES, Q_OUT, Q_FINAL_T, Q_N_CONTRIB };

#define EXPORTS(PFX, V, REALT)                                                                                         \
    void *PFX##_forward(int P, int H, int W, const float *means3D, const float *colors, const float *opac,             \
                        const float *scales, const float *rots, const float *bg, const float *view, const float *proj, \
                        float tanfovx, float tanfovy, float scale_modifier)                                            \
    {                                                                                                                  \
        return (void *)V##_forward(P, H, W, means3D, colors, opac, scales, rots, bg, view, proj, tanfovx, tanfovy,     \
                                   scale_modifier);                                                                    \
    }                                                                                                                  \
    long long PFX##_num_rendered(void *h) { return (long long)((V##_Ctx *)h)->R; }                                     \
    void PFX##_free(void *h) { V##_free_ctx((V##_Ctx *)h); }                                                           \
    static void PFX##_cpy_real(double *dst, const REALT *src, size_t n)                                                \
    {                                                                                                                  \
        for (size_t i = 0; i < n; ++i) dst[i] = (double)src[i];                                                        \
    }                                                                                                                  \
    int PFX##_get(void *h, int which, void *dst)                                                                       \
    {                                                                                                                  \
        V##_Ctx *c = (V##_Ctx *)h;                                                                                     \
        const size_t P = (size_t)c->P, HW = (size_t)c->H * c->W, R = (size_t)c->R, T = (size_t)c->T;                   \
        switch (which) {                                                                                               \
        case Q_DEPTH: PFX##_cpy_real((double *)dst, c->depth, P); break;                                               \
        case Q_RADII: memcpy(dst, c->radii, P * 4); break;                                                             \
        case Q_XY: PFX##_cpy_real((double *)dst, c->xy, 2 * P); break;                                                 \
        case Q_CONIC_O: PFX##_cpy_real((double *)dst, c->conic_o, 4 * P); break;                                       \
        case Q_COV3D: PFX##_cpy_real((double *)dst, c->cov3d, 6 * P); break;                                           \
        case Q_TILES: memcpy(dst, c->tiles, P * 4); break;                                                             \
        case Q_RECT: memcpy(dst, c->rect, 4 * P * 4); break;                                                           \
        case Q_OFFSETS: memcpy(dst, c->offsets, P * 4); break;                                                         \
        case Q_KEYS_UNSORTED: memcpy(dst, c->keys_unsorted, R * 8); break;                                             \
        case Q_VALS_UNSORTED: memcpy(dst, c->vals_unsorted, R * 4); break;                                             \
        case Q_KEYS: memcpy(dst, c->keys, R * 8); break;                                                               \
        case Q_VALS: memcpy(dst, c->vals, R * 4); break;                                                               \
        case Q_RANGES: memcpy(dst, c->ranges, 2 * T * 4); break;                                                       \
        case Q_OUT: PFX##_cpy_real((double *)dst, c->out, 3 * HW); break;                                              \
        case Q_FINAL_T: PFX##_cpy_real((double *)dst, c->final_T, HW); break;                                          \
        case Q_N_CONTRIB: memcpy(dst, c->n_contrib, HW * 4); break;                                                    \
        default: return -1;                                                                                            \
        }                                                                                                              \
        return 0;                                                                                                      \
    }                                                                                                                  \
    void PFX##_backward(void *h, const float *dL_dout, double *d_mean2D, double *d_conic, double *d_opacity,           \
                        double *d_colors, double *d_means3D, double *d_scales, double *d_rots, double *d_cov3d)        \
    {                                                                                                                  \
        V##_backward((V##_Ctx *)h, dL_dout, d_mean2D, d_conic, d_opacity, d_colors, d_means3D, d_scales, d_rots,       \
                     d_cov3d);                                                                                         \
    }

EXPORTS(oracle32, f32, float)
EXPORTS(oracle64, f64, double)

int oracle_version(void) { return 1; }
