/*
 * gavatar.h — C ABI of libgavatar_sm100.so, the B200-native (sm_100a) implementation of GaussianAvatar's per-frame
 * render hot path.  Flat extern "C" surface: plain device/host pointers, sizes, a CUDA stream handle passed as
 * void*.  No torch types.  Every function returns 0 (GA_OK) or a negative GaStatus; ga_last_error() gives the
 * message of the last failure on the calling thread.  The library never allocates persistent device memory: the
 * caller owns every buffer (sizes from the *_bytes queries) and keeps them alive until the matching backward ran.
 *
 * All work is enqueued on `stream`; calls are asynchronous unless stated otherwise.  Not re-entrant per workspace;
 * independent workspaces / streams may run concurrently.  One process per GPU.
 *
 * What each entry point replaces in the reference (paths relative to /root/reference; [UPSTREAM] = the un-vendored
 * diff-gaussian-rasterization extension the reference imports at gaussian_renderer/__init__.py:6):
 *
 *   ga_raster_*          GaussianRasterizer(raster_settings)(means3D=..., ...)      gaussian_renderer/__init__.py:36-48
 *                        ([UPSTREAM] _C.rasterize_gaussians / _C.rasterize_gaussians_backward)
 *   ga_lbs_*             pred_res*0.02 / mask-select / +query_points / two einsums / scale ramp / repeat / reg. losses
 *                                                                                    model/avatar_model.py:308-330
 *   ga_smpl_*            SMPL.forward(...).A @ inv_mats                              model/avatar_model.py:291-296,
 *                                                                                    submodules/smplx/lbs.py:152-252,349-405
 *   ga_decoder_* ga_conv* POP_no_unet.forward (GeomConvLayers, grid_sample, ShapeDecoder)
 *                                                                                    model/network.py:39-83, model/modules.py:114-137,508-582
 *   ga_loss_*            l1_loss_w + ssim                                            utils/loss_utils.py:7-53
 *   ga_adam_*            torch.optim.Adam step on net + geo_feature                  model/avatar_model.py:148-155,264-267
 */
#ifndef GAVATAR_H_
#define GAVATAR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum GaStatus {
    GA_OK = 0,
    GA_ERR_INVALID = -1,   /* bad shape / null pointer / unsupported option */
    GA_ERR_CAPACITY = -2,  /* a caller-provided buffer is too small */
    GA_ERR_CUDA = -3       /* a CUDA runtime call failed; see ga_last_error() */
} GaStatus;

int ga_version(void);
const char *ga_last_error(void);
/* Number of kernels this library has launched in the calling process since load (bench.py's gpu_launches). */
long long ga_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Rasterizer.  Layout contracts: contiguous fp32; means3D [P,3]; colors [P,3] (colors_precomp);
 * opacities [P]; scales [P,3]; rotations [P,4] (w,x,y,z, un-normalised); bg [3]; viewmatrix / projmatrix = the
 * reference's transposed 4x4s (scene/dataset_mono.py:248-250), i.e. flat[col*4+row]; image planar [3,H,W];
 * radii int32 [P].
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct GaRasterSettings {
    int32_t P, H, W;
    float tanfovx, tanfovy, scale_modifier;
} GaRasterSettings;

size_t ga_raster_geom_bytes(int32_t P);                       /* per-Gaussian state (upstream's geomBuffer)   */
size_t ga_raster_img_bytes(int32_t H, int32_t W);             /* final_T, n_contrib, tile ranges (imgBuffer)  */
size_t ga_raster_binning_bytes(int64_t num_rendered, int32_t H, int32_t W); /* keys/values + sort temp (binningBuffer) */
size_t ga_raster_bwd_scratch_bytes(int32_t P);                /* dL/dmean2D, dL/dconic, dL/dopacity scratch   */

/* Stage 1: K1 preprocess + K2 scan.  Synchronises `stream` once to return the number of (Gaussian, tile) instances
 * (upstream does the same D2H read to size its binning buffer). */
int ga_raster_forward_preprocess(const GaRasterSettings *s, const float *means3D, const float *scales,
                                 const float *rotations, const float *opacities, const float *viewmatrix,
                                 const float *projmatrix, void *geom, int32_t *radii, int64_t *num_rendered_host,
                                 void *stream);
/* Stage 2: K3 duplicate-with-keys, K4 radix sort, K5 tile ranges, K6 alpha compositing. */
int ga_raster_forward_render(const GaRasterSettings *s, const float *colors, const float *bg, void *geom, void *binning,
                             size_t binning_bytes, int64_t num_rendered, void *img, float *out_color, void *stream);
/* Backward (K7 + fused K8/K9).  d_opacities, d_rotations, d_means2D ([P,3], xy filled) may be NULL.
 * d_* outputs are overwritten (not accumulated). */
int ga_raster_backward(const GaRasterSettings *s, const float *means3D, const float *colors, const float *scales,
                       const float *rotations, const float *bg, const float *viewmatrix, const float *projmatrix,
                       const int32_t *radii, const void *geom, const void *binning, const void *img,
                       int64_t num_rendered, const float *dL_dout, void *scratch, float *d_means3D, float *d_colors,
                       float *d_scales, float *d_rotations, float *d_opacities, float *d_means2D, void *stream);

/* Debug / parity accessors into the opaque buffers (device pointers; valid after the forward that filled them). */
typedef struct GaRasterViews {
    const float *depth;          /* [P]   */
    const float *xy;             /* [P,2] */
    const float *conic_opacity;  /* [P,4] */
    const float *cov3d;          /* [P,6] */
    const uint32_t *tiles_touched; /* [P] */
    const uint32_t *offsets;     /* [P] inclusive scan */
    const uint16_t *rect;        /* [P,4] min.x min.y max.x max.y */
    const uint64_t *keys_unsorted, *keys_sorted;   /* [R] */
    const uint32_t *vals_unsorted, *vals_sorted;   /* [R] */
    const uint32_t *ranges;      /* [T,2] */
    const float *final_T;        /* [H*W] */
    const uint32_t *n_contrib;   /* [H*W] */
} GaRasterViews;
int ga_raster_views(const GaRasterSettings *s, const void *geom, const void *binning, const void *img,
                    int64_t num_rendered, GaRasterViews *out);

#ifdef __cplusplus
}
#endif
#endif /* GAVATAR_H_ */
