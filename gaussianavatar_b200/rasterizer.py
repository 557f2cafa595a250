"""`diff_gaussian_rasterization`-compatible Python surface over the sm_100a rasterizer.

Mirrors the API generation the reference binds at /root/reference gaussian_renderer/__init__.py:6,21-48:
``GaussianRasterizationSettings`` (12-field NamedTuple), ``GaussianRasterizer(raster_settings)(means3D=, means2D=, shs=,
colors_precomp=, opacities=, scales=, rotations=, cov3D_precomp=) -> (color[3,H,W], radii[P])``, gradients returned in
input order.  Validation behaviour follows [UPSTREAM]: exactly one of shs / colors_precomp and exactly
one of (scales+rotations) / cov3D_precomp, else ``Exception``.

Only the path GaussianAvatar exercises is implemented on the device: precomputed colours, scales + rotations.
SH evaluation and cov3D_precomp raise NotImplementedError (never used by the reference: shs=None at
model/avatar_model.py:350, cov3D_precomp=None at gaussian_renderer/__init__.py:37).
"""
from __future__ import annotations

import ctypes
from typing import NamedTuple

import torch
import torch.nn as nn

from . import _lib
from ._lib import GaRasterSettings, ptr


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def _c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class RasterContext:
    """Buffers one forward leaves behind for its backward (upstream's geomBuffer / binningBuffer / imgBuffer)."""
    __slots__ = ("settings", "geom", "binning", "img", "num_rendered", "radii")

    def views(self) -> dict:
        """Typed tensor copies of the internal per-stage state (parity tests)."""
        L = _lib.lib()
        v = _lib.GaRasterViews()
        _lib.check(L.ga_raster_views(ctypes.byref(self.settings), ptr(self.geom), ptr(self.binning), ptr(self.img),
                                     self.num_rendered, ctypes.byref(v)), "ga_raster_views")
        s = self.settings
        P, H, W, R = s.P, s.H, s.W, self.num_rendered
        T = ((W + 15) // 16) * ((H + 15) // 16)

        def grab(buf, addr, shape, dtype):
            n = 1
            for d in shape:
                n *= d
            if addr is None or n == 0:
                return torch.zeros(shape, dtype=dtype)
            off = addr - buf.data_ptr()
            nbytes = n * torch.empty((), dtype=dtype).element_size()
            return buf[off:off + nbytes].view(dtype).reshape(shape).clone().cpu()

        out = dict(depth=grab(self.geom, v.depth, (P,), torch.float32), xy=grab(self.geom, v.xy, (P, 2), torch.float32),
                   conic_opacity=grab(self.geom, v.conic_opacity, (P, 4), torch.float32),
                   cov3d=grab(self.geom, v.cov3d, (P, 6), torch.float32),
                   tiles_touched=grab(self.geom, v.tiles_touched, (P,), torch.int32),
                   offsets=grab(self.geom, v.offsets, (P,), torch.int32),
                   rect=grab(self.geom, v.rect, (P, 4), torch.int16),
                   ranges=grab(self.img, v.ranges, (T, 2), torch.int32),
                   final_T=grab(self.img, v.final_T, (H, W), torch.float32),
                   n_contrib=grab(self.img, v.n_contrib, (H, W), torch.int32),
                   radii=self.radii.clone().cpu(), num_rendered=R)
        if R > 0:
            out.update(keys_unsorted=grab(self.binning, v.keys_unsorted, (R,), torch.int64),
                       keys_sorted=grab(self.binning, v.keys_sorted, (R,), torch.int64),
                       vals_unsorted=grab(self.binning, v.vals_unsorted, (R,), torch.int32),
                       vals_sorted=grab(self.binning, v.vals_sorted, (R,), torch.int32))
        return out


def rasterize_forward(means3D, colors, opacities, scales, rotations, rs: GaussianRasterizationSettings):
    """K1..K6 on the current CUDA stream.  Returns (color [3,H,W], radii [P] int32, RasterContext)."""
    if not means3D.is_cuda:
        raise RuntimeError("gaussianavatar_b200 rasterizer needs CUDA tensors (no CPU fallback)")
    L = _lib.lib()
    dev = means3D.device
    P, H, W = int(means3D.shape[0]), int(rs.image_height), int(rs.image_width)
    s = GaRasterSettings(P, H, W, float(rs.tanfovx), float(rs.tanfovy), float(rs.scale_modifier))
    means3D, colors, scales, rotations = _c(means3D), _c(colors), _c(scales), _c(rotations)
    opacities = _c(opacities).reshape(-1)
    bg, view, proj = _c(rs.bg).to(dev), _c(rs.viewmatrix).to(dev), _c(rs.projmatrix).to(dev)

    ctx = RasterContext()
    ctx.settings = s
    ctx.geom = torch.empty(L.ga_raster_geom_bytes(P), dtype=torch.uint8, device=dev)
    ctx.img = torch.empty(L.ga_raster_img_bytes(H, W), dtype=torch.uint8, device=dev)
    ctx.radii = torch.empty(P, dtype=torch.int32, device=dev)
    n = ctypes.c_int64(0)
    st = _stream()
    _lib.check(L.ga_raster_forward_preprocess(ctypes.byref(s), ptr(means3D), ptr(scales), ptr(rotations), ptr(opacities),
                                              ptr(view), ptr(proj), ptr(ctx.geom), ptr(ctx.radii), ctypes.byref(n), st),
               "ga_raster_forward_preprocess")
    ctx.num_rendered = int(n.value)
    nb = L.ga_raster_binning_bytes(ctx.num_rendered, H, W)
    ctx.binning = torch.empty(nb, dtype=torch.uint8, device=dev)
    color = torch.empty(3, H, W, dtype=torch.float32, device=dev)
    _lib.check(L.ga_raster_forward_render(ctypes.byref(s), ptr(colors), ptr(bg), ptr(ctx.geom), ptr(ctx.binning), nb,
                                          ctx.num_rendered, ptr(ctx.img), ptr(color), st), "ga_raster_forward_render")
    return color, ctx.radii, ctx


def rasterize_backward(ctx: RasterContext, means3D, colors, scales, rotations, rs, grad_color, want_opacity=True,
                       want_rotations=True, want_means2D=True):
    L = _lib.lib()
    dev = means3D.device
    s = ctx.settings
    P = s.P
    grad_color = _c(grad_color)
    bg, view, proj = _c(rs.bg).to(dev), _c(rs.viewmatrix).to(dev), _c(rs.projmatrix).to(dev)
    scratch = torch.empty(L.ga_raster_bwd_scratch_bytes(P), dtype=torch.uint8, device=dev)
    d_means3D = torch.empty(P, 3, dtype=torch.float32, device=dev)
    d_colors = torch.empty(P, 3, dtype=torch.float32, device=dev)
    d_scales = torch.empty(P, 3, dtype=torch.float32, device=dev)
    d_rot = torch.empty(P, 4, dtype=torch.float32, device=dev) if want_rotations else None
    d_opac = torch.empty(P, 1, dtype=torch.float32, device=dev) if want_opacity else None
    d_m2d = torch.empty(P, 3, dtype=torch.float32, device=dev) if want_means2D else None
    _lib.check(L.ga_raster_backward(ctypes.byref(s), ptr(means3D), ptr(colors), ptr(scales), ptr(rotations), ptr(bg),
                                    ptr(view), ptr(proj), ptr(ctx.radii), ptr(ctx.geom), ptr(ctx.binning), ptr(ctx.img),
                                    ctx.num_rendered, ptr(grad_color), ptr(scratch), ptr(d_means3D), ptr(d_colors),
                                    ptr(d_scales), ptr(d_rot), ptr(d_opac), ptr(d_m2d), _stream()), "ga_raster_backward")
    return d_means3D, d_m2d, d_colors, d_opac, d_scales, d_rot


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings):
        if sh.numel() != 0:
            raise NotImplementedError("SH colour evaluation is not on GaussianAvatar's path (shs=None, "
                                      "model/avatar_model.py:350); pass colors_precomp")
        if cov3Ds_precomp.numel() != 0:
            raise NotImplementedError("cov3D_precomp is not on GaussianAvatar's path (gaussian_renderer/__init__.py:37)")
        means3D_c, colors_c, scales_c, rot_c = _c(means3D), _c(colors_precomp), _c(scales), _c(rotations)
        color, radii, rctx = rasterize_forward(means3D_c, colors_c, opacities, scales_c, rot_c, raster_settings)
        ctx.rctx = rctx
        ctx.raster_settings = raster_settings
        ctx.save_for_backward(means3D_c, colors_c, scales_c, rot_c)
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, grad_color, _grad_radii):
        means3D, colors, scales, rotations = ctx.saved_tensors
        need = ctx.needs_input_grad
        d_means3D, d_m2d, d_colors, d_opac, d_scales, d_rot = rasterize_backward(
            ctx.rctx, means3D, colors, scales, rotations, ctx.raster_settings, grad_color,
            want_opacity=need[4], want_rotations=need[6], want_means2D=need[1])
        return (d_means3D, d_m2d, None, d_colors, d_opac, d_scales, d_rot, None, None)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                                     raster_settings)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions: torch.Tensor) -> torch.Tensor:
        """[UPSTREAM] frustum test (view-space z > 0.2); unused by GaussianAvatar, kept for API compatibility."""
        with torch.no_grad():
            V = self.raster_settings.viewmatrix.to(positions.device).float()   # transposed: p_view = [p,1] @ V
            z = positions.float() @ V[:3, 2] + V[3, 2]
            return z > 0.2

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        empty = torch.Tensor([])
        shs = empty if shs is None else shs
        colors_precomp = empty if colors_precomp is None else colors_precomp
        scales = empty if scales is None else scales
        rotations = empty if rotations is None else rotations
        cov3D_precomp = empty if cov3D_precomp is None else cov3D_precomp
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                   self.raster_settings)
