"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front-end of ``oracle/build/liboracle_raster.so`` (C restatement of the rasterizer GaussianAvatar calls at
/root/reference gaussian_renderer/__init__.py:21-48; algorithm per SURVEY.md §8 a-8 / a-9).  PARITY UNPINNED —
upstream diff-gaussian-rasterization is not vendored and the reference has no golden vectors for this boundary.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "build", "liboracle_raster.so")

_Q = dict(depth=0, radii=1, xy=2, conic_o=3, cov3d=4, tiles=5, rect=6, offsets=7, keys_unsorted=8, vals_unsorted=9,
          keys=10, vals=11, ranges=12, out=13, final_T=14, n_contrib=15)


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, OpenMP).  Building the checker is not using it."""
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("raster_oracle.c", "raster_oracle_impl.inc", "Makefile"))
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < src_m:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        fp = ctypes.POINTER(ctypes.c_float)
        dp = ctypes.POINTER(ctypes.c_double)
        for pfx in ("oracle32", "oracle64"):
            f = getattr(L, pfx + "_forward")
            f.restype = ctypes.c_void_p
            f.argtypes = [ctypes.c_int] * 3 + [fp] * 8 + [ctypes.c_float] * 3
            g = getattr(L, pfx + "_get")
            g.restype = ctypes.c_int
            g.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
            n = getattr(L, pfx + "_num_rendered")
            n.restype = ctypes.c_longlong
            n.argtypes = [ctypes.c_void_p]
            fr = getattr(L, pfx + "_free")
            fr.restype = None
            fr.argtypes = [ctypes.c_void_p]
            b = getattr(L, pfx + "_backward")
            b.restype = None
            b.argtypes = [ctypes.c_void_p, fp] + [dp] * 8
        _lib = L
    return _lib


def _f32(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    if shape is not None:
        a = a.reshape(shape)
    return a


def _ptr(a, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty))


@dataclass
class OracleForward:
    """Result of one oracle forward; keeps the C context alive for backward()."""
    P: int
    H: int
    W: int
    precision: str
    _h: int = field(repr=False, default=0)
    _keep: tuple = field(repr=False, default=())

    @property
    def _pfx(self):
        return "oracle32" if self.precision == "f32" else "oracle64"

    @property
    def grid(self):
        return (self.W + 15) // 16, (self.H + 15) // 16

    @property
    def num_rendered(self) -> int:
        return int(getattr(lib(), self._pfx + "_num_rendered")(self._h))

    def get(self, name: str) -> np.ndarray:
        P, H, W, R = self.P, self.H, self.W, self.num_rendered
        T = self.grid[0] * self.grid[1]
        spec = dict(depth=((P,), np.float64), radii=((P,), np.int32), xy=((P, 2), np.float64), conic_o=((P, 4), np.float64),
                    cov3d=((P, 6), np.float64), tiles=((P,), np.int32), rect=((P, 4), np.int32), offsets=((P,), np.uint32),
                    keys_unsorted=((R,), np.uint64), vals_unsorted=((R,), np.uint32), keys=((R,), np.uint64),
                    vals=((R,), np.uint32), ranges=((T, 2), np.uint32), out=((3, H, W), np.float64),
                    final_T=((H, W), np.float64), n_contrib=((H, W), np.uint32))[name]
        arr = np.zeros(spec[0], dtype=spec[1])
        if arr.size:
            rc = getattr(lib(), self._pfx + "_get")(self._h, _Q[name], arr.ctypes.data_as(ctypes.c_void_p))
            assert rc == 0
        return arr

    @property
    def image(self) -> np.ndarray:
        return self.get("out")

    def backward(self, dL_dout) -> dict:
        P = self.P
        g = _f32(dL_dout, (3, self.H, self.W))
        outs = dict(d_mean2D=np.zeros((P, 2)), d_conic=np.zeros((P, 3)), d_opacity=np.zeros((P,)), d_colors=np.zeros((P, 3)),
                    d_means3D=np.zeros((P, 3)), d_scales=np.zeros((P, 3)), d_rots=np.zeros((P, 4)), d_cov3d=np.zeros((P, 6)))
        order = ["d_mean2D", "d_conic", "d_opacity", "d_colors", "d_means3D", "d_scales", "d_rots", "d_cov3d"]
        # numpy zero-size arrays still give a valid (dummy) pointer
        getattr(lib(), self._pfx + "_backward")(self._h, _ptr(g, ctypes.c_float), *[_ptr(outs[k], ctypes.c_double) for k in order])
        return outs

    def close(self):
        if self._h:
            getattr(lib(), self._pfx + "_free")(self._h)
            self._h = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def forward(means3D, colors, opacities, scales, rotations, bg, viewmatrix, projmatrix, tanfovx, tanfovy, H, W,
            scale_modifier=1.0, precision="f32") -> OracleForward:
    """Run K1..K6 on the CPU.  viewmatrix / projmatrix are the reference's transposed 4x4s (scene/dataset_mono.py:248-250),
    i.e. flat index = col*4+row of the column-vector matrix."""
    means3D = _f32(means3D).reshape(-1, 3)
    P = means3D.shape[0]
    colors = _f32(colors, (P, 3)); opacities = _f32(opacities, (P,)); scales = _f32(scales, (P, 3)); rotations = _f32(rotations, (P, 4))
    bg = _f32(bg, (3,)); view = _f32(viewmatrix, (16,)); proj = _f32(projmatrix, (16,))
    pfx = "oracle32" if precision == "f32" else "oracle64"
    fp = ctypes.c_float
    h = getattr(lib(), pfx + "_forward")(P, int(H), int(W), _ptr(means3D, fp), _ptr(colors, fp), _ptr(opacities, fp),
                                          _ptr(scales, fp), _ptr(rotations, fp), _ptr(bg, fp), _ptr(view, fp), _ptr(proj, fp),
                                          float(tanfovx), float(tanfovy), float(scale_modifier))
    return OracleForward(P=P, H=int(H), W=int(W), precision=precision, _h=h)
