"""CPU: pin the C oracle against (i) an independently written fp64 torch-autograd twin and (ii) analytic known answers.

The reference ships no golden vectors for the rasterizer boundary (PARITY UNPINNED), so these
self-consistency checks are what stands behind the oracle.
"""
import math

import numpy as np
import pytest
import torch

from oracle import raster_oracle as ro
from oracle.raster_twin import twin_forward
from scenes import oracle_args, random_scene


def _twin(sc, requires_grad=False):
    a = {k: sc[k].clone().double().requires_grad_(requires_grad) for k in ("means3D", "colors", "opacities", "scales", "rotations")}
    out, aux = twin_forward(a["means3D"], a["colors"], a["opacities"], a["scales"], a["rotations"], sc["bg"],
                            sc["cam"].world_view_transform, sc["cam"].full_proj_transform, sc["tanfovx"], sc["tanfovy"],
                            sc["H"], sc["W"], return_aux=True)
    return out, aux, a


@pytest.mark.parametrize("case", [
    dict(P=250, H=64, W=80, seed=1, scale_mean=0.03),
    dict(P=200, H=50, W=70, seed=2, scale_mean=0.02, aniso=False, opacity_one=True),   # alpha cap active (deviation 1)
    dict(P=200, H=64, W=64, seed=3, scale_mean=0.06, spread=2.5),                      # guard-band clamp (deviation 2)
])
def test_c_oracle_matches_autograd_twin(case):
    sc = random_scene(**case)
    o64 = ro.forward(**oracle_args(sc), precision="f64")
    out, aux, a = _twin(sc, requires_grad=True)
    assert np.abs(out.detach().numpy() - o64.image).max() < 1e-6
    np.testing.assert_array_equal(aux["radii"].numpy(), o64.get("radii"))
    np.testing.assert_array_equal(aux["n_contrib"].numpy(), o64.get("n_contrib"))
    gw = torch.randn(3, sc["H"], sc["W"], generator=torch.Generator().manual_seed(5)).double()
    (out * gw).sum().backward()
    gb = o64.backward(gw.numpy())
    for k, tk in (("d_means3D", "means3D"), ("d_colors", "colors"), ("d_scales", "scales"), ("d_rots", "rotations"),
                  ("d_opacity", "opacities")):
        ref = a[tk].grad.numpy().reshape(gb[k].shape)
        assert np.abs(ref - gb[k]).max() <= 1e-6 * max(1.0, np.abs(ref).max()), k
    if case.get("spread", 0) > 2:
        # the scene must actually exercise the clamp
        t = sc["means3D"].double() @ sc["cam"].world_view_transform.double()[:3, :3] + sc["cam"].world_view_transform.double()[3, :3]
        vis = t[:, 2] > 0.2
        assert ((t[vis, 0] / t[vis, 2]).abs() > 1.3 * sc["tanfovx"]).any()


def test_f32_oracle_close_to_f64():
    sc = random_scene(P=400, H=96, W=96, seed=9)
    o32 = ro.forward(**oracle_args(sc), precision="f32")
    o64 = ro.forward(**oracle_args(sc), precision="f64")
    assert o32.num_rendered == o64.num_rendered
    assert np.abs(o32.image - o64.image).max() < 1e-5
    gw = np.random.default_rng(0).standard_normal((3, 96, 96)).astype(np.float32)
    g32, g64 = o32.backward(gw), o64.backward(gw)
    for k in ("d_means3D", "d_colors", "d_scales", "d_rots"):
        assert np.linalg.norm(g32[k] - g64[k]) / np.linalg.norm(g64[k]) < 1e-4, k


def _single(mean, scale, H=64, W=64, opacity=1.0, color=(0.2, 0.5, 0.9), bg=(1, 1, 1)):
    from gaussianavatar_b200.camera import make_camera
    K = np.array([[80.0, 0, W / 2], [0, 80.0, H / 2], [0, 0, 1]], dtype=np.float32)
    cam = make_camera(K, np.eye(4), H, W)
    P = np.asarray(mean, dtype=np.float32).reshape(-1, 3)
    n = P.shape[0]
    return dict(means3D=P, colors=np.tile(np.asarray(color, np.float32), (n, 1)), opacities=np.full((n,), opacity, np.float32),
                scales=np.full((n, 3), scale, np.float32), rotations=np.tile(np.array([1, 0, 0, 0], np.float32), (n, 1)),
                bg=np.asarray(bg, np.float32), viewmatrix=cam.world_view_transform.numpy(),
                projmatrix=cam.full_proj_transform.numpy(), tanfovx=math.tan(cam.FovX / 2), tanfovy=math.tan(cam.FovY / 2), H=H, W=W), cam


def test_known_answer_single_centred_gaussian():
    """alpha(d) = min(0.99, exp(-d^2 / 2(sigma_px^2 + 0.3))) for an isotropic Gaussian on the optical axis."""
    z, s = 2.0, 0.05
    args, cam = _single([0, 0, z], s)
    o = ro.forward(**args, precision="f64")
    assert o.get("radii")[0] > 0
    f = 80.0
    var = (f * s / z) ** 2 + 0.3
    cx, cy = o.get("xy")[0]
    assert abs(cx - 31.5) < 1e-4 and abs(cy - 31.5) < 1e-4      # ndc 0 -> ((0+1)*64-1)/2
    assert o.get("radii")[0] == math.ceil(3 * math.sqrt(var))
    ys, xs = np.mgrid[0:64, 0:64]
    d2 = (xs - cx) ** 2 + (ys - cy) ** 2
    alpha = np.minimum(0.99, np.exp(-0.5 * d2 / var))
    alpha[alpha < 1 / 255] = 0
    rect = o.get("rect")[0]
    inside = (xs // 16 >= rect[0]) & (xs // 16 < rect[2]) & (ys // 16 >= rect[1]) & (ys // 16 < rect[3])
    alpha = alpha * inside
    expect = alpha * 0.2 + (1 - alpha) * 1.0
    assert np.abs(o.image[0] - expect).max() < 1e-6


def test_known_answer_order_and_ties():
    # two co-located Gaussians, nearer one must be composited first; identical depth -> index order
    args, _ = _single([[0, 0, 2.0], [0, 0, 1.5]], 0.05)
    args["colors"] = np.array([[1, 0, 0], [0, 1, 0]], np.float32)
    o = ro.forward(**args, precision="f32")
    keys, vals = o.get("keys"), o.get("vals")
    t0 = keys >> np.uint64(32) == (keys[0] >> np.uint64(32))
    assert list(vals[t0][:2]) == [1, 0]
    args["means3D"][1, 2] = 2.0
    o = ro.forward(**args, precision="f32")
    vals = o.get("vals")
    assert list(vals[:2]) == [0, 1]
    assert np.all(np.diff(o.get("keys").astype(np.uint64)) >= 0)


def test_known_answer_culls():
    # z == 0.2 exactly is culled, slightly beyond is kept; behind the camera is culled
    args, _ = _single([[0, 0, 0.2], [0, 0, np.nextafter(np.float32(0.2), np.float32(1))], [0, 0, -1.0]], 0.01)
    o = ro.forward(**args, precision="f32")
    r = o.get("radii")
    assert r[0] == 0 and r[1] > 0 and r[2] == 0
    # far outside the frustum: rect clamps to zero area -> no instances, image == background
    args, _ = _single([[50.0, 0, 2.0]], 0.01)
    o = ro.forward(**args, precision="f32")
    assert o.num_rendered == 0 and o.get("radii")[0] == 0
    assert np.all(o.image == 1.0)
    # P = 0
    args, _ = _single(np.zeros((0, 3)), 0.01)
    o = ro.forward(**args, precision="f32")
    assert o.num_rendered == 0 and np.all(o.image == 1.0)


def test_known_answer_tile_spans():
    # a Gaussian whose 3-sigma box straddles a tile corner touches exactly 4 tiles; deep inside a tile: 1
    f, z = 80.0, 2.0
    px_to_world = z / f
    args, _ = _single([[(16 - 31.5) * px_to_world, (16 - 31.5) * px_to_world, z]], 0.02)   # centre at pixel (16,16)
    o = ro.forward(**args, precision="f32")
    assert o.get("radii")[0] <= 8 and o.get("tiles")[0] == 4
    args, _ = _single([[(24 - 31.5) * px_to_world, (24 - 31.5) * px_to_world, z]], 0.02)   # centre of tile (1,1)
    o = ro.forward(**args, precision="f32")
    assert o.get("tiles")[0] == 1
    rg = o.get("ranges")
    assert rg[1 * 4 + 1, 1] - rg[1 * 4 + 1, 0] == 1 and rg.sum() == 1
