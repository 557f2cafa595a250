"""Seeded synthetic scenes for rasterizer parity tests (test helper, not product)."""
import math

import numpy as np
import torch

from gaussianavatar_b200.camera import TEST_POSE_EXTRINSIC, TEST_POSE_K, make_camera, scaled_intrinsics


def random_scene(P, H, W, seed=0, scale_mean=0.02, aniso=True, spread=0.6, opacity_one=False, z_extra=0.0):
    """Gaussians scattered in front of the reference's shipped camera (scaled to H x W)."""
    g = torch.Generator().manual_seed(seed)
    K = scaled_intrinsics(TEST_POSE_K, max(H, W))
    K[0, 2] = W / 2.0
    K[1, 2] = H / 2.0
    cam = make_camera(K, TEST_POSE_EXTRINSIC, H, W)
    means = (torch.rand(P, 3, generator=g) - 0.5) * torch.tensor([2 * spread, 2.4 * spread, 1.0]) + torch.tensor([0.0, -0.2, z_extra])
    colors = torch.rand(P, 3, generator=g)
    if aniso:
        scales = scale_mean * torch.exp(0.5 * torch.randn(P, 3, generator=g))
        rots = torch.randn(P, 4, generator=g)
        rots = rots / rots.norm(dim=1, keepdim=True)
    else:
        scales = (scale_mean * torch.exp(0.35 * torch.randn(P, 1, generator=g))).repeat(1, 3)
        rots = torch.zeros(P, 4)
        rots[:, 0] = 1
    opac = torch.ones(P, 1) if opacity_one else torch.rand(P, 1, generator=g) * 0.9 + 0.1
    bg = torch.tensor([1.0, 1.0, 1.0])
    return dict(means3D=means.float().contiguous(), colors=colors.float().contiguous(), opacities=opac.float().contiguous(),
                scales=scales.float().contiguous(), rotations=rots.float().contiguous(), bg=bg, cam=cam,
                tanfovx=math.tan(cam.FovX * 0.5), tanfovy=math.tan(cam.FovY * 0.5), H=H, W=W)


def oracle_args(sc):
    return dict(means3D=sc["means3D"].numpy(), colors=sc["colors"].numpy(), opacities=sc["opacities"].numpy(),
                scales=sc["scales"].numpy(), rotations=sc["rotations"].numpy(), bg=sc["bg"].numpy(),
                viewmatrix=sc["cam"].world_view_transform.numpy(), projmatrix=sc["cam"].full_proj_transform.numpy(),
                tanfovx=sc["tanfovx"], tanfovy=sc["tanfovy"], H=sc["H"], W=sc["W"])
