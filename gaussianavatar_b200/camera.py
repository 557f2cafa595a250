"""Camera matrices in the layout the rasterizer boundary expects.

Mirrors what the reference dataset hands to ``render_batch`` (scene/dataset_mono.py:225-257, using
utils/graphics_utils.py:27-71,99-100): ``world_view_transform`` and ``full_proj_transform`` are the TRANSPOSED 4x4
matrices (row-vector convention), i.e. their flat memory is the column-major form of the usual column-vector matrix.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch


def focal2fov(focal: float, pixels: float) -> float:
    """utils/graphics_utils.py:99-100."""
    return 2.0 * math.atan(pixels / (2.0 * focal))


def world_to_view(extrinsic: np.ndarray) -> np.ndarray:
    """Column-vector world->view 4x4.  The dataset takes R = extr[:3,:3]^T, t = extr[:3,3] and getWorld2View2
    (graphics_utils.py:27-38, translate=0, scale=1) transposes R back, so the result is the extrinsic itself,
    passed through the same inverse-of-inverse in float64 and narrowed to float32."""
    Rt = np.zeros((4, 4), dtype=np.float64)
    Rt[:3, :3] = np.asarray(extrinsic, dtype=np.float32)[:3, :3].astype(np.float64)
    Rt[:3, 3] = np.asarray(extrinsic, dtype=np.float32)[:3, 3].astype(np.float64)
    Rt[3, 3] = 1.0
    c2w = np.linalg.inv(Rt)
    return np.linalg.inv(c2w).astype(np.float32)


def projection_from_K(znear: float, zfar: float, K: np.ndarray, h: int, w: int) -> torch.Tensor:
    """Off-centre perspective matrix from intrinsics (graphics_utils.py:41-71, K is not None branch)."""
    near_fx, near_fy = znear / float(K[0, 0]), znear / float(K[1, 1])
    left, right = -(w - float(K[0, 2])) * near_fx, float(K[0, 2]) * near_fx
    bottom, top = (float(K[1, 2]) - h) * near_fy, float(K[1, 2]) * near_fy
    Pm = torch.zeros(4, 4)
    Pm[0, 0] = 2.0 * znear / (right - left)
    Pm[1, 1] = 2.0 * znear / (top - bottom)
    Pm[0, 2] = (right + left) / (right - left)
    Pm[1, 2] = (top + bottom) / (top - bottom)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    return Pm


@dataclass
class Camera:
    """One frame's camera in the batch-dict vocabulary of scene/dataset_mono.py:238-255."""
    FovX: float
    FovY: float
    height: int
    width: int
    world_view_transform: torch.Tensor   # [4,4] transposed
    full_proj_transform: torch.Tensor    # [4,4] transposed
    camera_center: torch.Tensor          # [3]

    def to(self, device):
        return Camera(self.FovX, self.FovY, self.height, self.width, self.world_view_transform.to(device),
                      self.full_proj_transform.to(device), self.camera_center.to(device))


def make_camera(K: np.ndarray, extrinsic: np.ndarray, height: int, width: int, znear: float = 0.01, zfar: float = 100.0) -> Camera:
    """K [3,3], extrinsic [4,4] world->camera (as in cam_parms.npz) -> Camera (dataset_mono.py:225-255)."""
    K = np.asarray(K, dtype=np.float32).reshape(3, 3)
    FovY = focal2fov(float(K[1, 1]), height)
    FovX = focal2fov(float(K[0, 0]), width)
    wvt = torch.tensor(world_to_view(extrinsic)).transpose(0, 1)
    proj = projection_from_K(znear, zfar, K, height, width).transpose(0, 1)
    full = (wvt.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0)
    center = wvt.inverse()[3, :3]
    return Camera(FovX, FovY, int(height), int(width), wvt.contiguous(), full.contiguous(), center.contiguous())


def scaled_intrinsics(K: np.ndarray, size: int, base: int = 1024) -> np.ndarray:
    """Scale the shipped 1024^2 intrinsics (assets/test_pose/cam_parms.npz) to another square resolution."""
    K = np.asarray(K, dtype=np.float32).copy()
    s = size / float(base)
    K[0, 0] *= s; K[1, 1] *= s; K[0, 2] *= s; K[1, 2] *= s
    return K


# The one camera the reference ships (assets/test_pose/cam_parms.npz), reproduced as constants so that nothing at
# run time needs /root/reference.  Values verified against the file by oracle/gen_golden.py.
TEST_POSE_K = np.array([[1100.0, 0.0, 512.0], [0.0, 1100.0, 512.0], [0.0, 0.0, 1.0]], dtype=np.float32)
TEST_POSE_EXTRINSIC = np.array([[0.99970485, 0.0, 0.02429441, -0.06073601],
                                [-0.00589733, -0.97009033, 0.24267256, -0.3156543],
                                [0.02356777, -0.24274421, -0.96980401, 2.49733328],
                                [0.0, 0.0, 0.0, 1.0]], dtype=np.float64)
