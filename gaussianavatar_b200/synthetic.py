"""Deterministic synthetic stand-ins for the licensed / un-shipped assets of the reference.

The reference needs the SMPL model pickle, UV masks, an LBS-weight map and canonical position maps from a OneDrive
archive (/root/reference README.md:41-54) that is not available offline.  This module produces tensors with the SAME
contracts (shapes, dtypes, value ranges, sparsity) so every kernel runs on realistic sizes:

  body      SMPL topology constants (24 joints, kinematic parents), a procedural 1.7 m T-pose skeleton, 6890
            pseudo-vertices on bone capsules, shapedirs / posedirs / J_regressor / lbs_weights of the right shapes
            (what submodules/smplx/lbs.py:152-163 consumes)
  avatar    N canonical "query" points in the star pose (arguments/__init__.py:44-53), their [N,24] skinning weights
            (<= 4 non-zeros per point like barycentric-interpolated SMPL weights, utils/general_utils.py:245-259), an
            S x S UV validity mask with exactly N valid pixels, inv(A_cano) (model/avatar_model.py:64,89)

Asset generation is one-off host work (numpy / torch CPU, float64 where it matters), not part of the per-frame path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]
NUM_JOINTS = 24
NUM_VERTS = 6890

# approximate SMPL-neutral rest joints (metres; y up, +x = subject's left)
_REST_JOINTS = np.array([
    [0.00, -0.22, 0.03], [0.07, -0.31, 0.02], [-0.07, -0.31, 0.02], [0.00, -0.11, 0.00],
    [0.10, -0.70, 0.02], [-0.10, -0.70, 0.02], [0.00, 0.02, 0.02], [0.09, -1.10, -0.02],
    [-0.09, -1.10, -0.02], [0.00, 0.08, 0.00], [0.11, -1.16, 0.10], [-0.11, -1.16, 0.10],
    [0.00, 0.29, -0.02], [0.08, 0.20, 0.00], [-0.08, 0.20, 0.00], [0.00, 0.37, 0.03],
    [0.18, 0.23, -0.01], [-0.18, 0.23, -0.01], [0.44, 0.23, -0.03], [-0.44, 0.23, -0.03],
    [0.69, 0.23, -0.02], [-0.69, 0.23, -0.02], [0.78, 0.22, -0.02], [-0.78, 0.22, -0.02]], dtype=np.float64)

# capsule radius of the bone ending at joint j (bone = parent(j) -> j); joint 0 gets a pelvis blob
_BONE_RADIUS = np.array([0.12, 0.10, 0.10, 0.13, 0.075, 0.075, 0.14, 0.05, 0.05, 0.14, 0.04, 0.04,
                         0.06, 0.08, 0.08, 0.095, 0.06, 0.06, 0.045, 0.045, 0.035, 0.035, 0.03, 0.03], dtype=np.float64)


def star_pose() -> np.ndarray:
    """Canonical pose of the reference: legs opened by 30 degrees (arguments/__init__.py:44-53)."""
    p = np.zeros(72, dtype=np.float64)
    p[5] = 30.0 / 180.0 * math.pi
    p[8] = -30.0 / 180.0 * math.pi
    return p


CANO_TRANSL = np.array([0.0, 0.30, 0.0], dtype=np.float64)   # scripts/gen_pose_map_cano_smpl.py:60-70


def rodrigues_np(r: np.ndarray) -> np.ndarray:
    """Axis-angle [J,3] -> [J,3,3] with the reference's shifted-norm convention (submodules/smplx/lbs.py:317)."""
    angle = np.linalg.norm(r + 1e-8, axis=1, keepdims=True)
    d = r / angle
    c, s = np.cos(angle)[:, :, None], np.sin(angle)[:, :, None]
    K = np.zeros((r.shape[0], 3, 3))
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -d[:, 2], d[:, 1], d[:, 2], -d[:, 0], -d[:, 1], d[:, 0]
    return np.eye(3)[None] + s * K + (1 - c) * (K @ K)


def joint_transforms_np(J: np.ndarray, pose72: np.ndarray, transl: np.ndarray | None = None) -> np.ndarray:
    """Relative joint transforms A [24,4,4] (float64) — closed form of lbs(..., return_affine_mat=True)
    (submodules/smplx/lbs.py:349-405, body_models.py:380-383).  Host-side, asset generation only."""
    R = rodrigues_np(pose72.reshape(24, 3).astype(np.float64))
    G = np.zeros((24, 4, 4))
    for j in range(24):
        L = np.eye(4)
        L[:3, :3] = R[j]
        L[:3, 3] = J[j] if j == 0 else J[j] - J[SMPL_PARENTS[j]]
        G[j] = L if j == 0 else G[SMPL_PARENTS[j]] @ L
    A = G.copy()
    for j in range(24):
        A[j, :3, 3] = G[j, :3, 3] - G[j, :3, :3] @ J[j]
    if transl is not None:
        A[:, :3, 3] += transl[None]
    return A


@dataclass
class SyntheticBody:
    """Tensor contract of the SMPL buffers `smplx.lbs.lbs` takes (submodules/smplx/lbs.py:152-163)."""
    v_template: torch.Tensor    # [6890,3]
    shapedirs: torch.Tensor     # [6890,3,10]
    posedirs: torch.Tensor      # [207, 20670]
    J_regressor: torch.Tensor   # [24,6890]
    parents: torch.Tensor       # [24] long
    lbs_weights: torch.Tensor   # [6890,24]
    betas: torch.Tensor         # [1,10]

    def rest_joints(self) -> torch.Tensor:
        """J = J_regressor (v_template + shapedirs beta): constant per subject (betas are fixed, avatar_model.py:95-98)."""
        v = self.v_template.double() + torch.einsum("l,mkl->mk", self.betas[0].double(), self.shapedirs.double())
        return (self.J_regressor.double() @ v).float()


def _sample_on_capsules(n: int, g: torch.Generator):
    """n points on the T-pose bone capsules (+ bone id and position), roughly area-uniform."""
    bones = [(SMPL_PARENTS[j], j) for j in range(1, 24)]
    length = np.array([np.linalg.norm(_REST_JOINTS[b] - _REST_JOINTS[a]) for a, b in bones])
    rad = np.array([_BONE_RADIUS[b] for _, b in bones])
    area = 2 * math.pi * rad * (length + 2 * rad)
    prob = torch.tensor(area / area.sum())
    bid = torch.multinomial(prob, n, replacement=True, generator=g)
    t = torch.rand(n, generator=g, dtype=torch.float64) * 1.2 - 0.1          # slight overshoot: rounded ends
    th = torch.rand(n, generator=g, dtype=torch.float64) * 2 * math.pi
    a = torch.tensor(_REST_JOINTS)[torch.tensor([x for x, _ in bones])[bid]]
    b = torch.tensor(_REST_JOINTS)[torch.tensor([y for _, y in bones])[bid]]
    axis = (b - a) / (b - a).norm(dim=1, keepdim=True)
    ref = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64).expand(n, 3).clone()
    par = (axis * ref).sum(1).abs() > 0.9
    ref[par] = torch.tensor([1.0, 0.0, 0.0], dtype=torch.float64)
    u = torch.linalg.cross(axis, ref)
    u = u / u.norm(dim=1, keepdim=True)
    v = torch.linalg.cross(axis, u)
    r = torch.tensor(rad)[bid]
    end = ((t < 0) | (t > 1)).double()
    r_eff = r * torch.sqrt(torch.clamp(1 - end * ((t.clamp(0, 1) - t).abs() * (b - a).norm(dim=1) / r) ** 2, min=0.05))
    pts = a + (b - a) * t[:, None] + r_eff[:, None] * (torch.cos(th)[:, None] * u + torch.sin(th)[:, None] * v)
    return pts, bid


def _skin_weights(pts: torch.Tensor, tau: float = 0.0035, k: int = 4) -> torch.Tensor:
    """softmax(-d^2/tau) over the k nearest bones; a bone's weight goes to its PARENT-side joint (the joint that moves it)."""
    bones = [(SMPL_PARENTS[j], j) for j in range(1, 24)]
    J = torch.tensor(_REST_JOINTS)
    a = J[torch.tensor([x for x, _ in bones])]
    b = J[torch.tensor([y for _, y in bones])]
    ab = b - a
    t = ((pts[:, None] - a[None]) * ab[None]).sum(-1) / (ab * ab).sum(-1)[None]
    q = a[None] + t.clamp(0, 1)[..., None] * ab[None]
    d2 = ((pts[:, None] - q) ** 2).sum(-1)                                    # [n, 23]
    near = torch.topk(-d2, k, dim=1)
    w = torch.softmax(near.values / tau, dim=1)
    W = torch.zeros(pts.shape[0], 24, dtype=torch.float64)
    owner = torch.tensor([x for x, _ in bones])                               # parent-side joint of each bone
    W.scatter_add_(1, owner[near.indices], w)
    return W


def make_body(seed: int = 0) -> SyntheticBody:
    g = torch.Generator().manual_seed(1000 + seed)
    v, _ = _sample_on_capsules(NUM_VERTS, g)
    W = _skin_weights(v)
    # J_regressor: softmax rows concentrated around each joint, then corrected so that J_regressor @ v_template
    # reproduces the rest skeleton closely (the regressor of the real model has the same property)
    d2 = ((torch.tensor(_REST_JOINTS)[:, None] - v[None]) ** 2).sum(-1)
    Jr = torch.softmax(-d2 / 0.002, dim=1)
    shapedirs = torch.randn(NUM_VERTS, 3, 10, generator=g, dtype=torch.float64) * 0.01
    posedirs = torch.randn(207, NUM_VERTS * 3, generator=g, dtype=torch.float64) * 0.001
    betas = torch.tensor([[-0.4732, -0.8652, 0.3936, 0.0133, -1.2825, -0.9797, -0.1731, 0.0270, -0.0786, -0.1593]])
    return SyntheticBody(v.float(), shapedirs.float(), posedirs.float(), Jr.float(), torch.tensor(SMPL_PARENTS, dtype=torch.long),
                         W.float(), betas.float())


@dataclass
class SyntheticAvatarAssets:
    body: SyntheticBody
    S: int                       # query_posmap_size
    N: int                       # valid UV pixels == Gaussians
    valid_idx: torch.Tensor      # [S*S] bool, exactly N True
    query_points: torch.Tensor   # [N,3] canonical (star-pose) positions
    query_lbs: torch.Tensor      # [N,24]
    cano_joint_mats: torch.Tensor  # [24,4,4] A_cano  (smpl_cano_joint_mat.pth)
    rest_joints: torch.Tensor    # [24,3]


def make_avatar_assets(N: int, S: int, seed: int = 0) -> SyntheticAvatarAssets:
    assert N <= S * S
    body = make_body(seed)
    g = torch.Generator().manual_seed(2000 + seed)
    pts, _ = _sample_on_capsules(N, g)
    W = _skin_weights(pts)
    J = body.rest_joints().double().numpy()
    A_cano = joint_transforms_np(J, star_pose(), CANO_TRANSL)
    M = torch.einsum("nj,jxy->nxy", W, torch.tensor(A_cano))
    cano = torch.einsum("nxy,ny->nx", M[:, :3, :3], pts) + M[:, :3, 3]
    valid = torch.zeros(S * S, dtype=torch.bool)
    valid[torch.randperm(S * S, generator=g)[:N]] = True
    return SyntheticAvatarAssets(body, S, N, valid, cano.float().contiguous(), W.float().contiguous(),
                                 torch.tensor(A_cano).float(), torch.tensor(J).float())


# BASELINE.json configs -> (N, S, image side)   
CONFIGS = {1: (10_000, 128, 256), 2: (50_000, 256, 512), 3: (200_000, 512, 1024), 4: (200_000, 512, 1024), 5: (500_000, 768, 2048)}


def synthetic_poses(n: int, seed: int = 0) -> tuple[torch.Tensor, torch.Tensor]:
    """Smooth random SMPL poses / translations in the range of assets/test_pose/smpl_parms.pth (used when the
    committed subset in tests/golden/ is not wanted)."""
    g = torch.Generator().manual_seed(3000 + seed)
    pose = torch.randn(n, 72, generator=g) * 0.15
    pose[:, :3] = torch.randn(n, 3, generator=g) * 0.05
    transl = torch.tensor([0.0012, 0.168, -0.021]) + torch.randn(n, 3, generator=g) * 0.01
    return pose.float(), transl.float()
