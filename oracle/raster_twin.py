"""ORACLE — TEST INFRASTRUCTURE ONLY.

Differentiable fp64 torch "twin" of the rasterizer forward, written independently of
oracle/raster_oracle.c so that ``torch.autograd`` of this file cross-checks the hand-derived backward (a-9) in the C
oracle and in the CUDA kernels.  The three deliberate deviations of upstream's backward from true autograd are encoded
as straight-through / detach tricks:
  (1) gradient passes the ``min(0.99, .)`` alpha cap as if unclamped,
  (2) the guard-band clamped ``t.x, t.y`` are constants w.r.t. ``t.z`` and gated w.r.t. ``t.x, t.y``,
  (3) conic backward uses ``1/(det^2 + 1e-7)``.
Reference call site: /root/reference gaussian_renderer/__init__.py:21-48.  PARITY UNPINNED (no upstream source here).

Small sizes only (loops over Gaussians in Python).
"""
from __future__ import annotations

import torch


class _Conic(torch.autograd.Function):
    """(a,b,c) -> (c/det, -b/det, a/det); backward with the det^2+1e-7 regulariser (deviation 3)."""

    @staticmethod
    def forward(ctx, a, b, c):
        det = a * c - b * b
        ctx.save_for_backward(a, b, c, det)
        inv = 1.0 / det
        return c * inv, -b * inv, a * inv

    @staticmethod
    def backward(ctx, gA, gB, gC):
        a, b, c, det = ctx.saved_tensors
        d2 = 1.0 / (det * det + 1e-7)
        # gB is the TRUE derivative w.r.t. the off-diagonal conic entry
        da = d2 * (-c * c * gA + b * c * gB + (det - a * c) * gC)
        dc = d2 * (-a * a * gC + a * b * gB + (det - a * c) * gA)
        db = d2 * (2 * b * c * gA - (det + 2 * b * b) * gB + 2 * a * b * gC)
        return da, db, dc


def quat_to_rot(q):
    r, x, y, z = q.unbind(-1)
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=-1)
    return R.reshape(*q.shape[:-1], 3, 3)


def twin_forward(means3D, colors, opacities, scales, rotations, bg, viewmatrix, projmatrix, tanfovx, tanfovy, H, W,
                 scale_modifier=1.0, return_aux=False):
    dt = torch.float64
    means3D = means3D.to(dt); colors = colors.to(dt); opacities = opacities.to(dt).reshape(-1)
    scales = scales.to(dt); rotations = rotations.to(dt)
    bg = torch.as_tensor(bg, dtype=dt)
    Mv = torch.as_tensor(viewmatrix, dtype=dt).reshape(4, 4).T      # column-vector matrices
    Mp = torch.as_tensor(projmatrix, dtype=dt).reshape(4, 4).T
    P = means3D.shape[0]
    gx, gy = (W + 15) // 16, (H + 15) // 16
    fx, fy = W / (2.0 * tanfovx), H / (2.0 * tanfovy)

    ones = torch.ones(P, 1, dtype=dt)
    ph = torch.cat([means3D, ones], 1)
    t = ph @ Mv.T                                                    # [P,4]
    hom = ph @ Mp.T
    p_w = 1.0 / (hom[:, 3] + 1e-7)
    ndc = hom[:, :2] * p_w[:, None]
    tz = t[:, 2]
    visible = tz > 0.2

    Rq = quat_to_rot(rotations)
    s = scale_modifier * scales
    Sigma = Rq @ torch.diag_embed(s * s) @ Rq.transpose(1, 2)

    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    txtz, tytz = t[:, 0] / tz, t[:, 1] / tz
    cx, cy = (txtz < -limx) | (txtz > limx), (tytz < -limy) | (tytz > limy)
    tx_u = txtz.clamp(-limx, limx) * tz
    ty_u = tytz.clamp(-limy, limy) * tz
    tx = torch.where(cx, tx_u.detach(), tx_u)                        # deviation (2)
    ty = torch.where(cy, ty_u.detach(), ty_u)
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zero, -fx * tx / (tz * tz), zero, fy / tz, -fy * ty / (tz * tz)], -1).reshape(P, 2, 3)
    Mat = J @ Mv[:3, :3]
    cov2 = Mat @ Sigma @ Mat.transpose(1, 2)
    a = cov2[:, 0, 0] + 0.3
    b = cov2[:, 0, 1]
    c = cov2[:, 1, 1] + 0.3
    det = a * c - b * b
    ok = visible & (det != 0)
    safe = lambda v, fill: torch.where(ok, v, torch.full_like(v, fill))
    conA, conB, conC = _Conic.apply(safe(a, 1.0), safe(b, 0.0), safe(c, 1.0))
    with torch.no_grad():
        mid = 0.5 * (a + c)
        lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
        radius = torch.ceil(3.0 * torch.sqrt(torch.clamp(lam, min=0.0))).to(torch.int64)
    pix = torch.stack([((ndc[:, 0] + 1.0) * W - 1.0) * 0.5, ((ndc[:, 1] + 1.0) * H - 1.0) * 0.5], -1)
    with torch.no_grad():
        rf = radius.to(dt)
        rminx = torch.trunc((pix[:, 0] - rf) / 16).to(torch.int64).clamp(0, gx)
        rminy = torch.trunc((pix[:, 1] - rf) / 16).to(torch.int64).clamp(0, gy)
        rmaxx = torch.trunc((pix[:, 0] + rf + 15) / 16).to(torch.int64).clamp(0, gx)
        rmaxy = torch.trunc((pix[:, 1] + rf + 15) / 16).to(torch.int64).clamp(0, gy)
        area = (rmaxx - rminx) * (rmaxy - rminy)
        ok = ok & (area > 0)
        radius = torch.where(ok, radius, torch.zeros_like(radius))
        depth32 = tz.to(torch.float32)
        idx = torch.nonzero(ok).reshape(-1)
        order = idx[torch.sort(depth32[idx], stable=True).indices]

    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    tyi, txi = ys // 16, xs // 16
    pxf, pyf = xs.to(dt), ys.to(dt)
    T = torch.ones(H, W, dtype=dt)
    C = torch.zeros(3, H, W, dtype=dt)
    done = torch.zeros(H, W, dtype=torch.bool)
    n_contrib = torch.zeros(H, W, dtype=torch.int64)
    count = torch.zeros(H, W, dtype=torch.int64)
    for g in order.tolist():
        member = (txi >= rminx[g]) & (txi < rmaxx[g]) & (tyi >= rminy[g]) & (tyi < rmaxy[g])
        act = member & ~done
        count = count + act.to(torch.int64)
        dx = pix[g, 0] - pxf
        dy = pix[g, 1] - pyf
        power = -0.5 * (conA[g] * dx * dx + conC[g] * dy * dy) - conB[g] * dx * dy
        raw = opacities[g] * torch.exp(power)
        alpha = raw + (raw.clamp(max=0.99) - raw).detach()          # deviation (1)
        valid = act & (power <= 0) & (alpha >= 1.0 / 255.0)
        test_T = T * (1 - alpha)
        newly_done = valid & (test_T < 1e-4)
        blend = valid & ~newly_done
        w = torch.where(blend, alpha * T, torch.zeros_like(T))
        C = C + colors[g][:, None, None] * w[None]
        T = torch.where(blend, test_T, T)
        n_contrib = torch.where(blend, count, n_contrib)
        done = done | newly_done
    out = C + T[None] * bg[:, None, None]
    if return_aux:
        return out, dict(radii=radius, pix=pix, depth=tz, conic=torch.stack([conA, conB, conC], -1), final_T=T,
                         n_contrib=n_contrib, rect=torch.stack([rminx, rminy, rmaxx, rmaxy], -1), order=order, cov3d=Sigma)
    return out
