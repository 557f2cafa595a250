"""Rasterizer forward+backward on the bench scene (config 3 poses through the real model), for ncu captures; also prints the
tile-list length distribution."""
import math, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussianavatar_b200.rasterizer import GaussianRasterizationSettings, rasterize_backward, rasterize_forward
from gaussianavatar_b200.workload import Stage1Workload

wl = Stage1Workload(int(sys.argv[1]) if len(sys.argv) > 1 else 3, 2, device="cuda:0")
m = wl.model
with torch.no_grad():
    bt = dict(pose_idx=torch.tensor([0, 1], device="cuda:0"), **wl.camera_fields(2))
    means, scales, colors, _ = m._posed_gaussians(bt["pose_idx"], 5000)
c = wl._cam_dev
rs = GaussianRasterizationSettings(c.height, c.width, math.tan(c.FovX / 2), math.tan(c.FovY / 2), m.background, 1.0, c.world_view_transform,
                                   c.full_proj_transform, 0, c.camera_center, False, False)
gw = torch.randn(3, c.height, c.width, device="cuda:0")
from gaussianavatar_b200 import _lib
NIT = int(os.environ.get("GA_RASTER_ITERS", "3"))
for it in range(NIT):
    if it == 2: torch.cuda.synchronize(); _lib.profile(True)
    color, radii, ctx = rasterize_forward(means[0], colors[0], m.fix_opacity, scales[0], m.fix_rotation, rs)
    rasterize_backward(ctx, means[0].contiguous(), colors[0].contiguous(), scales[0].contiguous(), m.fix_rotation, rs, gw, want_opacity=False,
                       want_rotations=False, want_means2D=False)
torch.cuda.synchronize()
if NIT > 2:
    rep = _lib.profile_report()
    print("per-frame ms:", {k: round(t / (NIT - 2), 4) for k, (n, t) in rep.items() if "render" in k or "sort" in k})
v = ctx.views()
rg = v["ranges"].numpy().astype(np.int64)
ln = rg[:, 1] - rg[:, 0]
nc = v["n_contrib"].numpy().reshape(-1, ).astype(np.int64)
print("num_rendered", ctx.num_rendered, "tiles", ln.size, "non-empty", int((ln > 0).sum()), "len mean/median/p90/p99/max of non-empty:",
      float(ln[ln > 0].mean()), float(np.median(ln[ln > 0])), float(np.percentile(ln[ln > 0], 90)), float(np.percentile(ln[ln > 0], 99)), int(ln.max()))
H = c.height
ncm = v["n_contrib"].numpy().astype(np.int64).reshape(H // 16, 16, H // 16, 16).transpose(0, 2, 1, 3).reshape(-1, 256)
tmax = ncm.max(1)
print("per-tile max n_contrib: mean", float(tmax[ln > 0].mean()), "p90", float(np.percentile(tmax[ln > 0], 90)), "max", int(tmax.max()),
      "| sum over tiles of len", int(ln.sum()), "sum of max n_contrib", int(tmax.sum()), "| mean n_contrib over covered pixels", float(nc[nc > 0].mean()))
