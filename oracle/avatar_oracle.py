"""ORACLE — TEST INFRASTRUCTURE ONLY.

Plain-torch CPU restatement (fp32 or fp64) of the non-rasterizer stages of GaussianAvatar's per-frame path, each
function citing the reference lines it follows.  Pinned against the reference's own modules by oracle/gen_golden.py
(which imports them from /root/reference and writes tests/golden/*.npz); tests/test_oracle_avatar.py re-checks the
restatement against those fixtures on every run.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]


# ------------------------------------------------------------------------------------------------------------------
# a-1 / a-2: SMPL joint transforms and cano2live
# ------------------------------------------------------------------------------------------------------------------
def batch_rodrigues(rot_vecs: torch.Tensor) -> torch.Tensor:
    """submodules/smplx/lbs.py:299-333 — note the norm of the SHIFTED vector (r + 1e-8) at :317."""
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    rot_dir = rot_vecs / angle
    cos, sin = torch.cos(angle)[:, None], torch.sin(angle)[:, None]
    rx, ry, rz = torch.split(rot_dir, 1, dim=1)
    zeros = torch.zeros_like(rx)
    K = torch.cat([zeros, -rz, ry, rz, zeros, -rx, -ry, rx, zeros], dim=1).view(-1, 3, 3)
    ident = torch.eye(3, dtype=rot_vecs.dtype)[None]
    return ident + sin * K + (1 - cos) * torch.bmm(K, K)


def smpl_joint_transforms(J: torch.Tensor, pose: torch.Tensor, transl: torch.Tensor) -> torch.Tensor:
    """A [B,24,4,4] as `SMPL.forward(...).A` returns it: lbs(..., return_affine_mat=True) (lbs.py:152-252) reduced to
    what feeds A — Rodrigues (:299-333), kinematic chain (:349-405), relative transform `G - pad(G J)` (:402-403) —
    plus `A[:,:,:3,3] += transl` (body_models.py:380-383).  J [24,3] are the rest joints (constant per subject)."""
    B = pose.shape[0]
    R = batch_rodrigues(pose.reshape(-1, 3)).view(B, 24, 3, 3)
    rel = J.clone()
    rel[1:] = J[1:] - J[torch.tensor(SMPL_PARENTS[1:])]
    local = torch.zeros(B, 24, 4, 4, dtype=pose.dtype)
    local[:, :, :3, :3] = R
    local[:, :, :3, 3] = rel[None]
    local[:, :, 3, 3] = 1
    chain = [local[:, 0]]
    for i in range(1, 24):
        chain.append(torch.matmul(chain[SMPL_PARENTS[i]], local[:, i]))
    G = torch.stack(chain, dim=1)
    Jh = F.pad(J, [0, 1])[None, :, :, None].to(pose.dtype)                       # [1,24,4,1]
    A = G - F.pad(torch.matmul(G, Jh), [3, 0])
    A = A.clone()
    A[:, :, :3, 3] = A[:, :, :3, 3] + transl[:, None, :]
    return A


def cano2live(A: torch.Tensor, inv_cano: torch.Tensor) -> torch.Tensor:
    """model/avatar_model.py:296."""
    return torch.matmul(A, inv_cano)


# ------------------------------------------------------------------------------------------------------------------
# a-3 / a-4: feature net (POP_no_unet, stage 1: pose_featmap=None)
# ------------------------------------------------------------------------------------------------------------------
def uv_coord_map(S: int, dtype=torch.float32) -> torch.Tensor:
    """utils/general_utils.py:165-176 (offset=False) as called at :188: uv[n] = (row/(S-1), col/(S-1)), n = row*S+col."""
    r = torch.arange(S, dtype=dtype)
    rows, cols = torch.meshgrid(r, r, indexing="ij")
    return torch.stack([rows.reshape(-1), cols.reshape(-1)], dim=1) / (S - 1)


def softplus(x):
    return F.softplus(x)   # beta=1, threshold=20 (nn.Softplus defaults, modules.py:549)


def bn_train(x, weight, bias, eps=1e-5):
    """BatchNorm1d in training mode: biased batch variance over (B, L) (modules.py:530-546; torch semantics)."""
    mean = x.mean(dim=(0, 2), keepdim=True)
    var = x.var(dim=(0, 2), unbiased=False, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) * weight[None, :, None] + bias[None, :, None], mean.flatten(), var.flatten()


def decoder_forward(p: dict, x: torch.Tensor, return_stats: bool = False):
    """ShapeDecoder.forward (model/modules.py:554-582).  p: state_dict-style names without the 'decoder.' prefix.
    x [B, in, L] -> residuals [B,3,L], scales [B,1,L], shs [B,3,L]."""
    stats = {}

    def block(inp, conv, bn):
        y = F.conv1d(inp, p[conv + ".weight"], p[conv + ".bias"])
        z, m, v = bn_train(y, p[bn + ".weight"], p[bn + ".bias"])
        stats[bn] = (m, v)
        return softplus(z)

    x1 = block(x, "conv1", "bn1")
    x2 = block(x1, "conv2", "bn2")
    x3 = block(x2, "conv3", "bn3")
    x4 = block(x3, "conv4", "bn4")
    x5 = block(torch.cat([x, x4], dim=1), "conv5", "bn5")
    x6 = block(x5, "conv6", "bn6")
    x7 = block(x6, "conv7", "bn7")
    x8 = F.conv1d(x7, p["conv8.weight"], p["conv8.bias"])
    xN6 = block(x5, "conv6N", "bn6N")
    xN7 = block(xN6, "conv7N", "bn7N")
    xN8 = F.conv1d(xN7, p["conv8N.weight"], p["conv8N.bias"])
    xS6 = block(x5, "conv6SH", "bn6SH")
    xS7 = block(xS6, "conv7SH", "bn7SH")
    xS8 = F.conv1d(xS7, p["conv8SH.weight"], p["conv8SH.bias"])
    out = (x8, torch.sigmoid(xN8), torch.sigmoid(xS8))
    return out + (stats,) if return_stats else out


def geom_convs(p: dict, geo: torch.Tensor) -> torch.Tensor:
    """GeomConvLayers.forward with use_relu=False (modules.py:122-137, network.py:26): three 5x5 convs, no bias, no act."""
    x = F.conv2d(geo, p["geom_proc_layers.conv1.weight"], padding=2)
    x = F.conv2d(x, p["geom_proc_layers.conv2.weight"], padding=2)
    return F.conv2d(x, p["geom_proc_layers.conv3.weight"], padding=2)


def upsample_features(feat: torch.Tensor, S: int) -> torch.Tensor:
    """network.py:61-67 + modules.py:745-754: bilinear grid_sample (align_corners=False, zero padding) onto the S x S
    query grid; closed form: pixel (i,j) reads x = Wf*j/(S-1) - 0.5, y = Hf*i/(S-1) - 0.5."""
    B, C, Hf, Wf = feat.shape
    if Hf == S:
        return feat
    uv = uv_coord_map(S, feat.dtype)[None].expand(B, -1, -1)
    grid = (uv.reshape(B, S, S, 2) * 2 - 1.0).transpose(1, 2)
    return F.grid_sample(feat, grid, mode="bilinear", align_corners=False)


def pop_forward(p: dict, geo_feature: torch.Tensor, S: int, B: int = 1, pose_featmap=None, return_stats=False):
    """POP_no_unet.forward (model/network.py:39-83).  p uses the reference's state_dict names."""
    geom = geom_convs(p, geo_feature.expand(B, -1, -1, -1))
    pix = geom if pose_featmap is None else pose_featmap + geom
    pix = upsample_features(pix, S).reshape(B, pix.shape[1], -1)
    uv = uv_coord_map(S, geo_feature.dtype).t()[None].expand(B, -1, -1)
    x = torch.cat([pix, uv], dim=1)
    dp = {k[len("decoder."):]: v for k, v in p.items() if k.startswith("decoder.")}
    return decoder_forward(dp, x, return_stats=return_stats)


# ------------------------------------------------------------------------------------------------------------------
# a-5 / a-6: post-decoder assembly + Gaussian LBS
# ------------------------------------------------------------------------------------------------------------------
def assemble_and_skin(pred_res, pred_scales, pred_shs, valid_idx, query_points, query_lbs, cano2live_mats, iteration,
                      geo_feature=None, ramp=True):
    """model/avatar_model.py:308-330.  pred_* are the decoder outputs [B,3,L] / [B,1,L] / [B,3,L];
    query_points [B,N,3], query_lbs [B,N,24], cano2live_mats [B,24,4,4]."""
    res = pred_res.permute(0, 2, 1) * 0.02
    point_res = res[:, valid_idx, :].contiguous()
    cano = point_res + query_points
    pt_mats = torch.einsum("bnj,bjxy->bnxy", query_lbs, cano2live_mats)
    full_pred = torch.einsum("bnxy,bny->bnx", pt_mats[..., :3, :3], cano) + pt_mats[..., :3, 3]
    sc = pred_scales.permute(0, 2, 1)
    if ramp and iteration < 1000:
        sc = sc * 1e-3 * iteration
    shs = pred_shs.permute(0, 2, 1)[:, valid_idx, :].contiguous()
    sc = sc[:, valid_idx, :].contiguous().repeat(1, 1, 3)
    out = dict(means3D=full_pred, scales=sc, colors=shs, offset_loss=torch.mean(res ** 2), scale_loss=torch.mean(sc))
    if geo_feature is not None:
        out["geo_loss"] = torch.mean(geo_feature ** 2)
    return out


# ------------------------------------------------------------------------------------------------------------------
# f-1: image losses
# ------------------------------------------------------------------------------------------------------------------
def l1_loss_w(a, b):
    """utils/loss_utils.py:7-8."""
    return torch.abs(a - b).mean()


def ssim(img1, img2, window_size=11):
    """utils/loss_utils.py:13-53: 11x11 Gaussian window sigma 1.5, depthwise conv, zero padding, C1=0.01^2, C2=0.03^2."""
    ch = img1.size(-3)
    g = torch.tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(window_size)])
    g = (g / g.sum()).unsqueeze(1)
    w2 = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0)
    window = w2.expand(ch, 1, window_size, window_size).contiguous().to(img1.dtype)
    pad = window_size // 2
    mu1 = F.conv2d(img1, window, padding=pad, groups=ch)
    mu2 = F.conv2d(img2, window, padding=pad, groups=ch)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = F.conv2d(img1 * img1, window, padding=pad, groups=ch) - mu1_sq
    s2 = F.conv2d(img2 * img2, window, padding=pad, groups=ch) - mu2_sq
    s12 = F.conv2d(img1 * img2, window, padding=pad, groups=ch) - mu12
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu12 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
    return m.mean()


def stage1_loss(image, gt, offset_loss, geo_loss, scale_loss, lambda_dssim=0.2, lambda_scale=3e-2, wdecay_rgl=10.0):
    """train.py:71-77 (before LPIPS starts): scale + offset + 0.8 L1 + 0.2 (1-SSIM) + geo."""
    return lambda_scale * scale_loss + wdecay_rgl * offset_loss + (1.0 - lambda_dssim) * l1_loss_w(image, gt) + \
        lambda_dssim * (1.0 - ssim(image, gt)) + geo_loss


# ------------------------------------------------------------------------------------------------------------------
# deterministic parameter initialisation shared by gen_golden.py and the tests (so fixtures need not store weights)
# ------------------------------------------------------------------------------------------------------------------
def decoder_param_shapes(in_size=66, hsize=128):
    sh = {}
    for name, (o, i) in dict(conv1=(hsize, in_size), conv2=(hsize, hsize), conv3=(hsize, hsize), conv4=(hsize, hsize),
                             conv5=(hsize, hsize + in_size), conv6=(hsize, hsize), conv7=(hsize, hsize), conv8=(3, hsize),
                             conv6SH=(hsize, hsize), conv7SH=(hsize, hsize), conv8SH=(3, hsize),
                             conv6N=(hsize, hsize), conv7N=(hsize, hsize), conv8N=(1, hsize)).items():
        sh[f"decoder.{name}.weight"] = (o, i, 1)
        sh[f"decoder.{name}.bias"] = (o,)
    for bn in ("bn1", "bn2", "bn3", "bn4", "bn5", "bn6", "bn7", "bn6N", "bn7N", "bn6SH", "bn7SH"):
        sh[f"decoder.{bn}.weight"] = (hsize,)
        sh[f"decoder.{bn}.bias"] = (hsize,)
    return sh


def seeded_pop_params(seed: int, c_geom=64, hsize=128, dtype=torch.float32) -> dict:
    """Deterministic (torch CPU generator) parameters with the reference's state_dict names and kaiming-uniform-like
    magnitudes.  BN affine params are perturbed away from (1,0) so their gradients are exercised."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    for k in (1, 2, 3):
        bound = 1.0 / math.sqrt(c_geom * 25)
        p[f"geom_proc_layers.conv{k}.weight"] = (torch.rand(c_geom, c_geom, 5, 5, generator=g) * 2 - 1) * bound
    for name, shape in decoder_param_shapes(c_geom + 2, hsize).items():
        if ".bn" in name:
            p[name] = (1.0 + 0.1 * torch.randn(shape, generator=g)) if name.endswith("weight") else 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = shape[1] if len(shape) == 3 else decoder_param_shapes(c_geom + 2, hsize)[name.replace("bias", "weight")][1]
            bound = 1.0 / math.sqrt(fan_in)
            p[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return {k: v.to(dtype) for k, v in p.items()}


# ---- stage 2: pose encoder (next scope row) -----------------------------------------------------------------
def bn_train_noaffine(x, eps=1e-5):
    """nn.BatchNorm2d(affine=False) in training mode (modules.py:69,98): per-channel batch statistics, biased variance."""
    mean = x.mean(dim=(0, 2, 3), keepdim=True)
    var = x.var(dim=(0, 2, 3), unbiased=False, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps)


def unet5ds_forward(p: dict, x: torch.Tensor) -> torch.Tensor:
    """UnetNoCond5DS.forward (model/modules.py:185-232; called at model/avatar_model.py:401,589) with up_mode='upconv', no
    dropout, training-mode BatchNorm.  Blocks: Conv2DBlock = [LeakyReLU(0.2)] -> conv4x4 s2 p1 -> [BN] (modules.py:62-78),
    UpConv2DBlock = ReLU -> ConvTranspose4x4 s2 p1 -> [BN] -> cat(skip) (modules.py:81-111).
    Reference quirk reproduced: Conv2DBlock's LeakyReLU is `inplace=True`, so it also rewrites its INPUT tensor -- the skip
    connections d1..d4 that reach the up path are the leaky-ReLU'd activations, not the raw conv / BN outputs."""
    lrelu = lambda t: F.leaky_relu(t, 0.2)
    conv = lambda t, k: F.conv2d(t, p[f"{k}.conv.weight"], stride=2, padding=1)
    up = lambda t, k: F.conv_transpose2d(t, p[f"{k}.up.weight"], bias=p.get(f"{k}.up.bias"), stride=2, padding=1)
    d1 = lrelu(conv(x, "conv1"))                       # conv1: no activation, no BN; conv2's in-place LeakyReLU then rewrites it
    d2 = lrelu(bn_train_noaffine(conv(d1, "conv2")))
    d3 = lrelu(bn_train_noaffine(conv(d2, "conv3")))
    d4 = lrelu(bn_train_noaffine(conv(d3, "conv4")))
    d5 = conv(d4, "conv5")                             # no BN; consumed through upconv1's (out-of-place) ReLU only
    u1 = torch.cat([bn_train_noaffine(up(F.relu(d5), "upconv1")), d4], 1)
    u2 = torch.cat([bn_train_noaffine(up(F.relu(u1), "upconv2")), d3], 1)
    u3 = torch.cat([bn_train_noaffine(up(F.relu(u2), "upconv3")), d2], 1)
    u4 = torch.cat([bn_train_noaffine(up(F.relu(u3), "upconv4")), d1], 1)
    return up(F.relu(u4), "upconv5")                   # bias, no BN


def unet5ds_param_shapes(input_nc=3, output_nc=64, nf=64) -> dict:
    """Reference state_dict names / shapes of UnetNoCond5DS parameters (ConvTranspose2d weights are [in, out, 4, 4])."""
    s = {}
    for k, (ci, co) in enumerate(((input_nc, nf), (nf, 2 * nf), (2 * nf, 4 * nf), (4 * nf, 8 * nf), (8 * nf, 8 * nf)), start=1):
        s[f"conv{k}.conv.weight"] = (co, ci, 4, 4)
    for k, (ci, co) in enumerate(((8 * nf, 8 * nf), (16 * nf, 4 * nf), (8 * nf, 2 * nf), (4 * nf, nf), (2 * nf, output_nc)), start=1):
        s[f"upconv{k}.up.weight"] = (ci, co, 4, 4)
    s["upconv5.up.bias"] = (output_nc,)
    return s


def seeded_unet_params(seed: int, input_nc=3, output_nc=64, nf=64, dtype=torch.float32) -> dict:
    g = torch.Generator().manual_seed(seed)
    p = {}
    for name, shape in unet5ds_param_shapes(input_nc, output_nc, nf).items():
        fan = (shape[1] if ".conv." in name else shape[0]) * 16 if len(shape) == 4 else 16 * 2 * nf
        p[name] = ((torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan)).to(dtype)
    return p
