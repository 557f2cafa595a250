"""CPU: pin oracle/avatar_oracle.py against fixtures generated FROM THE REFERENCE'S OWN MODULES (oracle/gen_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import avatar_oracle as ao

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def test_smpl_joint_transforms_match_reference_lbs():
    d = _load("smpl_A.npz")
    A = ao.smpl_joint_transforms(torch.tensor(d["rest_joints"]), torch.tensor(d["pose"]), torch.tensor(d["transl"]))
    assert np.abs(A.numpy() - d["A"]).max() < 2e-6
    # T-pose, zero translation (last row of the fixture) -> identity transforms
    assert np.abs(A[-1].numpy() - np.eye(4)[None]).max() < 1e-6
    # the synthetic body regenerates to the same rest joints the fixture was made with
    from gaussianavatar_b200 import synthetic as syn
    assert np.abs(syn.make_body(int(d["body_seed"])).rest_joints().numpy() - d["rest_joints"]).max() < 1e-6


@pytest.mark.parametrize("name", ["pop_s32_in16.npz", "pop_s32_in32.npz", "pop_s48_in128.npz", "pop_s32_in16_pose.npz"])
def test_pop_forward_backward_match_reference(name):
    d = _load(name)
    inp, S, B, seed = int(d["inp"]), int(d["S"]), int(d["B"]), int(d["seed"])
    p = {k: v.clone().requires_grad_(True) for k, v in ao.seeded_pop_params(seed, int(d["c_geom"]), int(d["hsize"])).items()}
    g = torch.Generator().manual_seed(seed + 1)
    geo = (torch.randn(1, int(d["c_geom"]), inp, inp, generator=g) * 0.01).requires_grad_(True)
    assert np.abs(ao.uv_coord_map(S).numpy() - d["uv"]).max() == 0
    pf = torch.tensor(d["pose_featmap"]).requires_grad_(True) if "pose_featmap" in d.files else None      # stage 2 (network.py:58)
    if pf is not None:
        assert np.array_equal((torch.randn(B, int(d["c_geom"]), inp, inp, generator=g) * 0.05).numpy(), d["pose_featmap"])      # keeps g in step
    res, sc, shs, stats = ao.pop_forward(p, geo, S, B=B, pose_featmap=pf, return_stats=True)
    for got, key in ((res, "res"), (sc, "scales"), (shs, "shs")):
        assert np.abs(got.detach().numpy() - d[key]).max() < 2e-5, key
    gr, gs, gc = (torch.randn(res.shape, generator=g), torch.randn(sc.shape, generator=g), torch.randn(shs.shape, generator=g))
    ((res * gr).sum() + (sc * gs).sum() + (shs * gc).sum()).backward()
    names = [str(n) for n in d["grad_names"]]
    norms = np.array([float(p[n].grad.norm()) for n in names])
    assert np.abs(norms - d["grad_norms"]).max() / d["grad_norms"].max() < 1e-3
    for k in d.files:
        if k.startswith("grad:"):
            n = k[5:]
            got = p[n].grad.numpy()
            got = got[:8, :8] if n.startswith("geom_proc") else got
            ref = d[k]
            if n in ("decoder.conv1.bias", "decoder.conv6N.bias"):
                # a bias in front of a training-mode BatchNorm has an exactly-zero gradient; both sides hold round-off
                assert np.abs(got).max() < 5e-3 and np.abs(ref).max() < 5e-3
                continue
            assert np.linalg.norm(got - ref) / (np.linalg.norm(ref) + 1e-30) < 2e-3, n
    assert abs(float(geo.grad.norm()) - float(d["geo_grad_norm"])) / float(d["geo_grad_norm"]) < 1e-3
    if pf is not None:
        assert np.linalg.norm(pf.grad.numpy() - d["pose_grad"]) / np.linalg.norm(d["pose_grad"]) < 1e-3
    # running-stat update of BatchNorm (momentum 0.1, unbiased variance): pins the batch statistics themselves
    m, v = stats["bn1"]
    n = B * S * S
    assert np.abs((0.1 * m).detach().numpy() - d["bn1_running_mean"]).max() < 1e-6
    assert np.abs((0.9 + 0.1 * v * n / (n - 1)).detach().numpy() - d["bn1_running_var"]).max() < 1e-6


def test_stage1_batch_dedup_property():
    """identical inputs across the batch => B=k outputs equal B=1 outputs."""
    p = ao.seeded_pop_params(3)
    geo = torch.randn(1, 64, 16, 16, generator=torch.Generator().manual_seed(1)) * 0.01
    r1 = ao.pop_forward(p, geo, 24, B=1)
    r3 = ao.pop_forward(p, geo, 24, B=3)
    for a, b in zip(r1, r3):
        assert torch.allclose(a[0], b[2], atol=5e-4)   # fp32 batch-stat round-off, amplified by BN


def test_losses_match_reference():
    d = _load("losses.npz")
    a = torch.tensor(d["img"]).requires_grad_(True)
    b = torch.tensor(d["gt"])
    l1, s = ao.l1_loss_w(a, b), ao.ssim(a, b)
    assert abs(l1.item() - float(d["l1"])) < 1e-7 and abs(s.item() - float(d["ssim"])) < 1e-6
    (0.8 * l1 + 0.2 * (1 - s)).backward()
    assert np.abs(a.grad.numpy() - d["grad"]).max() < 1e-8


def test_camera_matches_reference_dataset_math():
    from gaussianavatar_b200.camera import make_camera, scaled_intrinsics
    d, tp = _load("camera.npz"), _load("test_pose_subset.npz")
    from gaussianavatar_b200 import camera as cam_mod
    assert np.array_equal(tp["intrinsic"], cam_mod.TEST_POSE_K)
    assert np.abs(tp["extrinsic"] - cam_mod.TEST_POSE_EXTRINSIC).max() < 1e-8
    for side in (1024, 512):
        c = make_camera(scaled_intrinsics(tp["intrinsic"], side), tp["extrinsic"], side, side)
        assert np.abs(c.world_view_transform.numpy() - d[f"wvt{side}"]).max() < 1e-6
        assert np.abs(c.full_proj_transform.numpy() - d[f"full{side}"]).max() < 1e-5
        assert np.abs(c.camera_center.numpy() - d[f"center{side}"]).max() < 1e-5
        assert abs(c.FovX - d[f"fov{side}"][0]) < 1e-7


def test_tpose_identity_property():
    """live == canonical pose => cano2live == I => posed points == canonical + offset (SURVEY §8c known answers)."""
    from gaussianavatar_b200 import synthetic as syn
    a = syn.make_avatar_assets(500, 32, seed=1)
    pose = torch.tensor(syn.star_pose()).float()[None]
    A = ao.smpl_joint_transforms(a.rest_joints, pose, torch.tensor(syn.CANO_TRANSL).float()[None])
    C = ao.cano2live(A, torch.linalg.inv(a.cano_joint_mats)[None])
    assert (C - torch.eye(4)).abs().max() < 1e-5
    res = torch.randn(1, 3, 32 * 32) * 0.5
    out = ao.assemble_and_skin(res, torch.rand(1, 1, 1024), torch.rand(1, 3, 1024), a.valid_idx, a.query_points[None],
                               a.query_lbs[None], C, iteration=2000)
    expect = a.query_points + (res[0].t() * 0.02)[a.valid_idx]
    assert (out["means3D"][0] - expect).abs().max() < 1e-5


def test_unet5ds_oracle_matches_reference_fixture():
    """Stage-2 pose encoder (next scope row): the restatement of UnetNoCond5DS -- including the in-place LeakyReLU that rewrites
    the skip tensors -- against outputs / gradients of the reference module (oracle/gen_golden.py: gen_unet)."""
    from oracle import avatar_oracle as ao
    d = np.load(os.path.join(GOLD, "unet5ds_nf8_s32.npz"))
    nf, cin, cout, side, B, seed = (int(d[k]) for k in ("nf", "cin", "cout", "side", "B", "seed"))
    p = {k: v.requires_grad_(True) for k, v in ao.seeded_unet_params(seed, cin, cout, nf).items()}
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, cin, side, side, generator=g).requires_grad_(True)
    gout = torch.randn(B, cout, side, side, generator=g)
    y = ao.unet5ds_forward(p, x)
    assert np.abs(y.detach().numpy() - d["y"]).max() < 2e-5
    (y * gout).sum().backward()
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    assert rel(x.grad.numpy(), d["dx"]) < 1e-4
    assert rel(p["conv3.conv.weight"].grad.numpy(), d["d_conv3"]) < 1e-4
    assert rel(p["upconv4.up.weight"].grad.numpy(), d["d_upconv4"]) < 1e-4
    assert rel(p["upconv5.up.bias"].grad.numpy(), d["d_bias"]) < 1e-4


def test_product_pose_encoder_matches_reference_fixture():
    """gaussianavatar_b200.pose_encoder.UnetNoCond5DS (the product module; torch ops, runs on any device) against the outputs / gradients /
    running statistics of the reference's UnetNoCond5DS, and its state_dict carries the reference's names."""
    from gaussianavatar_b200.pose_encoder import UnetNoCond5DS
    d = np.load(os.path.join(GOLD, "unet5ds_nf8_s32.npz"))
    nf, cin, cout, side, B, seed = (int(d[k]) for k in ("nf", "cin", "cout", "side", "B", "seed"))
    net = UnetNoCond5DS(input_nc=cin, output_nc=cout, nf=nf, up_mode="upconv", use_dropout=False)
    ref_names = set(ao.unet5ds_param_shapes(cin, cout, nf))
    assert ref_names == {k for k, _ in net.named_parameters()}
    sd = net.state_dict()
    for k in ("conv2.bn.running_mean", "conv4.bn.running_var", "upconv1.bn.num_batches_tracked", "upconv4.bn.running_mean"):
        assert k in sd
    assert not any(k.startswith(("conv1.bn", "conv5.bn", "upconv5.bn")) for k in sd)
    missing, unexpected = net.load_state_dict(ao.seeded_unet_params(seed, cin, cout, nf), strict=False)
    assert not unexpected and all("running" in k or "num_batches" in k for k in missing)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, cin, side, side, generator=g).requires_grad_(True)
    gout = torch.randn(B, cout, side, side, generator=g)
    y = net(x)
    (y * gout).sum().backward()
    assert np.abs(y.detach().numpy() - d["y"]).max() < 2e-5
    assert np.abs(x.grad.numpy() - d["dx"]).max() < 2e-4 * max(1.0, np.abs(d["dx"]).max())
    grads = dict(net.named_parameters())
    for name, key in (("conv3.conv.weight", "d_conv3"), ("upconv4.up.weight", "d_upconv4"), ("upconv5.up.bias", "d_bias")):
        ref = d[key]
        assert np.abs(grads[name].grad.numpy() - ref).max() < 2e-4 * max(1.0, np.abs(ref).max()), name
    assert np.abs(net.conv2.bn.running_mean.numpy() - d["bn2_running_mean"]).max() < 1e-6
    assert int(net.conv2.bn.num_batches_tracked) == 1


@pytest.mark.parametrize("which", ["oracle", "product"])
def test_pose_encoder_at_the_reference_size_nf32_128(which):
    """input_nc=3, output_nc=64, nf=32 on 128 x 128 position maps (what model/avatar_model.py:139-146 builds): the oracle restatement
    and the product module against the reference module's (sub-sampled) outputs and gradients."""
    d = np.load(os.path.join(GOLD, "unet5ds_nf32_s128.npz"))
    nf, cin, cout, side, B, seed = (int(d[k]) for k in ("nf", "cin", "cout", "side", "B", "seed"))
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, cin, side, side, generator=g).requires_grad_(True)
    gout = torch.randn(B, cout, side, side, generator=g)
    params = ao.seeded_unet_params(seed, cin, cout, nf)
    if which == "oracle":
        p = {k: v.requires_grad_(True) for k, v in params.items()}
        y = ao.unet5ds_forward(p, x)
        grad_of = lambda name: p[name].grad
    else:
        from gaussianavatar_b200.pose_encoder import UnetNoCond5DS
        net = UnetNoCond5DS(input_nc=cin, output_nc=cout, nf=nf, up_mode="upconv", use_dropout=False)
        net.load_state_dict(params, strict=False)
        y = net(x)
        named = dict(net.named_parameters())
        grad_of = lambda name: named[name].grad
    (y * gout).sum().backward()
    rel = lambda a, b: np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)
    assert rel(y.detach().numpy()[:, :, ::8, ::8], d["y_sub"]) < 1e-4 and abs(float(y.detach().norm()) - float(d["y_norm"])) < 1e-3 * float(d["y_norm"])
    assert rel(x.grad.numpy()[:, :, ::4, ::4], d["dx_sub"]) < 2e-3
    assert rel(grad_of("conv3.conv.weight").numpy()[:16, :16], d["d_conv3_sub"]) < 2e-3
    assert rel(grad_of("upconv4.up.weight").numpy()[:16, :16], d["d_upconv4_sub"]) < 2e-3
    assert rel(grad_of("upconv5.up.bias").numpy(), d["d_bias"]) < 2e-3
    norms = np.array([float(grad_of(str(n)).norm()) for n in d["grad_names"]])
    assert np.abs(norms - d["grad_norms"]).max() / d["grad_norms"].max() < 2e-3
