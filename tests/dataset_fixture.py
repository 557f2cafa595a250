"""Writes a small synthetic dataset folder in the reference's layout (scene/dataset_mono.py:83-96 + the assets
model/avatar_model.py:41-98 reads), deterministic from a seed: the stand-in for the licensed / un-shipped archive of the reference's
README.md:41-54.  Test helper, not product."""
import os
import pickle
from os.path import join

import numpy as np
import torch
from PIL import Image

from gaussianavatar_b200 import synthetic as syn
from gaussianavatar_b200.camera import TEST_POSE_EXTRINSIC, TEST_POSE_K, scaled_intrinsics
from gaussianavatar_b200.config import ModelParams


def write_synthetic_dataset(root, N=600, S=32, num_frames=5, side=64, inp=16, seed=0, stage2=False) -> ModelParams:
    g = np.random.default_rng(seed)
    a = syn.make_avatar_assets(N, S, seed=seed)
    pose, transl = syn.synthetic_poses(num_frames, seed=seed)
    K = scaled_intrinsics(TEST_POSE_K, side)
    project, source, testf, smpl_dir = join(root, "project"), join(root, "data"), join(root, "novel"), join(root, "smpl_models")
    for split in ("train", "test"):
        d = join(source, split)
        os.makedirs(join(d, "images"), exist_ok=True); os.makedirs(join(d, "masks"), exist_ok=True)
        for f in range(num_frames):
            img = g.integers(0, 256, size=(side, side, 3), dtype=np.uint8)
            yy, xx = np.mgrid[0:side, 0:side]
            mask = (((yy - side / 2) ** 2 + (xx - side / 2 - f) ** 2) < (side / 3) ** 2).astype(np.uint8) * 255
            mask[::7, ::5] = 100                                  # values between 0 and 255 exercise the 128 threshold
            Image.fromarray(img, "RGB").save(join(d, "images", f"{f:05d}.png"))
            Image.fromarray(mask, "L").save(join(d, "masks", f"{f:05d}.png"))
        np.savez(join(d, "cam_parms.npz"), intrinsic=K, extrinsic=TEST_POSE_EXTRINSIC)
        smpl = dict(beta=a.body.betas.clone(), body_pose=pose.clone(), trans=transl.clone())
        torch.save(smpl, join(d, "smpl_parms.pth")); torch.save(smpl, join(d, "smpl_parms_pred.pth"))
        posmap = np.zeros((S * S, 3), np.float32); posmap[a.valid_idx.numpy()] = a.query_points.numpy()
        np.savez(join(d, f"query_posemap_{S}_cano_smpl.npz"), **{f"posmap{S}": posmap.reshape(S, S, 3)})
        torch.save(a.cano_joint_mats.clone(), join(d, "smpl_cano_joint_mat.pth"))
        if stage2:
            os.makedirs(join(d, "inp_map"), exist_ok=True)
            for f in range(num_frames):
                pm = g.normal(0, 0.3, size=(inp, inp, 3)).astype(np.float32)
                np.savez(join(d, "inp_map", "inp_posemap_%s_%s.npz" % (str(inp), str(f).zfill(8))), **{f"posmap{inp}": pm})
    os.makedirs(testf, exist_ok=True)
    np.savez(join(testf, "cam_parms.npz"), intrinsic=TEST_POSE_K, extrinsic=TEST_POSE_EXTRINSIC)
    npose, ntransl = syn.synthetic_poses(4, seed=seed + 1)
    torch.save(dict(beta=a.body.betas.clone(), body_pose=npose, trans=ntransl), join(testf, "smpl_parms.pth"))
    if stage2:
        os.makedirs(join(testf, "inp_map"), exist_ok=True)
        for f in range(4):
            pm = g.normal(0, 0.3, size=(inp, inp, 3)).astype(np.float32)
            np.savez(join(testf, "inp_map", "inp_posemap_%s_%s.npz" % (str(inp), str(f).zfill(8))), **{f"posmap{inp}": pm})
    os.makedirs(join(project, "assets", "uv_masks"), exist_ok=True)
    uv = np.full(S * S, -1, np.int32); uv[a.valid_idx.numpy()] = np.arange(N) % 13776     # face ids; -1 = not on the body
    np.save(join(project, "assets", "uv_masks", f"uv_mask{S}_with_faceid_smpl.npy"), uv.reshape(S, S))
    lbs = np.zeros((S * S, 24), np.float32); lbs[a.valid_idx.numpy()] = a.query_lbs.numpy()
    np.save(join(project, "assets", f"lbs_map_smpl_{S}.npy"), lbs.reshape(S, S, 24))
    os.makedirs(smpl_dir, exist_ok=True)
    with open(join(smpl_dir, "SMPL_NEUTRAL.pkl"), "wb") as f:
        pickle.dump(dict(v_template=a.body.v_template.numpy(), shapedirs=a.body.shapedirs.numpy(), J_regressor=a.body.J_regressor.numpy(),
                         kintree_table=np.stack([np.array(syn.SMPL_PARENTS), np.arange(24)])), f)
    return ModelParams(source_path=source, model_path=join(root, "out"), project_path=project, smpl_model_path=smpl_dir, test_folder=testf,
                       train_stage=2 if stage2 else 1, query_posmap_size=S, inp_posmap_size=inp, batch_size=2)
