"""Stage-1 training step — the loop body of the reference's train.py:66-97 (before LPIPS starts at epoch 30) — over the
fused kernels, with optional frame-level data parallelism (one process per GPU; the only collective is one all-reduce of
the shared feature-net / geo_feature gradients).

    loss = lambda_scale * scale_loss + wdecay_rgl * offset_loss + (1-lambda_dssim) * L1 + lambda_dssim * (1-SSIM) + geo_loss
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .losses import image_loss
from .optim import FusedAdam


def adjust_loss_weights(init_weight, current_epoch, mode="decay", start=400, every=20):
    """Loss-weight schedule of utils/general_utils.py:260-279 as train.py:60 uses it for the offset regulariser:
    constant before `start`, then x0.85 ('decay') or x1.05 ('rise') every `every` epochs."""
    if mode == "binary":
        raise ValueError("mode='binary' leaves the weight undefined in the reference")
    if current_epoch < start:
        return init_weight * 1e-6 if mode == "rise" else init_weight
    if every == 0:
        return init_weight
    factor = 1.05 if mode == "rise" else 0.85
    return init_weight * (factor ** ((current_epoch - start) // every))


def allreduce_gradients(grads, group=None):
    """SUM all-reduce of a list of gradient tensors (asynchronously issued, then waited).  The caller divides by the world
    size (FusedAdam.grad_scale).  Works with any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests)."""
    works = [dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group, async_op=True) for g in grads if g is not None]
    for w in works:
        w.wait()


class Stage1Trainer:
    def __init__(self, model, fused_adam: bool = True, process_group=None):
        self.model = model
        self.opt = model.opt_parms
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.group = process_group
        model.training_setup()
        if fused_adam:
            # same hyper-parameters / groups / state layout as the torch.optim.Adam the reference builds (avatar_model.py:150-155)
            model.optimizer = FusedAdam([{"params": list(model.net.parameters()), "lr": self.opt.lr_net},
                                         {"params": [model.geo_feature], "lr": self.opt.lr_geomfeat}])
            model.scheduler = torch.optim.lr_scheduler.MultiStepLR(model.optimizer, self.opt.sched_milestones, gamma=0.1)
        self._comm_stream = None

    def loss(self, batch, iteration: int, epoch: int = 0):
        """Forward only: returns (loss, image) exactly as train.py:70-77 forms them."""
        m, o = self.model, self.opt
        image, points, offset_loss, geo_loss, scale_loss = m.train_stage1(batch, iteration)
        wdecay_rgl = adjust_loss_weights(o.lambda_rgl, epoch, mode="decay", start=self.epoch_start, every=20)   # train.py:60
        loss = o.lambda_scale * scale_loss + wdecay_rgl * offset_loss + image_loss(image, batch["original_image"], o.lambda_dssim) + geo_loss
        return loss, image

    epoch_start = 0    # train.py:38-45: 0 unless resuming from a checkpoint

    def sync_gradients(self):
        """All-reduce(sum) of the shared parameters' gradients; the 1/world factor is folded into the Adam kernel.
        Pose / transl embedding rows are per-frame and stay rank-local (sparse grads)."""
        if self.world == 1:
            return
        m = self.model
        allreduce_gradients([m.net.flat.grad, m.geo_feature.grad], self.group)

    def step(self, batch, iteration: int, epoch: int = 0):
        m = self.model
        loss, _ = self.loss(batch, iteration, epoch)
        m.zero_grad(epoch)
        loss.backward()
        self.sync_gradients()
        if isinstance(m.optimizer, FusedAdam):
            m.optimizer.grad_scale = 1.0 / self.world
        elif self.world > 1:
            for g in (m.net.flat.grad, m.geo_feature.grad):
                g.div_(self.world)
        m.step(epoch)
        return loss
