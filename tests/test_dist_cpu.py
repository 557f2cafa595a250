"""CPU, world_size 2, gloo: the data-parallel plumbing of the path — frame sharding and the shared-gradient all-reduce
(+ 1/world folded into the update) give the same averaged gradient as one process holding the whole batch."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gaussianavatar_b200.trainer import allreduce_gradients
    g = torch.Generator().manual_seed(100 + rank)
    flat_grad = torch.randn(1000, generator=g)
    geo_grad = torch.randn(1, 4, 8, 8, generator=g)
    local = (flat_grad.clone(), geo_grad.clone())
    allreduce_gradients([flat_grad, geo_grad, None])
    scale = 1.0 / world                      # what FusedAdam.grad_scale applies
    gathered = [torch.zeros(1000) for _ in range(world)]
    dist.all_gather(gathered, local[0])
    expect = torch.stack(gathered).mean(0)
    ok = torch.allclose(flat_grad * scale, expect, atol=1e-6)
    if rank == 0:
        out.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


def test_frame_sharding_is_disjoint_and_covers_the_global_batch():
    """Stage1Workload.frame_ids: every step, the world's ranks take disjoint consecutive slices of the frame pool."""
    from gaussianavatar_b200.workload import Stage1Workload
    ids = Stage1Workload.frame_ids
    class W:  # minimal stand-in (no CUDA): only B and num_frames are used
        B, num_frames = 2, 32
    for world in (1, 2, 4, 8):
        for step in range(5):
            seen = []
            for rank in range(world):
                seen += ids(W, step, rank, world)
            assert len(set(seen)) == len(seen) == world * W.B
            base = step * world * W.B
            assert sorted(seen) == sorted((base + j) % W.num_frames for j in range(world * W.B))


def _pose_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from types import SimpleNamespace
    from gaussianavatar_b200.trainer import Stage1Trainer
    torch.manual_seed(0)
    pose = torch.nn.Embedding(10, 72, sparse=True); transl = torch.nn.Embedding(10, 3, sparse=True)      # identical replicas
    ids = torch.tensor([2 * rank, 2 * rank + 1])                                                        # this rank's frames
    w = torch.arange(72.0) * (rank + 1)
    ((pose(ids) * w).sum() + (transl(ids) * (rank + 2.0)).sum()).backward()
    tr = Stage1Trainer.__new__(Stage1Trainer)
    tr.model = SimpleNamespace(pose=pose, transl=transl)
    tr.world, tr.group, tr.opt = world, None, SimpleNamespace(pose_op_start_iter=5)
    tr.sync_pose_gradients(epoch=3)                       # inactive: untouched
    assert pose.weight.grad.coalesce().indices().shape[1] == 2
    tr.sync_pose_gradients(epoch=9)
    g = pose.weight.grad.to_dense()
    expect = torch.zeros(10, 72)
    for r in range(world):
        expect[2 * r] = expect[2 * r + 1] = torch.arange(72.0) * (r + 1) / world
    ok = torch.allclose(g, expect) and torch.allclose(transl.weight.grad.to_dense()[:4, 0], torch.tensor([1.0, 1.0, 1.5, 1.5]))
    # SparseAdam on the synchronised gradient keeps the replicas identical
    opt = torch.optim.SparseAdam(list(pose.parameters()) + list(transl.parameters()), 5e-3)
    opt.step()
    gathered = [torch.zeros_like(pose.weight.data) for _ in range(world)]
    dist.all_gather(gathered, pose.weight.data)
    ok = ok and torch.equal(gathered[0], gathered[1])
    if rank == 0:
        out.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_pose_embedding_gradients_are_synchronised_across_ranks_gloo():
    """with pose optimisation active, ranks must apply the same sparse update to their replicas of the pose tables."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 90)
    procs = [ctx.Process(target=_pose_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


def _syncbn_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from gaussianavatar_b200.pose_encoder import UnetNoCond5DS
    g = torch.Generator().manual_seed(7)
    x_all = torch.randn(world, 3, 32, 32, generator=g)             # one frame per rank
    w_all = torch.randn(world, 5, 32, 32, generator=g)             # a different loss weight per frame
    torch.manual_seed(11)
    ref = UnetNoCond5DS(input_nc=3, output_nc=5, nf=8).double()
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    # single process, whole batch: the statistics every BatchNorm should see
    (ref(x_all.double()) * w_all.double()).sum().backward()
    # this rank: its own frame only, BatchNorm sums all-reduced over the group
    net = UnetNoCond5DS(input_nc=3, output_nc=5, nf=8).double()
    net.load_state_dict(sd)
    net.sync_group = True
    y = net(x_all[rank:rank + 1].double())
    (y * w_all[rank:rank + 1].double()).sum().backward()
    ok = True
    with torch.no_grad():
        y_ref = ref(x_all.double())[rank:rank + 1]                 # (second forward of `ref`: only its running statistics move again)
    ok &= torch.allclose(y, y_ref, rtol=1e-9, atol=1e-10)
    for (n, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        gsum = p.grad.clone()
        dist.all_reduce(gsum)                                      # what the trainer's gradient bucket does for the shared parameters
        ok &= torch.allclose(gsum, q.grad, rtol=1e-7, atol=1e-9)
    # running statistics: momentum update with the GLOBAL batch mean / unbiased variance on every rank
    ref1 = UnetNoCond5DS(input_nc=3, output_nc=5, nf=8).double()
    ref1.load_state_dict(sd)
    ref1(x_all.double())
    for (n, b), (_, c) in zip(net.named_buffers(), ref1.named_buffers()):
        ok &= torch.allclose(b.double(), c.double(), rtol=1e-9, atol=1e-12)
    if rank == 0:
        out.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_pose_encoder_batchnorm_uses_global_batch_statistics_gloo():
    """SURVEY.md §8e, stage 2 under data parallelism: with `sync_group` set, every BatchNorm of the pose encoder normalises with the
    statistics of the GLOBAL batch — 2 ranks x 1 frame give the outputs, the (summed) parameter gradients and the running statistics of
    1 process x 2 frames, exactly (fp64)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 90)
    procs = [ctx.Process(target=_syncbn_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
