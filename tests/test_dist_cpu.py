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
