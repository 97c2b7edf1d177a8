"""world_size-2 gloo test (CPU) of the data-parallel gradient exchange used by the trainers (GradSync): after the exchange
every rank holds the mean of the per-rank gradients, parameters with no gradient are skipped, layout/views are preserved."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ctagan.trainers import GradSync, SyntheticSlices
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2, 2))]
    params[0].grad = torch.full((3, 4), float(rank + 1))
    params[1].grad = torch.arange(5.0) * (rank + 1)
    # params[2] has no gradient (a dead pre-InstanceNorm bias): must be skipped, not crash
    sync = GradSync(params)
    assert sync.world == world
    sync()
    ok = torch.allclose(params[0].grad, torch.full((3, 4), (1 + world) / 2.0)) and \
        torch.allclose(params[1].grad, torch.arange(5.0) * (1 + world) / 2.0) and params[2].grad is None
    # per-rank data streams differ (seed 42 + rank), same shapes
    a = SyntheticSlices(1, 16, 2, 42 + rank).batches[0]["A"]
    gathered = [torch.zeros_like(a) for _ in range(world)]
    dist.all_gather(gathered, a)
    ok = ok and not torch.equal(gathered[0], gathered[1])
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_grad_sync_world2_gloo():
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}
