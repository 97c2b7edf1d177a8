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


def test_bucket_chunks_never_split_a_layer():
    """GradSync cuts the optimiser's flat gradient bucket into all-reduce chunks at parameter boundaries; a layer's bias gradient is
    written by the same weight-gradient launches as its weight gradient, so the two must land in the same chunk (the chunk's readiness is
    tracked through the layer's ConvPrim), and the chunks must tile the bucket."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cta-gan_b200"))
    import torch
    from ctagan.trainers import GradSync

    class Prim:
        def __init__(self, w, b):
            self.w, self.b, self.expected_writes, self.final_hook = w, b, None, None

    class Opt:
        pass

    sizes = [(300, 7), (500, 3), (40, 40), (900, None), (100, 1)]          # (weight elements, bias elements or None)
    params, prims = [], []
    for nw, nb in sizes:
        w = torch.zeros(nw)
        b = torch.zeros(nb) if nb else None
        params.append(w)
        if b is not None:
            params.append(b)
        prims.append(Prim(w, b))
    opt = Opt()
    opt.params, opt.prims = params, prims
    opt._prim_of = {id(q.w): q for q in prims}
    offs = [0]
    for p in params:
        offs.append(offs[-1] + ((p.numel() + 3) & ~3))
    opt._offs = offs
    opt.grad_flat = torch.zeros(offs[-1])
    for chunk_elems in (1, 64, 301, 512, 10 ** 6):
        gs = GradSync(opt)                     # world size 1: nothing is built yet
        gs.chunk = chunk_elems
        gs._build_chunks()
        assert gs.chunks[0][0] == 0 and gs.chunks[-1][1] == offs[-1]
        for (s0, e0, _), (s1, e1, _) in zip(gs.chunks, gs.chunks[1:]):
            assert e0 == s1 and e0 > s0
        start_of = {id(p): o for p, o in zip(params, offs)}
        for ci, (s, e, cprims) in enumerate(gs.chunks):
            for q in cprims:
                assert s <= start_of[id(q.w)] < e and gs._chunk_of[id(q)] == ci
                if q.b is not None:
                    assert s <= start_of[id(q.b)] < e, (chunk_elems, ci)
        assert sorted(id(q) for _, _, cp in gs.chunks for q in cp) == sorted(id(q) for q in prims)
