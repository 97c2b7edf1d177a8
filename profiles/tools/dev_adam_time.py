"""Developer probe: time of the one-kernel optimiser step (ctagan.optim.FusedAdam) for the two Cyc generators (22.8 M parameters:
28 B/param of optimiser traffic + 4 B/param of packed bf16 weights = 730 MB) and for one discriminator."""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import _ctagan_path  # noqa
import torch
import ctagan
from ctagan.optim import FusedAdam
import Model.CycleGan as M

ctagan.set_precision("bf16")
torch.manual_seed(0)


def time_step(nets):
    params = list(itertools.chain(*[n.parameters() for n in nets]))
    opt = FusedAdam(params, 1e-4, nets)
    opt.grad_flat.normal_()
    for _ in range(3):
        opt.step()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        opt.step()
        with torch.cuda.graph(g, stream=st):
            for _ in range(5):
                opt.step()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    n = sum(p.numel() for p in params)
    print(f"{len(nets)} net(s), {n / 1e6:.1f} M params: {us:.1f} us/step, {n * 32 / us * 1e-3:.0f} GB/s of 32 B/param")


time_step([M.Generator(1, 1).cuda(), M.Generator(1, 1).cuda()])
time_step([M.Discriminator(1).cuda()])
