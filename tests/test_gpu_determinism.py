"""Bit reproducibility.  No kernel of the library uses floating-point atomics: every split reduction stores per-CTA partial results
in its own slot and adds the slots in a fixed order (conv statistics, weight / bias gradients of all three conv engines,
InstanceNorm statistics and backward sums, loss reductions), and the one true scatter-add (the warp's source gradient) accumulates
in 64-bit fixed point.  So the same inputs give the same bits -- whatever the CTA schedule, the stream overlap or the CUDA graph --
and a difference between two schedules of the same iteration can only be a missing dependency (a race), never "noise"."""
import random

import pytest
import torch

from util import nhwc

pytestmark = pytest.mark.gpu


def _seed(s=42):
    random.seed(s)
    torch.manual_seed(s)


WGRAD_CASES = [
    # name, N, Ci, Co, H, W, K, stride, pad      (one per weight-gradient engine path)
    ("tc_res3x3", 2, 256, 256, 34, 34, 3, 1, 0),
    ("tc_reg_32ch_bias", 2, 32, 32, 64, 64, 3, 1, 1),
    ("thin_vec_head7x7", 2, 1, 64, 70, 70, 7, 1, 0),
    ("thin_vec_tail7x7", 2, 64, 1, 70, 70, 7, 1, 0),
    ("thin_rows_disc_first", 2, 2, 64, 64, 64, 4, 2, 1),
    ("thin_vec_disc_last", 2, 512, 1, 31, 31, 4, 1, 1),
    ("simt_small_map", 2, 64, 64, 8, 8, 3, 1, 1),
]


@pytest.mark.parametrize("case", WGRAD_CASES, ids=[c[0] for c in WGRAD_CASES])
@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_weight_gradients_are_bit_reproducible(case, dtype):
    from ctagan import engine as E
    name, N, Ci, Co, H, W, K, s, p = case
    T = torch.bfloat16 if dtype == "bf16" else torch.float32
    g = torch.Generator().manual_seed(11)
    Ho, Wo = (H + 2 * p - K) // s + 1, (W + 2 * p - K) // s + 1
    x = torch.randn(N, H, W, Ci, generator=g).cuda().to(T)
    dy = torch.randn(N, Ho, Wo, Co, generator=g).cuda().to(T)
    prim = E.ConvPrim(torch.zeros(Co, Ci, K, K).cuda(), torch.zeros(Co).cuda(), s, p)
    dw0, db0 = prim.wgrad(dy, x, want_bias=True)
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (Co, Ci, K, K), dy.float().permute(0, 3, 1, 2), stride=s, padding=p)
    assert float((dw0 - ref).abs().max() / ref.abs().max()) <= 1e-3
    filler = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    for k in range(4):
        filler.fill_(k + 1)                   # different garbage in the recycled workspaces: every partial-sum slot must be written
        del filler
        dw1, db1 = prim.wgrad(dy, x, want_bias=True)
        filler = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        assert torch.equal(dw0, dw1) and torch.equal(db0, db1), (name, k)


def test_norm_and_loss_reductions_are_bit_reproducible():
    from ctagan import lib as L, ops
    g = torch.Generator().manual_seed(12)
    for T in (torch.bfloat16, torch.float32):
        x = (torch.randn(2, 128, 128, 64, generator=g) * 2 + 0.5).cuda().to(T)
        gout = torch.randn(2, 130, 130, 64, generator=g).cuda().to(T)
        s0 = ops.instnorm_stats(x)
        d0 = ops.norm_act_pad_bwd(gout, x, s0, L.ACT_RELU, 1)           # 128x128 maps: the two-kernel (reduce + apply) path
        for _ in range(4):
            assert torch.equal(ops.instnorm_stats(x), s0)
            assert torch.equal(ops.norm_act_pad_bwd(gout, x, s0, L.ACT_RELU, 1), d0)
    a = torch.randn(8, 1, 256, 256, generator=g).cuda(); b = torch.randn(8, 1, 256, 256, generator=g).cuda()
    fl = torch.randn(8, 2, 256, 256, generator=g).cuda()
    l0 = (ops.l1_fwd(a, b), ops.mse_const_fwd(a, 1.0), ops.smooth_fwd(fl), ops.masked_l1_fwd(a, b, a))
    for _ in range(4):
        l1 = (ops.l1_fwd(a, b), ops.mse_const_fwd(a, 1.0), ops.smooth_fwd(fl), ops.masked_l1_fwd(a, b, a))
        assert all(torch.equal(p, q) for p, q in zip(l0, l1))


def test_warp_source_gradient_is_bit_reproducible_and_exact():
    """gsrc (a scatter-add) in 64-bit fixed point: reproducible to the bit, and equal to the fp64 scatter-add of the same products."""
    from ctagan import ops
    g = torch.Generator().manual_seed(13)
    B, H, W = 3, 96, 128
    src = torch.randn(B, 1, H, W, generator=g).cuda()
    flow = (torch.randn(B, 2, H, W, generator=g) * 6).cuda()           # displacements beyond the 8-pixel window: the global path too
    gout = (torch.randn(B, 1, H, W, generator=g) * 3e-4).cuda()
    gs0, gf0 = ops.warp_bwd(gout, src, flow)
    for _ in range(4):
        gs1, gf1 = ops.warp_bwd(gout, src, flow)
        assert torch.equal(gs0, gs1) and torch.equal(gf0, gf1)
    s64 = src.double().cpu().requires_grad_(True)
    f64 = flow.double().cpu()
    ii, jj = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    gy = 2 * ((ii + f64[:, 0]) / (H - 1) - 0.5); gx = 2 * ((jj + f64[:, 1]) / (W - 1) - 0.5)
    out = torch.nn.functional.grid_sample(s64, torch.stack([gx, gy], -1), mode="bilinear", padding_mode="border", align_corners=True)
    out.backward(gout.double().cpu())
    err = float((gs0.cpu().double() - s64.grad).abs().max() / s64.grad.abs().max())
    assert err <= 2e-5, err              # fp32 coordinate arithmetic of the reference op order vs fp64


def _run_steps(kind, precision, steps, graph=False, size=64, batch=1):
    from oracle import restate as R
    from test_gpu_steps import _cfg
    import trainer as TR
    import ctagan
    from ctagan.graphs import GraphedTrainer
    _seed()
    if kind == "cyc":
        tr = TR.Cyc_Trainer(_cfg("CycleGan", size, batch=batch, precision=precision))
        keys = ("A", "B")
    else:
        tr = TR.Reg_Trainer(_cfg("RegGan", size, batch=batch, precision=precision))
        keys = ("A", "B")
    runner = GraphedTrainer(tr, warmup=1, replay_first=False) if graph else None
    random.seed(7)
    for i in range(steps):
        a, b = R.synthetic_pair(batch, size, seed=900 + i, phantom=True)
        if graph:
            runner.step_host(dict(zip(keys, (a, b))))
        else:
            tr.step(dict(zip(keys, (a, b))))
    torch.cuda.synchronize()
    nets = [m for m in tr.__dict__.values() if isinstance(m, torch.nn.Module) and len(list(m.parameters()))]
    state = [p.detach().clone() for m in nets for p in m.parameters()]
    losses = {k: float(v) for k, v in tr.last_losses.items()}
    ctagan.set_precision("bf16")
    return state, losses


@pytest.mark.parametrize("kind,precision,size", [("cyc", "bf16", 64), ("cyc", "fp32", 64), ("reg", "bf16", 256)])
def test_training_iterations_are_bit_reproducible(kind, precision, size):
    """Three iterations (multi-stream schedule, weight-gradient lanes, Adam) twice from the same seed: every parameter of every network
    is bit-identical, eagerly and as CUDA-graph replays."""
    s0, l0 = _run_steps(kind, precision, 3, size=size)
    s1, l1 = _run_steps(kind, precision, 3, size=size)
    bad = [i for i, (p, q) in enumerate(zip(s0, s1)) if not torch.equal(p, q)]
    assert not bad and l0 == l1, (bad[:8], l0, l1)
    s2, l2 = _run_steps(kind, precision, 3, graph=True, size=size)
    bad = [i for i, (p, q) in enumerate(zip(s0, s2)) if not torch.equal(p, q)]
    assert not bad and l0 == l2, ("graph", bad[:8], l0, l2)
