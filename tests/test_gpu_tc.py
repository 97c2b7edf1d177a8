"""GPU parity of the tcgen05/TMEM/TMA convolution engine (forced: engine='tc') against PyTorch fp32 convolutions on the same
bf16-rounded operands, at the tensor-core-sized shapes of the hot path.  bf16 mode tolerance (north_star): <= 2e-2 outputs,
<= 5e-2 gradients; measured ~3e-3 (one bf16 ulp of the stored output) / ~1e-6 for the fp32 weight gradients."""
import pytest
import torch
import torch.nn.functional as F

from util import maxrel, nchw, nhwc

pytestmark = pytest.mark.gpu

CASES = [
    # name, N, Ci, Co, H, W, K, stride, pad  (H, W = unpadded input)
    ("res3x3_valid_b1", 1, 256, 256, 66, 66, 3, 1, 0),
    ("res3x3_valid_b3", 3, 256, 256, 66, 66, 3, 1, 0),
    ("down1_s2", 1, 64, 128, 256, 256, 3, 2, 1),
    ("down2_s2", 2, 128, 256, 128, 128, 3, 2, 1),
    ("disc1_k4s2", 2, 64, 128, 128, 128, 4, 2, 1),
    ("disc3_k4s1_odd", 2, 256, 512, 32, 32, 4, 1, 1),
    ("reg_3x3_same", 1, 64, 64, 64, 64, 3, 1, 1),
    ("wide_512", 1, 64, 64, 40, 512, 3, 1, 1),
    ("reg_up1_96to32", 2, 96, 32, 64, 64, 3, 1, 1),
    ("reg_res32_valid", 2, 32, 32, 66, 66, 3, 1, 0),
    ("reg_down2_32to64", 1, 32, 64, 128, 128, 3, 1, 1),
    ("reg_c1_1x1", 2, 64, 128, 32, 32, 1, 1, 0),
    ("reg_tiny_4x4", 8, 64, 64, 4, 4, 3, 1, 1),
    ("reg_tiny_2x2", 8, 64, 64, 2, 2, 3, 1, 1),
    ("reg_tiny_8x8_b1", 1, 64, 64, 8, 8, 3, 1, 1),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_tc_conv_engine(case):
    from ctagan import engine as E, lib as L
    name, N, Ci, Co, H, W, K, s, p = case
    g = torch.Generator().manual_seed(11)
    x = torch.randn(N, Ci, H, W, generator=g).bfloat16().float().requires_grad_(True)
    w = (torch.randn(Co, Ci, K, K, generator=g) / (Ci * K * K) ** 0.5).bfloat16().float().requires_grad_(True)
    b = torch.randn(Co, generator=g)
    y = F.conv2d(x, w, b, stride=s, padding=p)
    dy = torch.randn(y.shape, generator=g).bfloat16().float()
    y.backward(dy)
    Ho, Wo = y.shape[2], y.shape[3]
    prim = E.ConvPrim(w.detach().cuda(), b.cuda(), s, p)
    xd, dyd = nhwc(x.detach()).cuda().bfloat16(), nhwc(dy).cuda().bfloat16()
    E.set_conv_engine("tc")
    try:
        yd = prim.fprop(xd, use_bias=True)
        assert maxrel(nchw(yd.float()), y) <= 2e-2, ("fprop", maxrel(nchw(yd.float()), y))
        if s == 1 and p == 0:
            m = K - 1                      # the engine's native dgrad form: zero-margined dy, VALID conv
            dyz = torch.zeros(N, Ho + 2 * m, Wo + 2 * m, Co, device="cuda", dtype=torch.bfloat16)
            dyz[:, m:m + Ho, m:m + Wo] = dyd
            dxd = prim.bprop(dyz, (H, W), pad=m)
            dw, db = prim.wgrad(dyz, xd, want_bias=True, pad=m, gy_margin=m)
        else:
            dxd = prim.bprop(dyd, (H, W))
            if N * Ho * Wo < 512:
                E.set_conv_engine("auto")          # tiny maps: the weight gradient stays on the CUDA-core kernels
            dw, db = prim.wgrad(dyd, xd, want_bias=True)
        assert maxrel(nchw(dxd.float()), x.grad) <= 2e-2, ("dgrad", maxrel(nchw(dxd.float()), x.grad))
        assert maxrel(dw, w.grad) <= 1e-3, ("wgrad", maxrel(dw, w.grad))
        assert maxrel(db, dy.sum((0, 2, 3))) <= 1e-3
    finally:
        E.set_conv_engine("auto")


@pytest.mark.parametrize("shape", [(1, 256, 128, 64, 64), (2, 128, 64, 128, 128)])
def test_tc_conv_transpose(shape):
    from ctagan import engine as E
    N, Ci, Co, H, W = shape
    g = torch.Generator().manual_seed(3)
    x = torch.randn(N, Ci, H, W, generator=g).bfloat16().float().requires_grad_(True)
    wt = (torch.randn(Ci, Co, 3, 3, generator=g) / (Ci * 9) ** 0.5).bfloat16().float().requires_grad_(True)
    y = F.conv_transpose2d(x, wt, None, stride=2, padding=1, output_padding=1)
    dy = torch.randn(y.shape, generator=g).bfloat16().float()
    y.backward(dy)
    prim = E.ConvPrim(wt.detach().cuda(), None, 2, 1)
    xd, dyd = nhwc(x.detach()).cuda().bfloat16(), nhwc(dy).cuda().bfloat16()
    E.set_conv_engine("tc")
    try:
        yd = prim.bprop(xd, (2 * H, 2 * W))
        assert maxrel(nchw(yd.float()), y) <= 2e-2
        dxd = prim.fprop(dyd, use_bias=False)
        assert maxrel(nchw(dxd.float()), x.grad) <= 2e-2
        dw, _ = prim.wgrad(xd, dyd)
        assert maxrel(dw, wt.grad) <= 1e-3
    finally:
        E.set_conv_engine("auto")


@pytest.mark.parametrize("shape", [(1, 256, 256, 66, 66, 3, 1, 0), (3, 64, 128, 64, 64, 3, 2, 1), (2, 256, 512, 32, 32, 4, 1, 1)])
def test_tc_fused_instancenorm_statistics(shape):
    """The (mean, rstd) that come out of the conv epilogue equal the statistics of the stored output."""
    from ctagan import engine as E, ops
    N, Ci, Co, H, W, K, s, p = shape
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(N, H, W, Ci, generator=g) * 1.5 + 0.3).cuda().bfloat16()
    w = (torch.randn(Co, Ci, K, K, generator=g) / (Ci * K * K) ** 0.5).cuda()
    prim = E.ConvPrim(w, None, s, p)
    Ho, Wo = (H + 2 * p - K) // s + 1, (W + 2 * p - K) // s + 1
    pool = ops.ZeroPool(2 * N * Co + 1, x.device)
    y, stats = prim.fprop_stats(x, pool)
    assert y.shape == (N, Ho, Wo, Co) and pool.off > 0      # the fused (tcgen05) path was taken (tickets came from the pool)
    yf = y.float()
    mean = yf.mean((1, 2)); var = yf.var((1, 2), unbiased=False)
    assert float((stats[..., 0] - mean).abs().max()) <= 2e-3 * float(var.sqrt().max())
    assert maxrel(stats[..., 1], (var + 1e-5).rsqrt()) <= 2e-3
    # transposed convolution (4 output-phase launches accumulate into the same sums)
    if s == 2:
        wt = (torch.randn(Co, Ci, 3, 3, generator=g) / (Co * 9) ** 0.5).cuda()
        primT = E.ConvPrim(wt, None, 2, 1)
        xt = torch.randn(N, 32, 32, Co, generator=g).cuda().bfloat16()
        pool = ops.ZeroPool(2 * N * Ci + 1, x.device)
        yt, st = primT.bprop_stats(xt, (64, 64), pool)
        ytf = yt.float()
        assert maxrel(st[..., 1], (ytf.var((1, 2), unbiased=False) + 1e-5).rsqrt()) <= 2e-3


GROUPED_CASES = [
    # name, images per group, Ci, Co, H, W, K, stride, pad
    ("res3x3_valid", 1, 256, 256, 66, 66, 3, 1, 0),
    ("down_s2_b2", 2, 64, 128, 64, 64, 3, 2, 1),
    ("disc_k4s1_odd", 2, 256, 512, 32, 32, 4, 1, 1),
    ("reg_96to32", 1, 96, 32, 64, 64, 3, 1, 1),
]


@pytest.mark.parametrize("case", GROUPED_CASES, ids=[c[0] for c in GROUPED_CASES])
def test_grouped_launch_equals_one_launch_per_group(case):
    """ctagan_conv_gather_grouped / ctagan_conv_wgrad_grouped (two weight sets, two image groups, one launch) against the ungrouped
    entry points run per group -- both orders of the slots, since the cycle pass uses the pair in swapped order."""
    from ctagan import engine as E, ops, lib as L
    name, B, Ci, Co, H, W, K, s, p = case
    g = torch.Generator().manual_seed(5)
    prims = [E.ConvPrim((torch.randn(Co, Ci, K, K, generator=g) / (Ci * K * K) ** 0.5).cuda(), torch.randn(Co, generator=g).cuda(), s, p)
             for _ in range(2)]
    x = torch.randn(2 * B, H, W, Ci, generator=g).cuda().bfloat16()
    for order in ((0, 1), (1, 0)):
        members = [prims[k] for k in order]
        gp = E.GroupedPrim(members)
        pool = ops.ZeroPool(2 * 2 * B * Co + 8, x.device)
        y, st = gp.fprop_stats(x, pool)
        refs = [q.fprop_stats(xk, ops.ZeroPool(2 * B * Co + 8, x.device)) for q, xk in zip(members, (x[:B], x[B:]))]
        y_ref, st_ref = torch.cat([a for a, _ in refs]), torch.cat([b for _, b in refs])
        assert ops.launch_count() > 0
        assert torch.equal(y, y_ref), (order, "fprop", maxrel(y.float(), y_ref.float()))
        assert torch.equal(st, st_ref), (order, "stats", maxrel(st, st_ref))
        assert torch.equal(gp.fprop(x, use_bias=False), y_ref)
        dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(6)).cuda().bfloat16()
        dx = gp.bprop(dy, (H, W))
        dx_ref = torch.cat([q.bprop(dk, (H, W)) for q, dk in zip(members, (dy[:B], dy[B:]))])
        assert torch.equal(dx, dx_ref), (order, "dgrad", maxrel(dx.float(), dx_ref.float()))
        dw, db = gp.wgrad(dy, x, want_bias=True)
        for k, q in enumerate(members):
            dwk, dbk = q.wgrad(dy[k * B:(k + 1) * B], x[k * B:(k + 1) * B], want_bias=True)
            assert maxrel(dw[k], dwk) <= 1e-5, (order, "wgrad", k, maxrel(dw[k], dwk))      # (the split-K counts differ)
            assert maxrel(db[k], dbk) <= 1e-5, (order, "bgrad", k)


def test_grouped_cyc_schedule_matches_stream_schedule():
    """bf16 Cyc iteration with the generators / discriminators run as grouped launches against the default two-stream schedule."""
    import random
    from oracle import restate as R
    from trainer import Cyc_Trainer
    from test_gpu_steps import _cfg, _close
    out = {}
    for sched in ("streams", "grouped"):
        random.seed(42); torch.manual_seed(42)
        tr = Cyc_Trainer(_cfg("CycleGan", 128, precision="bf16", cyc_schedule=sched))
        out[sched] = []
        for it in range(2):
            rA, rB = R.synthetic_pair(1, 128, seed=700 + it, phantom=True)
            out[sched].append({k: float(v) for k, v in tr.step({"A": rA, "B": rB}).items()})
    import ctagan
    ctagan.set_precision("bf16")
    for it in range(2):
        for k in out["streams"][it]:
            assert _close(out["grouped"][it][k], out["streams"][it][k], 1e-3 if it == 0 else 3e-2), (it, k, out["grouped"][it][k], out["streams"][it][k])


@pytest.mark.parametrize("N", [1, 3, 8])
def test_fused_statistics_are_bit_reproducible(N):
    """The statistics epilogue has no floating-point atomics (per-tile partial sums, added in tile order by the last CTA): the
    same launch gives the same bits every time, and the ticket buffer is zero again afterwards."""
    from ctagan import engine as E, ops
    g = torch.Generator().manual_seed(3)
    prim = E.ConvPrim((torch.randn(256, 256, 3, 3, generator=g) / 48).cuda(), None, 1, 0)
    x = torch.randn(N, 66, 66, 256, generator=g).cuda().bfloat16()
    pool = ops.ZeroPool(64, x.device)
    y0, s0 = prim.fprop_stats(x, pool)
    for _ in range(5):
        pool.off = 0                                  # the same ticket slice again: the kernel must have re-armed it
        y1, s1 = prim.fprop_stats(x, pool)
        torch.cuda.synchronize()
        assert torch.equal(y0, y1) and torch.equal(s0, s1)
    assert float(pool.buf.abs().max()) == 0.0
