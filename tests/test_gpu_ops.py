"""GPU parity of the individual kernels (through the C ABI) against plain PyTorch fp32 on CPU (the ops the reference
dispatches).  fp32 validation mode: max relative error <= 1e-4 (BASELINE.json north_star)."""
import pytest
import torch
import torch.nn.functional as F

from util import l2rel, maxrel, nchw, nhwc

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def ct():
    import ctagan
    from ctagan import engine, lib, ops
    return ctagan, engine, lib, ops


CONV_CASES = [
    # name, N, Ci, Co, H, W, K, stride, pad
    ("head7x7_cin1", 2, 1, 64, 38, 38, 7, 1, 0),
    ("down_s2", 1, 64, 128, 32, 32, 3, 2, 1),
    ("res3x3", 2, 256, 256, 18, 18, 3, 1, 0),
    ("disc_k4s2", 1, 64, 128, 32, 32, 4, 2, 1),
    ("disc_k4s1_odd", 1, 256, 512, 8, 8, 4, 1, 1),
    ("disc_last_co1", 2, 512, 1, 7, 7, 4, 1, 1),
    ("disc_first_cin2", 1, 2, 64, 32, 32, 4, 2, 1),
    ("reg_1x1", 1, 64, 128, 2, 2, 1, 1, 0),
    ("reg_cin96", 1, 96, 32, 16, 16, 3, 1, 1),
    ("reg_out_co2", 1, 32, 2, 16, 24, 3, 1, 1),
    ("tail7x7_co1", 1, 64, 1, 22, 22, 7, 1, 0),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_conv_fprop_bprop_wgrad(ct, case, dtype):
    ctagan, E, L, ops = ct
    name, N, Ci, Co, H, W, K, s, p = case
    T = torch.float32 if dtype == "fp32" else torch.bfloat16
    tol = TOL if dtype == "fp32" else 2e-2
    g = torch.Generator().manual_seed(hash(name) % 1000)
    x = torch.randn(N, Ci, H, W, generator=g)
    w = torch.randn(Co, Ci, K, K, generator=g) / (Ci * K * K) ** 0.5
    b = torch.randn(Co, generator=g)
    if dtype == "bf16":   # identical operand rounding on both sides: the test then checks the kernel, not bf16
        x, w = x.bfloat16().float(), w.bfloat16().float()
    x.requires_grad_(True); w.requires_grad_(True)
    y = F.conv2d(x, w, b, stride=s, padding=p)
    dy = torch.randn(y.shape, generator=g)
    if dtype == "bf16":
        dy = dy.bfloat16().float()
    y.backward(dy)

    wd, bd = w.detach().cuda(), b.cuda()
    prim = E.ConvPrim(wd, bd, s, p)
    xd = nhwc(x.detach()).cuda().to(T)
    yd = prim.fprop(xd, act=L.ACT_NONE, use_bias=True)
    assert maxrel(nchw(yd.float()), y) <= tol, ("fprop", maxrel(nchw(yd.float()), y))
    dyd = nhwc(dy).cuda().to(T)
    dxd = prim.bprop(dyd, (H, W))
    assert maxrel(nchw(dxd.float()), x.grad) <= tol, ("bprop", maxrel(nchw(dxd.float()), x.grad))
    dw, db = prim.wgrad(dyd, xd, want_bias=True)
    assert maxrel(dw, w.grad) <= tol, ("wgrad", maxrel(dw, w.grad))
    assert maxrel(db, dy.sum((0, 2, 3))) <= tol, ("bgrad", maxrel(db, dy.sum((0, 2, 3))))


@pytest.mark.parametrize("shape", [(1, 256, 128, 8, 8), (2, 128, 64, 16, 16)])
def test_conv_transpose(ct, shape):
    ctagan, E, L, ops = ct
    N, Ci, Co, H, W = shape
    g = torch.Generator().manual_seed(3)
    x = torch.randn(N, Ci, H, W, generator=g, requires_grad=True)
    wt = (torch.randn(Ci, Co, 3, 3, generator=g) / (Ci * 9) ** 0.5).requires_grad_(True)
    y = F.conv_transpose2d(x, wt, None, stride=2, padding=1, output_padding=1)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    prim = E.ConvPrim(wt.detach().cuda(), None, 2, 1)        # ConvT weight [Ci][Co] == conv weight W[O=Ci][I=Co]
    xd = nhwc(x.detach()).cuda()
    yd = prim.bprop(xd, (2 * H, 2 * W))
    assert maxrel(nchw(yd), y) <= TOL
    dyd = nhwc(dy).cuda()
    dxd = prim.fprop(dyd, use_bias=False)
    assert maxrel(nchw(dxd), x.grad) <= TOL
    dw, _ = prim.wgrad(xd, dyd)
    assert maxrel(dw, wt.grad) <= TOL


@pytest.mark.parametrize("cfg", [(2, 64, 12, 10, 1, "relu", True), (1, 256, 16, 16, 1, "none", True), (2, 8, 9, 7, 3, "relu", False),
                                 (1, 128, 8, 8, 0, "lrelu", False), (1, 1, 20, 20, 3, "none", False), (1, 128, 2, 2, 1, "relu", True)])
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_norm_act_pad_fwd_bwd(ct, cfg, dtype):
    ctagan, E, L, ops = ct
    N, C, H, W, pad, act, with_res = cfg
    T = torch.float32 if dtype == "fp32" else torch.bfloat16
    tol = TOL if dtype == "fp32" else 3e-2
    actc = {"none": L.ACT_NONE, "relu": L.ACT_RELU, "lrelu": L.ACT_LRELU}[act]
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(N, C, H, W, generator=g) * 2 + 0.7)
    res = torch.randn(N, C, H, W, generator=g) if with_res else None
    if dtype == "bf16":
        x = x.bfloat16().float()
        res = res.bfloat16().float() if with_res else None
    x.requires_grad_(True)
    use_norm = C > 1
    y = F.instance_norm(x, eps=1e-5) if use_norm else x
    y = {"none": lambda t: t, "relu": F.relu, "lrelu": lambda t: F.leaky_relu(t, 0.2)}[act](y)
    if with_res:
        res.requires_grad_(True)
        y = y + res
    if pad:
        y = F.pad(y, (pad,) * 4, mode="reflect")
    gout = torch.randn(y.shape, generator=g)
    if dtype == "bf16":
        gout = gout.bfloat16().float()
    y.backward(gout)

    xd = nhwc(x.detach()).cuda().to(T)
    stats = ops.instnorm_stats(xd) if use_norm else None
    if use_norm:
        mean = x.detach().mean((2, 3)); var = x.detach().var((2, 3), unbiased=False)
        assert maxrel(stats[..., 0], mean) <= 1e-5 and maxrel(stats[..., 1], (var + 1e-5).rsqrt()) <= 1e-5
    resd = nhwc(res.detach()).cuda().to(T) if with_res else None
    out = ops.norm_act_pad(xd, stats, actc, pad, res=resd, res_pad=0)
    assert maxrel(nchw(out.float()), y) <= tol, maxrel(nchw(out.float()), y)
    goutd = nhwc(gout).cuda().to(T)
    dx = ops.norm_act_pad_bwd(goutd, xd, stats, actc, pad)
    assert maxrel(nchw(dx.float()), x.grad) <= tol, maxrel(nchw(dx.float()), x.grad)
    if with_res:
        gres = ops.norm_act_pad_bwd(goutd, None, None, L.ACT_NONE, pad)
        assert maxrel(nchw(gres.float()), res.grad) <= tol


@pytest.mark.parametrize("cfg", [(1, 256, 64, 64, 1, 2, "none"), (2, 64, 24, 20, 1, 0, "relu"), (1, 128, 70, 70, 1, 2, "lrelu"), (3, 32, 16, 16, 0, 1, "relu"),
                                 (1, 128, 128, 128, 0, 0, "relu"), (2, 24, 100, 90, 1, 2, "none"), (1, 8, 132, 132, 0, 1, "relu")])
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_norm_bwd_skip_fusion_and_margin(ct, cfg, dtype, monkeypatch):
    """InstanceNorm backward with the reflection-pad fold, the skip-connection addend, the zero output margin and the second output
    (the skip gradient itself) against autograd -- covers both one-kernel cluster/DSMEM variants (bf16, <= 64x64 and <= 128x128) and the
    two-kernel path (fp32, larger maps)."""
    ctagan, E, L, ops = ct
    N, C, H, W, pad, out_pad, act = cfg
    if H * W > 4096 and C % 16:
        monkeypatch.setenv("CTAGAN_NORM_BWD_CLUSTER", "2")        # the opt-in 128x128 cluster variant
    T = torch.float32 if dtype == "fp32" else torch.bfloat16
    tol = TOL if dtype == "fp32" else 3e-2
    actc = {"none": L.ACT_NONE, "relu": L.ACT_RELU, "lrelu": L.ACT_LRELU}[act]
    g = torch.Generator().manual_seed(9)
    rnd = lambda *sh: (torch.randn(*sh, generator=g).bfloat16().float() if dtype == "bf16" else torch.randn(*sh, generator=g))
    x = rnd(N, C, H, W) * 1.5 + 0.3
    x = (x.bfloat16().float() if dtype == "bf16" else x).requires_grad_(True)
    gout = rnd(N, C, H + 2 * pad, W + 2 * pad)
    skip = rnd(N, C, H, W)
    y = F.instance_norm(x, eps=1e-5)
    y = {"none": lambda t: t, "relu": F.relu, "lrelu": lambda t: F.leaky_relu(t, 0.2)}[act](y)
    probe = torch.zeros(N, C, H, W, requires_grad=True)          # d(loss)/d(probe) = fold(gout) + skip
    yp = F.pad(y + probe, (pad,) * 4, mode="reflect") if pad else y + probe
    ((yp * gout).sum() + ((y + probe) * skip).sum()).backward()
    xd = nhwc(x.detach()).cuda().to(T)
    stats = ops.instnorm_stats(xd)
    dx, gsk = ops.norm_act_pad_bwd(nhwc(gout).cuda().to(T), xd, stats, actc, pad, addend=nhwc(skip).cuda().to(T), out_pad=out_pad, want_g=True)
    assert dx.shape == (N, H + 2 * out_pad, W + 2 * out_pad, C)
    inner = dx[:, out_pad:out_pad + H, out_pad:out_pad + W]
    assert maxrel(nchw(inner.float()), x.grad) <= tol, maxrel(nchw(inner.float()), x.grad)
    assert maxrel(nchw(gsk.float()), probe.grad) <= tol
    if out_pad:
        border = dx.clone()
        border[:, out_pad:out_pad + H, out_pad:out_pad + W] = 0
        assert float(border.abs().max()) == 0.0


def test_pool_upsample(ct):
    ctagan, E, L, ops = ct
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 64, 8, 12, generator=g, requires_grad=True)
    y = F.max_pool2d(x, 2)
    gy = torch.randn(y.shape, generator=g)
    add = torch.randn(x.shape, generator=g)
    y.backward(gy)
    xd = nhwc(x.detach()).cuda()
    yd = ops.maxpool2_fwd(xd)
    assert torch.equal(nchw(yd).cpu(), y.detach())
    gx = ops.maxpool2_bwd(nhwc(gy).cuda(), xd, addend=nhwc(add).cuda())
    assert maxrel(nchw(gx), x.grad + add) <= 1e-6

    for (H, W) in ((2, 2), (8, 6)):
        a = torch.randn(2, 64, H, W, generator=g, requires_grad=True)
        sk = torch.randn(2, 32, 2 * H, 2 * W, generator=g, requires_grad=True)
        u = torch.cat([F.interpolate(a, (2 * H, 2 * W), mode="bilinear"), sk], 1)
        gu = torch.randn(u.shape, generator=g)
        u.backward(gu)
        ud = ops.upsample2x_cat_fwd(nhwc(a.detach()).cuda(), nhwc(sk.detach()).cuda())
        assert maxrel(nchw(ud), u) <= 1e-6
        ga, gs = ops.upsample2x_cat_bwd(nhwc(gu).cuda(), 64)
        assert maxrel(nchw(ga), a.grad) <= 1e-5 and maxrel(nchw(gs), sk.grad) <= 1e-6


def test_warp_golden(ct, golden):
    ctagan, E, L, ops = ct
    tr = ctagan.Transformer_2D()
    for name in ("small", "sq", "tiny_flow", "oob"):
        gd = golden[f"warp.{name}"]
        src = gd["src"].cuda().requires_grad_(True)
        flow = gd["flow"].cuda().requires_grad_(True)
        o = tr(src, flow)
        (o * gd["wt"].cuda()).sum().backward()
        assert (o.cpu() - gd["out"]).abs().max() <= 2e-5, name
        assert (src.grad.cpu() - gd["gsrc"]).abs().max() <= 2e-5, name
        # the flow gradient is discontinuous where a sample sits on an integer pixel: compare away from those
        d = (flow.grad.cpu() - gd["gflow"]).abs()
        frac_bad = float((d > 1e-3 * gd["gflow"].abs().max()).float().mean())
        assert frac_bad <= 2e-3, (name, frac_bad, float(d.max()))
        sm = ctagan.smooothing_loss(flow.detach().requires_grad_(False))
        assert abs(float(sm) - float(golden[f"smooth.{name}"])) <= 1e-5 * float(golden[f"smooth.{name}"])


def test_losses(ct):
    ctagan, E, L, ops = ct
    from oracle import restate as R
    g = torch.Generator().manual_seed(2)
    a = torch.randn(2, 1, 33, 20, generator=g, requires_grad=True)
    b = torch.randn(2, 1, 33, 20, generator=g)
    ref = R.l1_loss(a, b); ref.backward()
    ad = a.detach().cuda().requires_grad_(True)
    l = ctagan.l1_loss(ad, b.cuda()); (l * 3.0).backward()
    assert abs(float(l) - float(ref)) <= 1e-6 * abs(float(ref)) and maxrel(ad.grad, 3 * a.grad) <= 1e-6
    p = torch.randn(5, 1, generator=g, requires_grad=True)
    for tgt in (1.0, 0.0):
        p.grad = None
        ref = R.mse_vs_const(p, tgt); ref.backward()
        pd = p.detach().cuda().requires_grad_(True)
        l = ctagan.mse_const(pd, tgt); l.backward()
        assert abs(float(l) - float(ref)) <= 1e-6 * abs(float(ref)) and maxrel(pd.grad, p.grad) <= 1e-6
    f = (torch.randn(2, 2, 17, 23, generator=g) * 2).requires_grad_(True)
    ref = R.smoothing_loss(f); ref.backward()
    fd = f.detach().cuda().requires_grad_(True)
    l = ctagan.smooothing_loss(fd); l.backward()
    assert abs(float(l) - float(ref)) <= 1e-5 * abs(float(ref)) and maxrel(fd.grad, f.grad) <= 1e-5
    w = torch.rand(2, 1, 16, 16, generator=g) * 2 - 1
    w[0, 0, 0, :4] = 0.0
    w.requires_grad_(True)
    b1 = torch.rand(2, 1, 16, 16, generator=g) * 2 - 1
    b2 = torch.rand(2, 1, 16, 16, generator=g) * 2 - 1
    ref = R.masked_l1(w, b1, b2); ref.backward()
    wd = w.detach().cuda().requires_grad_(True)
    l = ctagan.masked_l1_loss(wd, b1.cuda(), b2.cuda()); l.backward()
    assert abs(float(l) - float(ref)) <= 1e-6 * abs(float(ref)) and maxrel(wd.grad, w.grad) <= 1e-6
