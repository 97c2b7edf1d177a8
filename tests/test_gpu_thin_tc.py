"""Round-2 kernels of the layers with 1-2 channels on one side (patch-matrix tensor-core kernels, csrc/conv_tc.cu) and of the
channels-last weight-gradient storage, each against the generic CUDA-core engine / PyTorch on identical bf16 operands.
Reference layers: Model/CycleGan.py:27-28 (7x7 head, Cin = 1), :58-60 (7x7 tail, Cout = 1, tanh), :78 (discriminator layer 0),
trainer/reg.py first conv (2 -> 32) and flow head (32 -> 2)."""
import pytest
import torch
import torch.nn.functional as F

from util import maxrel, nchw, nhwc

pytestmark = pytest.mark.gpu

THIN = [
    # name, N, Ci, Co, H, W (conv input incl. physical padding), K, stride, pad
    ("head7_1to64", 1, 1, 64, 134, 134, 7, 1, 0),
    ("tail7_64to1", 2, 64, 1, 134, 134, 7, 1, 0),
    ("tail7_64to1_ragged", 2, 64, 1, 70, 102, 7, 1, 0),
    ("disc0_k4s2_2to64", 2, 2, 64, 128, 128, 4, 2, 1),
    ("disc0_k4s2_1to64", 2, 1, 64, 128, 128, 4, 2, 1),
    ("reg_first_2to32", 2, 2, 32, 128, 128, 3, 1, 1),
    ("reg_flow_32to2", 2, 32, 2, 128, 128, 3, 1, 1),
]


@pytest.mark.parametrize("case", THIN, ids=[c[0] for c in THIN])
def test_thin_layers_on_tensor_cores(case):
    """forward (+ bias, activation), fused InstanceNorm statistics where the layer has them, weight and bias gradient (overwrite and
    accumulate) against PyTorch fp32 on the same bf16-rounded operands; bit-reproducible."""
    from ctagan import engine as E, lib as L, ops
    name, N, Ci, Co, H, W, K, s, p = case
    g = torch.Generator().manual_seed(5)
    x = torch.randn(N, Ci, H, W, generator=g).bfloat16().float()
    w = (torch.randn(Co, Ci, K, K, generator=g) / (Ci * K * K) ** 0.5).bfloat16().float().requires_grad_(True)
    b = torch.randn(Co, generator=g).requires_grad_(True)
    y = F.conv2d(x, w, b, stride=s, padding=p)
    dy = torch.randn(y.shape, generator=g).bfloat16().float()
    y.backward(dy)
    prim = E.ConvPrim(w.detach().cuda(), b.detach().cuda(), s, p)
    xd, dyd = nhwc(x).cuda().bfloat16(), nhwc(dy).cuda().bfloat16()
    act = L.ACT_TANH if Co == 1 else L.ACT_LRELU
    ref = torch.tanh(y) if Co == 1 else F.leaky_relu(y, 0.2)
    yd = prim.fprop(xd, act=act, use_bias=True)
    assert maxrel(nchw(yd.float()), ref.detach()) <= 2e-2, ("fprop", maxrel(nchw(yd.float()), ref.detach()))
    assert torch.equal(yd, prim.fprop(xd, act=act, use_bias=True))
    if Ci <= 2:                                    # statistics of the raw output (the layer in front of an InstanceNorm)
        pool = ops.ZeroPool(2 * N * Co + 8, xd.device)
        raw, st = prim.fprop_stats(xd, pool)
        y0 = F.conv2d(x, w.detach(), None, stride=s, padding=p)
        mean, var = y0.mean((2, 3)), y0.var((2, 3), unbiased=False)
        assert float((st[..., 0].cpu() - mean).abs().max()) <= 2e-3 * float(y0.abs().max())
        assert maxrel(st[..., 1].cpu(), 1.0 / torch.sqrt(var + 1e-5)) <= 2e-3
    dw, db = prim.wgrad(dyd, xd, want_bias=True)
    assert maxrel(dw, w.grad) <= 1e-3 and maxrel(db, b.grad) <= 1e-3, (maxrel(dw, w.grad), maxrel(db, b.grad))
    dw2, db2 = prim.wgrad(dyd, xd, want_bias=True)
    assert torch.equal(dw, dw2) and torch.equal(db, db2)
    geom = ops.make_geom(N, H, W, Ci, y.shape[2], y.shape[3], Co, K, s, 1, p, L.ACT_NONE, ops.dt(dyd), 0)
    aw, ab = dw.clone(), db.clone()
    ops.conv_wgrad(dyd, xd, geom, True, L.ENGINE_AUTO, out_w=aw, out_b=ab, accumulate=True)
    assert torch.equal(aw, dw + dw) and torch.equal(ab, db + db)


@pytest.mark.parametrize("shape", [(1, 256, 256, 66, 3, 1), (2, 128, 256, 64, 3, 2), (2, 256, 512, 32, 4, 1), (3, 64, 96, 34, 3, 1)])
def test_channels_last_weight_gradient_equals_oihw(shape):
    """CTAGAN_WGRAD_PACKED: the same weight gradient, stored [Co][KH][KW][Ci] -- bit for bit the OIHW result, overwrite and accumulate,
    on the tcgen05 engine (cluster split-K) and on the CUDA-core engine."""
    from ctagan import lib as L, ops
    N, Ci, Co, H, K, s = shape
    p = 0 if K == 3 and s == 1 else 1
    Ho = (H + 2 * p - K) // s + 1
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn(N, H, H, Ci, device="cuda", generator=g).bfloat16()
    dy = torch.randn(N, Ho, Ho, Co, device="cuda", generator=g).bfloat16()
    geom = ops.make_geom(N, H, H, Ci, Ho, Ho, Co, K, s, 1, p, L.ACT_NONE, ops.dt(dy), 0)
    for engine in (L.ENGINE_TC, L.ENGINE_GENERIC):
        dw, _ = ops.conv_wgrad(dy, x, geom, False, engine)
        raw = torch.full((Co, K, K, Ci), float("nan"), device="cuda")
        ops.conv_wgrad(dy, x, geom, False, engine, out_w=raw, packed=True)
        assert torch.equal(raw.permute(0, 3, 1, 2), dw), engine
        ops.conv_wgrad(dy, x, geom, False, engine, out_w=raw, packed=True, accumulate=True)
        assert torch.equal(raw.permute(0, 3, 1, 2), dw + dw), engine
