"""Developer probe: forward convolution of the 1-2 input channel layers -- tensor-core patch-matrix kernel (default) against the generic
CUDA-core engine (parity, incl. the fused InstanceNorm statistics) and timing; CTAGAN_THIN_TC=0 times the CUDA-core kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import _ctagan_path  # noqa
import torch
from ctagan import engine as E, lib as L, ops
from dev_thin_time import graph_time, rel  # noqa: E402  (prints that probe's table first)

CASES = [
    # name, N, Ci, Co, H, W (conv input incl. physical padding), K, stride, pad, act
    ("G head 7x7 1->64 b1", 1, 1, 64, 262, 262, 7, 1, 0, L.ACT_NONE),
    ("G head 7x7 1->64 b8", 8, 1, 64, 262, 262, 7, 1, 0, L.ACT_NONE),
    ("D first k4s2 1->64 b2", 2, 1, 64, 256, 256, 4, 2, 1, L.ACT_LRELU),
    ("D first k4s2 2->64 b2", 2, 2, 64, 256, 256, 4, 2, 1, L.ACT_LRELU),
    ("Reg first 3x3 2->32 b8", 8, 2, 32, 256, 256, 3, 1, 1, L.ACT_LRELU),
    ("G head 7x7 1->64 512 b4", 4, 1, 64, 518, 518, 7, 1, 0, L.ACT_NONE),
    ("small map 3x3 2->32 64^2 b1", 1, 2, 32, 64, 64, 3, 1, 1, L.ACT_RELU),
]
for name, N, Ci, Co, H, W, K, s, p, act in CASES:
    x = torch.randn(N, H, W, Ci, device="cuda").bfloat16()
    w = torch.randn(Co, Ci, K, K, device="cuda") / (Ci * K * K) ** 0.5
    b = torch.randn(Co, device="cuda")
    prim = E.ConvPrim(w, b, s, p)
    E.set_conv_engine("generic"); ref = prim.fprop(x, act=act, use_bias=True)
    E.set_conv_engine("auto"); y = prim.fprop(x, act=act, use_bias=True)
    y2 = prim.fprop(x, act=act, use_bias=True)
    pool = ops.ZeroPool(2 * N * Co + 8, x.device)
    ys, st = prim.fprop_stats(x, pool)
    E.set_conv_engine("generic"); yr = prim.fprop(x, use_bias=False)
    E.set_conv_engine("auto")
    yf = yr.float()
    mean = yf.mean((1, 2)); rstd = 1.0 / torch.sqrt(yf.var((1, 2), unbiased=False) + 1e-5)
    e_m = float((st[..., 0] - mean).abs().max() / mean.abs().max()); e_r = rel(st[..., 1], rstd)
    us = graph_time(lambda: prim.fprop(x, act=act, use_bias=True))

    def with_stats():
        pl = ops.ZeroPool(2 * N * Co + 8, x.device)
        return prim.fprop_stats(x, pl)
    us_s = graph_time(with_stats)
    by = x.numel() * 2 + y.numel() * 2
    print(f"{name:28s} {us:7.1f} us {by / us * 1e-3:7.1f} GB/s | +stats {us_s:7.1f} us | err {rel(y, ref):.1e} stats err {e_m:.1e} {e_r:.1e} | reproducible {bool(torch.equal(y, y2))}",
          flush=True)
