"""Developer probe (not a pytest): tcgen05 weight-gradient kernel -- parity against the CUDA-core engine and graph-replay timing at the
hot-path shapes (usage: python profiles/tools/dev_wgrad_time.py; CTAGAN_WG_MIN_CHUNKS / CTAGAN_WG_MAX_CLUSTER select the split plan)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import _ctagan_path  # noqa
import torch
from ctagan import engine as E, lib as L, ops

torch.manual_seed(0)


def graph_time(fn, reps=20, iters=10):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(); fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (iters * reps)


def rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max())


CASES = [
    # name, N, Ci, Co, H, W (conv input incl. physical padding), K, stride, pad
    ("res 3x3 256->256 64^2 b1", 1, 256, 256, 66, 66, 3, 1, 0),
    ("res 3x3 256->256 64^2 b8", 8, 256, 256, 66, 66, 3, 1, 0),
    ("res 3x3 256->256 128^2 b4", 4, 256, 256, 130, 130, 3, 1, 0),
    ("down1 64->128 s2 b1", 1, 64, 128, 256, 256, 3, 2, 1),
    ("down2 128->256 s2 b1", 1, 128, 256, 128, 128, 3, 2, 1),
    ("disc 64->128 k4s2 b2", 2, 64, 128, 128, 128, 4, 2, 1),
    ("disc 128->256 k4s2 b2", 2, 128, 256, 64, 64, 4, 2, 1),
    ("disc 256->512 k4s1 b2", 2, 256, 512, 32, 32, 4, 1, 1),
    ("reg 32->32 3x3 256^2 b8", 8, 32, 32, 258, 258, 3, 1, 0),
    ("reg 64->64 3x3 128^2 b8", 8, 64, 64, 130, 130, 3, 1, 0),
    ("reg 96->32 3x3 256^2 b8", 8, 96, 32, 256, 256, 3, 1, 1),
]

for name, N, Ci, Co, H, W, K, s, p in CASES:
    x = torch.randn(N, H, W, Ci, device="cuda").bfloat16()
    w = torch.randn(Co, Ci, K, K, device="cuda") / (Ci * K * K) ** 0.5
    prim = E.ConvPrim(w, None, s, p)
    Ho, Wo = (H + 2 * p - K) // s + 1, (W + 2 * p - K) // s + 1
    dy = torch.randn(N, Ho, Wo, Co, device="cuda").bfloat16()
    small = N * H * W * Ci <= 8 * 130 * 130 * 64
    if small:
        E.set_conv_engine("simt"); wref, _ = prim.wgrad(dy, x)
    E.set_conv_engine("tc"); dw, _ = prim.wgrad(dy, x)
    dw2, _ = prim.wgrad(dy, x)
    err = rel(dw, wref) if small else float("nan")
    same = bool(torch.equal(dw, dw2))
    us = graph_time(lambda: prim.wgrad(dy, x))
    g = ops.make_geom(N, H, W, Ci, Ho, Wo, Co, K, s, 1, p, L.ACT_NONE, ops.dt(dy), 0)
    raw = torch.empty(Co, K, K, Ci, device="cuda")
    ops.conv_wgrad(dy, x, g, False, L.ENGINE_TC, out_w=raw, packed=True)
    same_p = bool(torch.equal(raw.permute(0, 3, 1, 2), dw))
    us_p = graph_time(lambda: ops.conv_wgrad(dy, x, g, False, L.ENGINE_TC, out_w=raw, packed=True))
    us_a = graph_time(lambda: ops.conv_wgrad(dy, x, g, False, L.ENGINE_TC, out_w=raw, packed=True, accumulate=True))
    fl = 2.0 * N * Ho * Wo * Ci * Co * K * K
    by = (x.numel() + dy.numel()) * 2 + w.numel() * 4
    print(f"{name:32s} OIHW {us:7.1f} us | packed {us_p:7.1f} us {fl / us_p * 1e-6:7.1f} TFLOP/s | packed+acc {us_a:7.1f} us | err {err:.1e} reproducible {same} packed==OIHW {same_p}", flush=True)
E.set_conv_engine("auto")
