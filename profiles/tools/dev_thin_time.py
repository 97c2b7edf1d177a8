"""Developer probe: weight gradient of the 1-2 channel layers -- tensor-core patch-matrix kernel (default) against the generic CUDA-core
engine (parity) and timing; run a second time with CTAGAN_THIN_TC=0 for the CUDA-core thin kernels' times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import _ctagan_path  # noqa
import torch
from ctagan import engine as E, lib as L, ops

torch.manual_seed(0)


def graph_time(fn, reps=10, iters=10):
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
    ("G head 7x7 1->64 b1", 1, 1, 64, 262, 262, 7, 1, 0),
    ("G tail 7x7 64->1 b1", 1, 64, 1, 262, 262, 7, 1, 0),
    ("G head 7x7 1->64 b8", 8, 1, 64, 262, 262, 7, 1, 0),
    ("G tail 7x7 64->1 b8", 8, 64, 1, 262, 262, 7, 1, 0),
    ("D first k4s2 1->64 b2", 2, 1, 64, 256, 256, 4, 2, 1),
    ("D first k4s2 2->64 b2", 2, 2, 64, 256, 256, 4, 2, 1),
    ("Reg first 3x3 2->32 b8", 8, 2, 32, 256, 256, 3, 1, 1),
    ("Reg out 3x3 32->2 b8", 8, 32, 2, 256, 256, 3, 1, 1),
    ("G head 7x7 1->64 512 b4", 4, 1, 64, 518, 518, 7, 1, 0),
]
for name, N, Ci, Co, H, W, K, s, p in (CASES if __name__ == "__main__" else []):
    x = torch.randn(N, H, W, Ci, device="cuda").bfloat16()
    w = torch.randn(Co, Ci, K, K, device="cuda") / (Ci * K * K) ** 0.5
    b = torch.zeros(Co, device="cuda")
    prim = E.ConvPrim(w, b, s, p)
    Ho, Wo = (H + 2 * p - K) // s + 1, (W + 2 * p - K) // s + 1
    dy = torch.randn(N, Ho, Wo, Co, device="cuda").bfloat16()
    E.set_conv_engine("generic"); wref, bref = prim.wgrad(dy, x, want_bias=True)
    E.set_conv_engine("auto"); dw, db = prim.wgrad(dy, x, want_bias=True)
    dw2, db2 = prim.wgrad(dy, x, want_bias=True)
    us = graph_time(lambda: prim.wgrad(dy, x, want_bias=True))
    g = ops.make_geom(N, H, W, Ci, Ho, Wo, Co, K, s, 1, p, L.ACT_NONE, ops.dt(dy), 0)
    acc_w, acc_b = dw.clone(), db.clone()
    ops.conv_wgrad(dy, x, g, True, L.ENGINE_AUTO, out_w=acc_w, out_b=acc_b, accumulate=True)
    by = (x.numel() + dy.numel()) * 2
    print(f"{name:28s} {us:8.1f} us  {by / us * 1e-3:7.1f} GB/s | dw err {rel(dw, wref):.1e} db err {rel(db, bref):.1e} | reproducible "
          f"{bool(torch.equal(dw, dw2) and torch.equal(db, db2))} | accumulate err {rel(acc_w, 2 * dw):.1e} {rel(acc_b, 2 * db):.1e}", flush=True)
E.set_conv_engine("auto")
