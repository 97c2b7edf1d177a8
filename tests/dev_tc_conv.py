"""Developer probe (not a pytest): tcgen05 conv engine vs the CUDA-core engine on the res-block shapes + timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _ctagan_path  # noqa
import torch
from ctagan import engine as E, lib as L, ops

torch.manual_seed(0)


def run(N, H, W, Ci, Co, K, iters=50):
    x = (torch.randn(N, H, W, Ci, device="cuda")).bfloat16()
    w = (torch.randn(Co, Ci, K, K, device="cuda") / (Ci * K * K) ** 0.5)
    b = torch.randn(Co, device="cuda")
    prim = E.ConvPrim(w, b, 1, 0)
    E.set_conv_engine("simt"); ref = prim.fprop(x, act=L.ACT_RELU, use_bias=True).float()
    E.set_conv_engine("tc"); out = prim.fprop(x, act=L.ACT_RELU, use_bias=True).float()
    torch.cuda.synchronize()
    err = float((out - ref).abs().max() / ref.abs().max())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        prim.fprop(x, use_bias=False)
    e0.record()
    for _ in range(iters):
        prim.fprop(x, use_bias=False)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    fl = 2.0 * N * (H - K + 1) * (W - K + 1) * Ci * Co * K * K
    print(f"N={N} {H}x{W} {Ci}->{Co} k{K}: maxrel {err:.3e}  {us:.1f} us  {fl / us / 1e6:.1f} TFLOP/s", flush=True)
    # input-gradient form (zero-padded dy, flipped weights)
    E.set_conv_engine("simt"); dref = prim.bprop(x[..., :Co].contiguous() if Co <= Ci else x.repeat(1, 1, 1, Co // Ci), (H + K - 1, W + K - 1)).float()
    return err


errs = []
errs.append(run(1, 66, 66, 256, 256, 3))
errs.append(run(1, 68, 68, 256, 256, 3))
errs.append(run(8, 66, 66, 256, 256, 3))
errs.append(run(2, 34, 34, 256, 512, 4))
errs.append(run(1, 130, 130, 64, 64, 3))
errs.append(run(4, 66, 66, 128, 128, 3))
print("worst", max(errs))
assert max(errs) < 2e-2
