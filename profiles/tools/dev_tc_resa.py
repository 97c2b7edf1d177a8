import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _ctagan_path  # noqa
import torch
from ctagan import engine as E, lib as L
from dev_tc_conv import graph_time, rel
torch.manual_seed(0)
for (N, H, Ci, Co, K) in ((1, 66, 256, 256, 3), (2, 66, 256, 256, 3), (8, 66, 256, 256, 3), (1, 68, 256, 256, 3), (2, 35, 256, 512, 4), (2, 66, 64, 64, 3)):
    x = torch.randn(N, H, H, Ci, device="cuda").bfloat16()
    w = torch.randn(Co, Ci, K, K, device="cuda") / (Ci * K * K) ** 0.5
    prim = E.ConvPrim(w, None, 1, 0)
    E.set_conv_engine("simt"); ref = prim.fprop(x, use_bias=False)
    E.set_conv_engine("tc"); out = prim.fprop(x, use_bias=False)
    us = graph_time(lambda: prim.fprop(x, use_bias=False))
    fl = 2.0 * N * (H - K + 1) ** 2 * Ci * Co * K * K
    print(f"RESA={os.environ.get('CTAGAN_TC_RESA','1')} BN={os.environ.get('CTAGAN_TC_BN','auto')} N={N} {H}x{H} {Ci}->{Co} k{K}: err {rel(out, ref):.2e}  {us:.1f} us {fl/us/1e6:.0f} TF", flush=True)
