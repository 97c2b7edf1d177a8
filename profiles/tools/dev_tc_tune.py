import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import _ctagan_path  # noqa
import torch
from ctagan import engine as E, lib as L, ops
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from dev_tc_conv import graph_time  # noqa

torch.manual_seed(0)
for N in (1, 2, 8):
    x = torch.randn(N, 66, 66, 256, device="cuda").bfloat16()
    w = torch.randn(256, 256, 3, 3, device="cuda") * 0.02
    prim = E.ConvPrim(w, None, 1, 0)
    us = graph_time(lambda: prim.fprop(x, use_bias=False))
    fl = 2.0 * N * 64 * 64 * 256 * 256 * 9
    print(f"BN={os.environ.get('CTAGAN_TC_BN', 'auto')} N={N}: fprop {us:.1f} us {fl / us / 1e6:.0f} TF", flush=True)
