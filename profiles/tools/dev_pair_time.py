import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _ctagan_path  # noqa
import torch
from ctagan import engine as E, ops
from dev_tc_conv import graph_time
torch.manual_seed(0)
prim = E.ConvPrim((torch.randn(256, 256, 3, 3) / 48).cuda(), None, 1, 0)
ref = None
for N in (1, 2, 4, 8):
    x = torch.randn(N, 66, 66, 256, device="cuda").bfloat16()
    def run():
        pool = ops.ZeroPool(2 * N * 256 + 8, x.device)
        return prim.fprop_stats(x, pool)
    y, st = run()
    us = graph_time(run)
    fl = 2 * N * 64 * 64 * 256 * 256 * 9
    print(f"PAIR={os.environ.get('CTAGAN_TC_PAIR','0')} BN={os.environ.get('CTAGAN_TC_BN','auto')} N={N}: {us:.1f} us  {fl / us / 1e6:.0f} TFLOP/s  checksum {float(y.float().abs().sum()):.6e} {float(st.abs().sum()):.6e}")
