import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _ctagan_path  # noqa
import torch
from ctagan import ops, lib as L
from dev_tc_conv import graph_time
torch.manual_seed(0)
for N in (1, 8):
    C, H = 256, 64
    x = torch.randn(N, H, H, C, device="cuda").bfloat16()
    st = ops.instnorm_stats(x)
    for pad in (0, 1):
        gout = torch.randn(N, H + 2 * pad, H + 2 * pad, C, device="cuda").bfloat16()
        us = graph_time(lambda: ops.norm_act_pad_bwd(gout, x, st, L.ACT_RELU, pad, out_pad=2))
        mb = (gout.numel() * 2 * 2 + x.numel() * 2 * 2 + N * 68 * 68 * C * 2) / 1e6
        print(f"N={N} norm_bwd(reduce+apply) pad={pad}: {us:.1f} us  ({mb:.1f} MB algorithmic -> {mb / us * 1e-3:.2f} TB/s)")
    us = graph_time(lambda: ops.norm_act_pad(x, st, L.ACT_RELU, 1))
    mb = (x.numel() * 2 + N * 66 * 66 * C * 2) / 1e6
    print(f"N={N} norm_act_pad pad=1: {us:.1f} us ({mb / us * 1e-3:.2f} TB/s)")
    us = graph_time(lambda: ops.norm_act_pad(x, st, L.ACT_NONE, 1, res=gout, res_pad=1))
    print(f"N={N} norm_act_pad+res: {us:.1f} us")
    us = graph_time(lambda: ops.instnorm_stats(x))
    print(f"N={N} instnorm_stats (memset+partial+finalize): {us:.1f} us")
