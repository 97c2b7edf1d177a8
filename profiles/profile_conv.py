"""ncu target: the dominant kernel alone -- the res-block 3x3 256->256 convolution forward (+ fused InstanceNorm statistics) at the
batch sizes of the two bench workloads, L2-warm, between cudaProfilerStart/Stop.

  ncu --profile-from-start off --set full --clock-control none --cache-control none --import-source on -k regex:conv_tc_valid \
      -o gpurun_out/conv_full python profiles/profile_conv.py [N]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cta-gan_b200"))
import torch  # noqa: E402

from ctagan import engine as E, ops  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1
torch.manual_seed(0)
prim = E.ConvPrim((torch.randn(256, 256, 3, 3) / 48).cuda(), None, 1, 0)
x = torch.randn(N, 66, 66, 256, device="cuda").bfloat16()
for _ in range(5):
    pool = ops.ZeroPool(2 * N * 256 + 8, x.device)
    y, st = prim.fprop_stats(x, pool)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(3):
    pool = ops.ZeroPool(2 * N * 256 + 8, x.device)
    y, st = prim.fprop_stats(x, pool)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled conv N =", N, tuple(y.shape))
