"""ncu target: the tcgen05 weight-gradient kernel alone (res-block 3x3 256->256 on 64x64 maps), L2-warm, between cudaProfilerStart/Stop.

  ncu --profile-from-start off --set full --clock-control none --cache-control none --import-source on -k regex:conv_wgrad_tc \
      -o gpurun_out/wgrad_full python profiles/profile_wgrad.py [N] [Ci] [Co] [H]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cta-gan_b200"))
import torch  # noqa: E402

from ctagan import engine as E  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1
Ci = int(sys.argv[2]) if len(sys.argv) > 2 else 256
Co = int(sys.argv[3]) if len(sys.argv) > 3 else 256
H = int(sys.argv[4]) if len(sys.argv) > 4 else 66
torch.manual_seed(0)
prim = E.ConvPrim((torch.randn(Co, Ci, 3, 3) / 48).cuda(), None, 1, 0)
x = torch.randn(N, H, H, Ci, device="cuda").bfloat16()
dy = torch.randn(N, H - 2, H - 2, Co, device="cuda").bfloat16()
E.set_conv_engine("tc")
for _ in range(2):
    dw, _ = prim.wgrad(dy, x)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(3):
    dw, _ = prim.wgrad(dy, x)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled wgrad N =", N, tuple(dw.shape))
