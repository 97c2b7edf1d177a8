"""ncu target (round 2): one launch of every hot kernel at the shapes of the bench workloads, L2-warm, between cudaProfilerStart/Stop.

  ncu --profile-from-start off --set full --clock-control none --cache-control none --import-source on -o gpurun_out/kernels_r2 \
      python profiles/profile_kernels.py
  ncu -i gpurun_out/kernels_r2.ncu-rep --page raw --csv > profiles/ncu_full_kernels_r2.raw.csv

Launch order (the ids in the report): for (b, s) in [(1,256), (8,256), (4,512), (16,512)]: res-block conv + statistics, its input
gradient, its weight gradient (overwrite), norm_act_pad, norm_act_pad + residual, norm backward; then the 7x7 head (forward +
statistics, weight gradient) and tail (forward, weight gradient) at (1,256) and (8,256); then the optimiser kernel of one generator."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cta-gan_b200"))
import torch  # noqa: E402

import ctagan  # noqa: E402
from ctagan import engine as E, lib as L, ops  # noqa: E402

ctagan.set_precision("bf16")
torch.manual_seed(0)
T = torch.bfloat16


def run_all():
    for b, s in [(1, 256), (8, 256), (4, 512), (16, 512)]:
        h = s // 4
        prim = E.ConvPrim((torch.randn(256, 256, 3, 3) / 48).cuda(), None, 1, 0)
        x = torch.randn(b, h + 2, h + 2, 256, device="cuda").to(T)
        pool = ops.ZeroPool(64 * b + 64, x.device)
        y, st = prim.fprop_stats(x, pool)
        dyz = torch.randn(b, h + 4, h + 4, 256, device="cuda").to(T)
        prim.bprop(dyz, (h + 2, h + 2), pad=2)
        prim.wgrad(y, x)
        ops.norm_act_pad(y, st, L.ACT_RELU, 1)
        ops.norm_act_pad(y, st, L.ACT_NONE, 1, res=x, res_pad=1)
        gout = torch.randn(b, h + 2, h + 2, 256, device="cuda").to(T)
        ops.norm_act_pad_bwd(gout, y, st, L.ACT_RELU, 1, out_pad=2, pool=ops.ZeroPool(4096 * b + 64, x.device))
    for b in (1, 8):
        head = E.ConvPrim((torch.randn(64, 1, 7, 7) / 7).cuda(), torch.zeros(64).cuda(), 1, 0)
        tail = E.ConvPrim((torch.randn(1, 64, 7, 7) / 56).cuda(), torch.zeros(1).cuda(), 1, 0)
        img = torch.randn(b, 262, 262, 1, device="cuda").to(T)
        feat = torch.randn(b, 262, 262, 64, device="cuda").to(T)
        yh, _ = head.fprop_stats(img, ops.ZeroPool(64 * b + 64, img.device))
        head.wgrad(yh, img)
        yt = tail.fprop(feat, act=L.ACT_TANH, use_bias=True)
        tail.wgrad(yt, feat, want_bias=True)
    import Model.CycleGan as M
    from ctagan.optim import FusedAdam
    net = M.Generator(1, 1).cuda()
    opt = FusedAdam(net.parameters(), 1e-4, [net])
    opt.grad_flat.normal_()
    opt.step()


run_all()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run_all()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled")
