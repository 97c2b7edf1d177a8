import sys, os, random
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import _ctagan_path
import torch
from ctagan import engine as E
from oracle import restate as R
from trainer.reg import Reg
from trainer.utils import smooothing_loss
from util import maxrel
E.set_precision("fp32")
golden = torch.load("tests/golden/golden_v1.pt", weights_only=False)
def _seed(s=42):
    random.seed(s); torch.manual_seed(s)
for it in range(12):
    _seed(); sd = R.init_reg(1, 1)
    _seed(); net = Reg(256, 256, 1, 1).cuda()
    ra, rb = R.synthetic_pair(1, 256, seed=3, phantom=True)
    with torch.no_grad():
        fl = net(ra.cuda(), rb.cuda())
    e0 = maxrel(fl, golden["reg.flow_256"])
    _seed(1); wbig = torch.randn_like(sd["offset_map.output.conv2d.weight"]) * 0.05
    with torch.no_grad():
        net.offset_map.output.conv2d.weight.copy_(wbig.cuda())
    xa = ra.cuda().requires_grad_(True)
    fl = net(xa, rb.cuda())
    e1 = maxrel(fl, golden["reg.flow_256_bigw"])
    sm = smooothing_loss(fl)
    sm2 = smooothing_loss(fl.detach())
    g = float(golden["reg.smooth_bigw"])
    print(it, e0, e1, float(sm), float(sm2), g, abs(float(sm) - g) / g)
