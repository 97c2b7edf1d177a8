import sys, os, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "cta-gan_b200"))
import torch, argparse
import bench
from ctagan import trainers as TR, engine as E
args = argparse.Namespace(workload="cyc", batch=None, size=256, precision="bf16")
cfg = bench.workload_config(args)
random.seed(42); torch.manual_seed(42)
tr = TR.Cyc_Trainer(cfg)
loader = TR.SyntheticSlices(1, 256, 2, 42, tr.data_keys, pool=2)
names = {id(tr.netG_A2B._get_plan()): "G_A2B", id(tr.netG_B2A._get_plan()): "G_B2A", id(tr.netD_A._get_plan()): "D_A", id(tr.netD_B._get_plan()): "D_B"}
ogf, ogb, odf, odb = E.generator_forward, E.generator_backward, E.discriminator_forward, E.discriminator_backward
def wrap(fn, tag):
    def w(plan, *a, **k):
        print(f"{tag:6s} {names.get(id(plan))}  stream={torch.cuda.current_stream().cuda_stream:#x}")
        return fn(plan, *a, **k)
    return w
tr.step(loader.batches[0])
E.generator_forward, E.generator_backward = wrap(ogf, "G fwd"), wrap(ogb, "G bwd")
E.discriminator_forward, E.discriminator_backward = wrap(odf, "D fwd"), wrap(odb, "D bwd")
sA, sB = tr._side_streams()
print(f"main={torch.cuda.current_stream().cuda_stream:#x} sA={sA.cuda_stream:#x} sB={sB.cuda_stream:#x}")
tr.phase_G(*[loader.batches[0][k].cuda() for k in tr.data_keys])
