"""ncu target: warm up, then run exactly one eager training iteration between cudaProfilerStart/Stop.

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python profiles/profile_step.py [cyc|reg] [batch]
"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cta-gan_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from ctagan import trainers as TR  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "cyc"
default_batch = {"cyc": 1, "reg": 8, "hd": 4}[workload]
args = bench.parse.__globals__["argparse"].Namespace(workload=workload, batch=int(sys.argv[2]) if len(sys.argv) > 2 else default_batch,
                                                     size=512 if workload == "hd" else 256, precision="bf16", hd_stage=2)
cfg = bench.workload_config(args)
random.seed(42); torch.manual_seed(42)
trainer = {"cyc": TR.Cyc_Trainer, "reg": TR.Reg_Trainer, "hd": TR.Hd_Trainer_x2}[workload](cfg)
loader = TR.SyntheticSlices(cfg["batchSize"], cfg["size"], 4, 42, trainer.data_keys, pool=2)
for b in loader.batches:
    trainer.step(b)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
trainer.step(loader.batches[0])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one", workload, "step")
