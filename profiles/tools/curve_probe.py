"""Developer probe: running-mean loss curves of the product trainers (graph / eager, bf16 / fp32) against tests/golden/curves_v1.pt."""
import os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _ctagan_path  # noqa
import torch
from oracle import restate as R
from test_gpu_steps import _cfg
import trainer as TR
from ctagan.graphs import GraphedTrainer

curves = torch.load(os.path.join(ROOT, "tests", "golden", "curves_v1.pt"), weights_only=True)
kind, size = sys.argv[1], int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
ref = curves["reg_256" if kind == "reg" else "cyc_128"]


def rm(x):
    x = torch.as_tensor(x, dtype=torch.float64)
    return x.cumsum(0) / torch.arange(1, len(x) + 1, dtype=torch.float64)


for graph, prec in ((True, "bf16"), (False, "bf16"), (False, "fp32")):
    random.seed(42); torch.manual_seed(42)
    tr = (TR.Reg_Trainer if kind == "reg" else TR.Cyc_Trainer)(_cfg("x", size, precision=prec))
    runner = GraphedTrainer(tr, warmup=1, replay_first=False, enabled=graph)
    rows = []
    for i in range(steps):
        a, b = R.synthetic_pair(1, size, seed=1000 + i, phantom=True)
        rows.append({k: v.detach().clone() for k, v in runner.step_host({"A": a, "B": b}).items()})
    torch.cuda.synchronize()
    for k in rows[0]:
        if k not in ref:
            continue
        mine = [float(r[k]) for r in rows]
        a_, b_ = rm(mine), rm(ref[k][:steps])
        pts = [0, 1, 2, 4, 9, 19, 49, 99, steps - 1]
        print(f"graph={graph} {prec} {k}: " + " ".join(f"{p + 1}:{mine[p]:.4f}/{float(ref[k][p]):.4f}({float((a_[p] - b_[p]) / b_[p]):+.1%})" for p in pts if p < steps), flush=True)
