import sys, os, random
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import _ctagan_path
import torch
from test_gpu_steps import _run_cyc
runs = {m: _run_cyc(m.split("#")[0], 7) for m in ("two_phase#1", "two_phase#2", "fused", "graph")}
for i in range(7):
    for m, r in runs.items():
        print(i, m, None if r[i] is None else {k: round(v, 5) for k, v in r[i].items()})
