"""200-step reference loss curves for the curve-tracking test -- TEST INFRASTRUCTURE ONLY.

Run in the build container:   python oracle/make_curves.py        (about 15 minutes of CPU)
Writes tests/golden/curves_v1.pt: per-step losses of the restated reference iterations (oracle/restate.py, which
oracle/make_golden.py pins bit for bit against the reference's own modules and source lines), fp32 CPU, seed 42 init, on the
deterministic synthetic phantom stream `restate.synthetic_pair(batch, size, seed=1000 + step, phantom=True)`:

  reg_256:  Reg_Trainer iteration (RegTrainer.py:170-198), batch 1, 256x256, lr 1e-4       -- the configuration of SURVEY.md App. C.4
  cyc_128:  Cyc_Trainer iteration (CycTrainer.py:138-197), batch 1, 128x128, lr 1e-4

The GPU test replays the same stream through the product trainers and compares the cumulative running means that the reference's
Logger prints (trainer/utils.py:81) at every step >= 20 (BASELINE.json: within 2 %).
"""
from __future__ import annotations

import os
import random
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import restate as R  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "curves_v1.pt")
STEPS = 200


def seed(s=42):
    random.seed(s)
    torch.manual_seed(s)


def main():
    torch.set_num_threads(int(os.environ.get("CURVE_THREADS", "8")))
    out = {}
    t0 = time.time()
    seed(); st = R.RegState()
    rows = []
    for i in range(STEPS):
        a, b = R.synthetic_pair(1, 256, seed=1000 + i, phantom=True)
        rows.append(R.reg_step(st, a, b))
        if i % 20 == 0:
            print("reg", i, rows[-1], f"{time.time() - t0:.0f}s", flush=True)
    out["reg_256"] = {k: torch.tensor([r[k] for r in rows], dtype=torch.float64) for k in rows[0]}
    seed(); st = R.CycState()
    rows = []
    for i in range(STEPS):
        a, b = R.synthetic_pair(1, 128, seed=1000 + i, phantom=True)
        rows.append(R.cyc_step(st, a, b))
        if i % 20 == 0:
            print("cyc", i, rows[-1], f"{time.time() - t0:.0f}s", flush=True)
    out["cyc_128"] = {k: torch.tensor([r[k] for r in rows], dtype=torch.float64) for k in rows[0]}
    out["meta"] = {"torch": str(torch.__version__), "steps": STEPS, "stream": "synthetic_pair(1, size, seed=1000+step, phantom=True)",
                   "threads": torch.get_num_threads()}
    torch.save(out, OUT)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
