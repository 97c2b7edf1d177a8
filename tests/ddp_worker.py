"""Worker of tests/test_gpu_ddp.py: one rank of a data-parallel Cyc / Reg iteration; rank 0 saves every parameter gradient."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "cta-gan_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402


def run(kind, size, per_rank, total, out_path, steps=1, precision="fp32"):
    from oracle import restate as R
    from test_gpu_steps import _cfg
    import trainer as TR
    rank = int(os.environ.get("RANK", "0"))
    random.seed(42); torch.manual_seed(42)
    cls = TR.Cyc_Trainer if kind == "cyc" else TR.Reg_Trainer
    tr = cls(_cfg("x", size, batch=per_rank, precision=precision))
    a, b = R.synthetic_pair(total, size, seed=77, phantom=True)
    sl = slice(rank * per_rank, (rank + 1) * per_rank)
    for _ in range(steps):
        losses = tr.step({"A": a[sl].contiguous(), "B": b[sl].contiguous()})
    torch.cuda.synchronize()
    if rank == 0:
        nets = [m for m in tr.__dict__.values() if isinstance(m, torch.nn.Module) and len(list(m.parameters()))]
        grads = [None if p.grad is None else p.grad.detach().cpu().clone() for m in nets for p in m.parameters()]
        weights = [p.detach().cpu().clone() for m in nets for p in m.parameters()]
        syncs = [v for v in tr.__dict__.values() if hasattr(v, "n_early")]
        early = sum(s_.n_early for s_ in syncs)
        torch.save({"grads": grads, "weights": weights, "early_chunks": early, "losses": {k: float(v) for k, v in losses.items()}}, out_path)


if __name__ == "__main__":
    kind, size, per_rank, total, out_path = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    steps = int(sys.argv[6]) if len(sys.argv) > 6 else 1
    precision = sys.argv[7] if len(sys.argv) > 7 else "fp32"
    run(kind, size, per_rank, total, out_path, steps, precision)
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()
    sys.stdout.flush()
    os._exit(0)
