"""Multi-rank equivalence on real GPUs: N ranks x b slices == 1 rank x N*b slices.  Every op on the path is per sample
(InstanceNorm statistics are per (n, c)) and the losses are batch means, so the all-reduced (averaged) gradients of two ranks with one
slice each must equal the gradients of one rank holding both slices, up to summation order.  Needs 2 GPUs (gpurun --gpus 2)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.parametrize("kind,size", [("cyc", 64), ("reg", 256)])
def test_two_ranks_equal_one_rank_with_twice_the_batch(kind, size, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    worker = os.path.join(ROOT, "tests", "ddp_worker.py")
    two, one = str(tmp_path / "two.pt"), str(tmp_path / "one.pt")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                    "--master-port", str(_free_port()), worker, kind, str(size), "1", "2", two], check=True, env=env, timeout=600)
    env1 = {k: v for k, v in env.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    subprocess.run([sys.executable, worker, kind, str(size), "2", "2", one], check=True, env=env1, timeout=600)
    a, b = torch.load(two), torch.load(one)
    ref_max = max(float(g.abs().max()) for g in b["grads"] if g is not None)
    num = den = 0.0
    for k, (g2, g1) in enumerate(zip(a["grads"], b["grads"])):
        assert (g2 is None) == (g1 is None), k
        if g1 is None:
            continue
        num += float((g2.double() - g1.double()).pow(2).sum()); den += float(g1.double().pow(2).sum())
        assert float((g2 - g1).abs().max()) <= 2e-4 * ref_max, (k, float((g2 - g1).abs().max()), ref_max)
    # fp32 validation mode: the two runs differ only in summation order (batch-of-2 reductions vs the all-reduce); kinks of ReLU / |.|
    # limit the agreement of whole-network gradients to ~1e-4, far below any scaling or routing error (O(1))
    assert (num / den) ** 0.5 <= 1e-4, (num / den) ** 0.5
    for k, v in b["losses"].items():
        pass        # (the per-rank losses are over different slices: only the gradients are comparable)


@pytest.mark.parametrize("kind,size", [("cyc", 64), ("reg", 256)])
def test_overlapped_all_reduce_equals_all_reduce_after_backward(kind, size, tmp_path):
    """The bucket chunks that go to the all-reduce while the backward pass is still running (GradSync._final, from the second step on)
    must give exactly what one all-reduce after the whole backward gives: three bf16 training steps on two ranks, every weight of
    every network bit for bit (the kernels are deterministic and an average of two ranks does not depend on the chunking)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    worker = os.path.join(ROOT, "tests", "ddp_worker.py")
    outs = {}
    for early in ("1", "0"):
        out = str(tmp_path / f"early{early}.pt")
        env = dict(os.environ, MASTER_ADDR="127.0.0.1", CTAGAN_DDP_EARLY=early, CTAGAN_DDP_CHUNK_MB="2")
        subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(_free_port()), worker, kind, str(size), "1", "2", out, "3", "bf16"], check=True, env=env, timeout=600)
        outs[early] = torch.load(out)
    assert outs["1"]["early_chunks"] > 0 and len(outs["1"]["weights"]) == len(outs["0"]["weights"])
    for k, (a, b) in enumerate(zip(outs["1"]["weights"], outs["0"]["weights"])):
        assert torch.equal(a, b), (k, float((a - b).abs().max()))
