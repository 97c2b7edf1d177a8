import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cta-gan_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def golden():
    """Vectors frozen from the real reference (oracle/make_golden.py): plain tensors / numbers / strings only (weights_only load)."""
    import torch
    g = torch.load(os.path.join(ROOT, "tests", "golden", "golden_v1.pt"), weights_only=True)
    g.update(torch.load(os.path.join(ROOT, "tests", "golden", "golden_v2.pt"), weights_only=True))
    g.update(torch.load(os.path.join(ROOT, "tests", "golden", "golden_v3.pt"), weights_only=True))
    return g


_ORDER = ["test_abi_and_host", "test_oracle", "test_ddp", "test_gpu_ops", "test_gpu_tc", "test_gpu_thin_tc", "test_gpu_data_eval", "test_gpu_optim", "test_gpu_determinism", "test_gpu_modules",
          "test_gpu_bf16", "test_gpu_steps", "test_gpu_curves", "test_gpu_ddp"]


def pytest_collection_modifyitems(config, items):
    """Kernel-level tests first (ops, tcgen05 engine, determinism), whole modules next, training iterations last: under `-x` a
    failure high in the stack can then never hide the state of the layers below it."""
    def rank(item):
        name = item.fspath.basename
        for k, prefix in enumerate(_ORDER):
            if name.startswith(prefix):
                return k
        return len(_ORDER)
    items.sort(key=rank)          # stable: keeps the order inside a file
