"""On-GPU evaluation for test(): the metrics the reference computes per slice on the CPU after a `.cpu().numpy()`
round trip (trainer/CycTrainer.py:286-330: display window, 0.3-threshold masks, MAE / PSNR / SSIM / UQI on the windowed and on the
masked raw pair; :337-341 int16 DICOM pixels) as fused, deterministic reduction kernels over whole batches.  Per-slice results stay on
the device; `result()` makes the single host read.  LPIPS (a third-party AlexNet) is not computed."""
from __future__ import annotations

import ctypes

import torch

from . import lib as L
from . import ops

NAMES = ("MAEw", "PSNRw", "SSIMw", "UQIw", "MAE", "PSNR", "SSIM", "UQI")


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def slice_metrics(fake: torch.Tensor, real: torch.Tensor, wc: float = 40.0, ww: float = 400.0) -> torch.Tensor:
    """fake, real: [B, 1, H, W] (or [B, H, W]) fp32 in [-1, 1] -> [B, 8] fp64 on the device, columns = NAMES."""
    ops.ensure_device()
    fake, real = fake.detach().float().contiguous(), real.detach().float().contiguous()
    H, W = fake.shape[-2:]
    B = fake.numel() // (H * W)
    lib = L.load()
    out = torch.empty((B, 8), dtype=torch.float64, device=fake.device)
    scratch = torch.empty((int(lib.ctagan_eval_metrics_scratch_doubles(B, H, W)),), dtype=torch.float64, device=fake.device)
    ops._count(3)
    L.check(lib.ctagan_eval_metrics(_p(fake), _p(real), _p(out), _p(scratch), B, H, W, float(wc), float(ww),
                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out


def to_dicom_int16(x: torch.Tensor) -> torch.Tensor:
    """(x + 1) * 0.5 * 4095 truncated to int16: the pixel array the reference writes back into the DICOM file."""
    ops.ensure_device()
    x = x.detach().float().contiguous()
    out = torch.empty(x.shape, dtype=torch.int16, device=x.device)
    ops._count(1)
    L.check(L.load().ctagan_to_dicom_i16(_p(x), _p(out), x.numel(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out


class Evaluator:
    """Accumulates the per-slice metrics of a test() run on the device."""

    def __init__(self, device, wc: float = 40.0, ww: float = 400.0):
        self.sum = torch.zeros((8,), dtype=torch.float64, device=device)
        self.n = 0
        self.wc, self.ww = wc, ww

    def add(self, fake, real):
        m = slice_metrics(fake, real, self.wc, self.ww)
        self.sum += m.sum(0)
        self.n += m.shape[0]
        return m

    def result(self):
        vals = (self.sum / max(self.n, 1)).cpu().tolist()          # the one host read of the run
        out = dict(zip(NAMES, vals))
        out["slices"] = self.n
        return out
