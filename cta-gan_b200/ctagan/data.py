"""GPU input pipeline: what the reference does per slice on the CPU inside a one-worker DataLoader -- DICOM pixel
array -> numpy -> [-1, 1] normalisation / display window (trainer/datasets.py:36-82) -> PIL RandomAffine (trainer/CycTrainer.py:91-95)
-> nearest Resize (trainer/utils.py:13-30) -> pinned H2D -- done here as: a background reader thread that fills pinned int16 staging
buffers (double-buffered), an asynchronous H2D copy of the RAW 16-bit slices on a copy stream (a quarter of the fp32 bytes), and
batched sm_100a kernels for normalisation, window, affine augmentation and resize.  The training stream only waits on an event.

Slice lists are the reference's text files (one path per line; the target slice is the same path with "SE0" -> "SE1",
trainer/datasets.py:92-99).  Files: `.npy` holding the stored 16-bit pixel array of one slice, or DICOM when `pydicom` is importable
(it is not part of this image; the error says so)."""
from __future__ import annotations

import ctypes
import math
import os
import queue
import threading
from typing import Dict, List, Sequence

import numpy as np
import torch

from . import lib as L
from . import ops


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def hu_to_unit(raw: torch.Tensor, add: int = 0) -> torch.Tensor:
    """int16 stored pixel values [..] -> fp32 in [-1, 1] (read_dicom, trainer/datasets.py:74-82)."""
    ops.ensure_device()
    out = torch.empty(raw.shape, dtype=torch.float32, device=raw.device)
    ops._count(1)
    L.check(L.load().ctagan_hu_to_unit(_p(raw), _p(out), raw.numel(), int(add), _st()))
    return out


def hu_window(hu: torch.Tensor, center: float = 50.0, width: float = 400.0) -> torch.Tensor:
    """int16 HU -> display-windowed fp32 in [-1, 1] (read_ori_w, trainer/datasets.py:45-56)."""
    ops.ensure_device()
    out = torch.empty(hu.shape, dtype=torch.float32, device=hu.device)
    ops._count(1)
    L.check(L.load().ctagan_hu_window(_p(hu), _p(out), hu.numel(), float(center), float(width), _st()))
    return out


def resize_nearest(x: torch.Tensor, size: Sequence[int]) -> torch.Tensor:
    """[B, (1,) H, W] fp32 -> [B, (1,) size] with F.interpolate's default nearest rule (trainer/utils.py:28)."""
    shape = x.shape
    B, Hs, Ws = int(np.prod(shape[:-2])), shape[-2], shape[-1]
    if (Hs, Ws) == tuple(size):
        return x
    out = torch.empty((*shape[:-2], size[0], size[1]), dtype=torch.float32, device=x.device)
    ops._count(1)
    L.check(L.load().ctagan_resize_nearest(_p(x.contiguous()), _p(out), B, Hs, Ws, size[0], size[1], _st()))
    return out


def random_affine_matrices(n: int, size: Sequence[int], level: float, generator=None) -> torch.Tensor:
    """n inverse affine matrices [n, 6] (fp64, CPU) drawn the way RandomAffine(degrees=level, translate=[0.02*level]*2,
    scale=[1-0.02*level, 1+0.02*level]) of trainer/CycTrainer.py:92 draws them (torchvision's get_params, torch RNG) and inverted the
    way torchvision hands them to PIL (_get_inverse_affine_matrix about the image centre)."""
    H, W = size
    out = torch.empty((n, 6), dtype=torch.float64)
    for k in range(n):
        angle = float(torch.empty(1).uniform_(-float(level), float(level), generator=generator).item())
        max_dx, max_dy = float(0.02 * level * W), float(0.02 * level * H)
        tx = int(round(torch.empty(1).uniform_(-max_dx, max_dx, generator=generator).item()))
        ty = int(round(torch.empty(1).uniform_(-max_dy, max_dy, generator=generator).item()))
        scale = float(torch.empty(1).uniform_(1 - 0.02 * level, 1 + 0.02 * level, generator=generator).item())
        out[k] = torch.tensor(inverse_affine_matrix((W * 0.5, H * 0.5), angle, (tx, ty), scale), dtype=torch.float64)
    return out


def inverse_affine_matrix(center, angle, translate, scale, shear=(0.0, 0.0)):
    """torchvision.transforms.functional._get_inverse_affine_matrix (the matrix PIL's Image.transform receives)."""
    rot = math.radians(angle)
    sx, sy = math.radians(shear[0]), math.radians(shear[1])
    cx, cy = center
    tx, ty = translate
    a = math.cos(rot - sy) / math.cos(sy)
    b = -math.cos(rot - sy) * math.tan(sx) / math.cos(sy) - math.sin(rot)
    c = math.sin(rot - sy) / math.cos(sy)
    d = -math.sin(rot - sy) * math.tan(sx) / math.cos(sy) + math.cos(rot)
    m = [d, -b, 0.0, -c, a, 0.0]
    m = [x / scale for x in m]
    m[2] += m[0] * (-cx - tx) + m[1] * (-cy - ty)
    m[5] += m[3] * (-cx - tx) + m[4] * (-cy - ty)
    m[2] += cx
    m[5] += cy
    return m


def affine_nearest(x: torch.Tensor, inv_matrices: torch.Tensor, fill: float = -1.0) -> torch.Tensor:
    """x [B, (1,) H, W] fp32, inv_matrices [B, 6] fp64: PIL's nearest-neighbour affine resampling, pixels from outside = fill."""
    shape = x.shape
    B, H, W = int(np.prod(shape[:-2])), shape[-2], shape[-1]
    m = inv_matrices.to(device=x.device, dtype=torch.float64, non_blocking=True).contiguous()
    assert m.shape == (B, 6)
    out = torch.empty_like(x)
    ops._count(1)
    L.check(L.load().ctagan_affine_nearest(_p(x.contiguous()), _p(out), B, H, W, _p(m), float(fill), _st()))
    return out


def read_slice(path: str) -> np.ndarray:
    """The stored 16-bit pixel array of one slice."""
    if path.endswith(".npy"):
        return np.load(path)
    try:
        import pydicom
    except ImportError as exc:  # pragma: no cover - pydicom is not in this image
        raise RuntimeError(f"{path}: reading DICOM needs pydicom, which is not installed here; convert the slices to .npy "
                           "(the stored 16-bit pixel array) or install pydicom") from exc
    return pydicom.dcmread(path.replace("../../../", "../../"), force=True).pixel_array      # trainer/datasets.py:75


class SliceListLoader:
    """Iterable of batches {key: device tensor [B, 1, size, size]} from a slice list, prepared by the GPU pipeline.

    keys ("A", "B") -> read_dicom normalisation of the SE0 / SE1 slices (ImageDataset, trainer/datasets.py:86-119);
    keys ("A2", "B1", "B2") -> the raw-normalised input slice, the windowed and the raw-normalised target slice (ImageDataset_x, :190-232).
    augment: RandomAffine with fill -1 on every image (the same matrix for the images of one slice pair is NOT shared: the reference
    draws per transform call).  Batches are sharded over ranks by slice index; a short last batch is dropped."""

    def __init__(self, list_path: str, batch: int, size: int, keys, device, rank=0, world=1, augment=False, noise_level=1, seed=42,
                 depth=2):
        with open(list_path) as f:
            files = [ln.strip("\n") for ln in f.readlines() if ln.strip()]
        self.files_a = sorted(files)[rank::world]
        self.batch, self.size, self.keys, self.device = batch, size, tuple(keys), device
        self.augment, self.level = bool(augment) and noise_level > 0, noise_level
        self.gen = torch.Generator().manual_seed(seed)
        self.depth = depth
        self.copy_stream = ops.named_stream("data.copy")

    def __len__(self):
        return len(self.files_a) // self.batch

    def _reader(self, q: "queue.Queue"):
        try:
            staging = None
            for bi in range(len(self)):
                paths = self.files_a[bi * self.batch:(bi + 1) * self.batch]
                a = [read_slice(p_) for p_ in paths]
                b = [read_slice(p_.replace("SE0", "SE1")) for p_ in paths]
                H, W = a[0].shape
                if staging is None or staging[0].shape[-2:] != (H, W):
                    staging = [torch.empty((2, self.batch, H, W), dtype=torch.int16).pin_memory() for _ in range(self.depth + 1)]
                buf = staging[bi % len(staging)]
                for k in range(self.batch):
                    buf[0, k].copy_(torch.from_numpy(np.ascontiguousarray(a[k]).astype(np.int16, copy=False)))
                    buf[1, k].copy_(torch.from_numpy(np.ascontiguousarray(b[k]).astype(np.int16, copy=False)))
                q.put(buf)
            q.put(None)
        except BaseException as exc:  # noqa: BLE001 - surfaced in the consumer
            q.put(exc)

    def _prepare(self, buf: torch.Tensor) -> Dict[str, torch.Tensor]:
        """pinned int16 [2, B, H, W] -> device batch (runs on the copy stream)."""
        raw = buf.to(self.device, non_blocking=True)
        H, W = raw.shape[-2:]
        out = {}
        for key in self.keys:
            src = raw[0] if key.startswith("A") else raw[1]
            if key == "B1":          # windowed target (read_ori_w image1; the stored values are HU + 1024, :41)
                img = hu_window(src - 1024, 50.0, 400.0)
            else:
                img = hu_to_unit(src, 0)
            if self.augment:
                img = affine_nearest(img, random_affine_matrices(self.batch, (H, W), self.level, self.gen), -1.0)
            out[key] = resize_nearest(img, (self.size, self.size)).unsqueeze(1)
        return out

    def __iter__(self):
        q: "queue.Queue" = queue.Queue(maxsize=self.depth)
        th = threading.Thread(target=self._reader, args=(q,), daemon=True)
        th.start()
        pending: List = []
        cur = torch.cuda.current_stream()
        while True:
            item = q.get()
            if isinstance(item, BaseException):
                raise item
            if item is not None:
                with torch.cuda.stream(self.copy_stream):
                    batch = self._prepare(item)
                    ev = torch.cuda.Event()
                    ev.record(self.copy_stream)
                pending.append((batch, ev))
            if pending and (item is None or len(pending) > 1):       # one batch of look-ahead: H2D + kernels of batch i+1 under step i
                batch, ev = pending.pop(0)
                cur.wait_event(ev)
                for t in batch.values():
                    t.record_stream(cur)
                yield batch
            if item is None and not pending:
                break
