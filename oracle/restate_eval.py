"""CPU restatement (numpy, float64 like the reference) of the data-side and evaluation-side arithmetic around the hot path --
TEST INFRASTRUCTURE ONLY (same rules as oracle/restate.py).  Each function cites the reference lines it follows.

Pinning (oracle/make_golden.py:main_v3 -> tests/golden/golden_v3.pt): to_windowdata, MAE, PSNR, UQI and the mask block are checked against
the reference's own source lines exec'd from /root/reference; affine_nearest against PIL itself (Image.transform, the library call the
reference's RandomAffine makes); resize against torch's F.interpolate.  SSIM is `skimage.measure.compare_ssim` (scikit-image <= 0.17, not
pinned by the reference and absent from this image): its published algorithm (Wang et al. 2004; 7x7 uniform window, sample covariance,
K1 = 0.01, K2 = 0.03, data range 2 for float images) is restated here with scipy.ndimage.uniform_filter -- "parity unpinned" for SSIM."""
from __future__ import annotations

import numpy as np


def read_dicom_norm(pixel_array: np.ndarray) -> np.ndarray:
    """trainer/datasets.py:74-82."""
    image2 = pixel_array.astype(np.int64).copy()
    image2[image2 < 0] = 0
    image2 = image2 / 4095
    return (image2 - 0.5) / 0.5


def window_image(hu: np.ndarray, center: float, width: float) -> np.ndarray:
    """The display window shared by read_ori_w (trainer/datasets.py:45-56) and to_windowdata (trainer/CycTrainer.py:41-57)."""
    win_min = (2 * center - width) / 2.0 + 0.5
    win_max = (2 * center + width) / 2.0 + 0.5
    dFactor = 255.0 / (win_max - win_min)
    image = hu - win_min
    image = np.trunc(image * dFactor)
    image[image > 255] = 255
    image[image < 0] = 0
    image = image / 255
    return (image - 0.5) / 0.5


def to_windowdata(image: np.ndarray, WC: float, WW: float) -> np.ndarray:
    """trainer/CycTrainer.py:34-57."""
    image = (image + 1) * 0.5 * 4095
    image[image == 0] = -2000
    image = image - 1024
    return window_image(image, WC, WW)


def eval_images(fake_B: np.ndarray, real_B: np.ndarray, WC: float, WW: float):
    """The four images test() compares (trainer/CycTrainer.py:286-318), aliasing included: returns (c, b, fake_masked, real_masked)."""
    b = to_windowdata(real_B, WC, WW)
    bb = b
    bb[bb < 0.3] = 0
    bb[bb >= 0.3] = 1
    b = b * bb
    b[b == 0] = -1
    c = to_windowdata(fake_B, WC, WW) * bb
    cc = c
    cc[cc < 0.3] = 0
    cc[cc >= 0.3] = 1
    c = c * cc
    c[c == 0] = -1
    real_m = real_B * bb
    real_m[real_m == 0] = -1
    fake_m = fake_B * cc
    fake_m[fake_m == 0] = -1
    return c, b, fake_m, real_m


def PSNR(fake, real):
    """trainer/CycTrainer.py:362-375."""
    a = np.where(real != -1)
    x, y = a[0], a[1]
    if x.size == 0 or y.size == 0:
        mse = np.mean(((fake + 1) / 2. - (real + 1) / 2.) ** 2) + 1e-10
    else:
        mse = np.mean(((fake[x, y] + 1) / 2. - (real[x, y] + 1) / 2.) ** 2)
    if mse < 1.0e-10:
        return 100
    return 20 * np.log10(1 / (np.sqrt(mse) + 1e-10))


def MAE(fake, real):
    """trainer/CycTrainer.py:377-388."""
    a = np.where(real != -1)
    x, y = a[0], a[1]
    if x.size == 0 or y.size == 0:
        mae = np.nanmean(np.abs(fake - real)) + 1e-10
    else:
        mae = np.nanmean(np.abs(fake[x, y] - real[x, y]))
    return mae / 2


def UQI(fake, real):
    """trainer/CycTrainer.py:390-398."""
    meanf, meanr = np.mean(fake), np.mean(real)
    m, n = np.shape(fake)
    varf = np.sqrt(np.sum((fake - meanf) ** 2) / (m * n - 1))
    varr = np.sqrt(np.sum((real - meanr) ** 2) / (m * n - 1))
    cov = np.sum((fake - meanf) * (real - meanr)) / (m * n - 1)
    return 4 * meanf * meanr * cov / ((meanf ** 2 + meanr ** 2) * (varf ** 2 + varr ** 2) + 1e-10)


def SSIM(X, Y, win_size=7, data_range=2.0):
    """skimage.measure.compare_ssim defaults (see the module docstring)."""
    from scipy.ndimage import uniform_filter
    X, Y = X.astype(np.float64), Y.astype(np.float64)
    NP = win_size ** 2
    cov_norm = NP / (NP - 1)
    ux, uy = uniform_filter(X, size=win_size), uniform_filter(Y, size=win_size)
    uxx, uyy, uxy = uniform_filter(X * X, size=win_size), uniform_filter(Y * Y, size=win_size), uniform_filter(X * Y, size=win_size)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    pad = (win_size - 1) // 2
    return S[pad:-pad, pad:-pad].mean()


def slice_metrics(fake_B: np.ndarray, real_B: np.ndarray, WC=40.0, WW=400.0):
    """(MAEw, PSNRw, SSIMw, UQIw, MAE, PSNR, SSIM, UQI) of one slice, trainer/CycTrainer.py:300-330."""
    c, b, fm, rm = eval_images(fake_B.copy(), real_B.copy(), WC, WW)
    return (MAE(c, b), PSNR(c, b), SSIM(c, b), UQI(c, b), MAE(fm, rm), PSNR(fm, rm), SSIM(fm, rm), UQI(fm, rm))


def to_dicom_int16(fake: np.ndarray) -> np.ndarray:
    """trainer/CycTrainer.py:337-341."""
    return ((fake + 1) * 0.5 * 4095).astype(np.int16)


def affine_nearest(img: np.ndarray, m, fill=-1.0) -> np.ndarray:
    """PIL Image.transform(size, AFFINE, m, resample=NEAREST) for a float image (libImaging/Geometry.c, affine_fixed): what
    torchvision's RandomAffine does to the 'F'-mode slice (trainer/CycTrainer.py:91-95)."""
    H, W = img.shape
    FIX = lambda v: int(np.floor(v * 65536.0 + 0.5))
    a0, a1, a3, a4 = FIX(m[0]), FIX(m[1]), FIX(m[3]), FIX(m[4])
    a2, a5 = FIX(m[2] + m[0] * 0.5 + m[1] * 0.5), FIX(m[5] + m[3] * 0.5 + m[4] * 0.5)
    ys, xs = np.mgrid[0:H, 0:W]
    xin = (a2 + ys * a1 + xs * a0) >> 16
    yin = (a5 + ys * a4 + xs * a3) >> 16
    ok = (xin >= 0) & (xin < W) & (yin >= 0) & (yin < H)
    out = np.full((H, W), fill, dtype=img.dtype)
    out[ok] = img[yin[ok], xin[ok]]
    return out
