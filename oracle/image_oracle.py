"""TEST INFRASTRUCTURE — not part of the product path.

CPU restatement of the reference's per-sample image path, PreprocessRGB.__call__ with image_aspect_ratio='pad'
(dexbotic/data/dataset/rgb_preprocess.py:13-44): expand2square -> HF CLIPImageProcessor.preprocess (PIL bicubic resize,
rescale 1/255, normalize).  The resize lives in a third-party dependency (Pillow, src/libImaging/Resample.c: 8-bit
separable convolution in 22-bit fixed point, horizontal pass first, uint8 between the passes); its published algorithm is
restated here in numpy integer arithmetic and PINNED against Pillow itself (tests/test_input_pipeline_oracle.py runs
both wherever Pillow is importable, and oracle/make_golden.py wrote tests/golden/image_preprocess.npz from Pillow + the
HF processor).  ActionNorm._normalize: data/dataset/transform/action.py:268-275."""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bicubic_filter(x: float) -> float:
    """Resample.c bicubic_filter, a = -0.5."""
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs (box = the whole axis) + normalize_coeffs_8bpc."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ww = 0.0
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = np.zeros(ksize, dtype=np.float64)
        for x in range(xmax):
            w = bicubic_filter((x + xmin - center + 0.5) * ss)
            k[x] = w
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
        for x in range(ksize):
            kk[xx, x] = int(-0.5 + k[x] * (1 << PRECISION_BITS)) if k[x] < 0 else int(0.5 + k[x] * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return kk, bounds


def _resample_axis0(img: np.ndarray, out_size: int) -> np.ndarray:
    """One pass of ImagingResampleVertical_8bpc / Horizontal_8bpc along axis 0 of an [n, m, 3] uint8 array."""
    kk, bounds = precompute_coeffs(img.shape[0], out_size)
    out = np.zeros((out_size,) + img.shape[1:], dtype=np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        xmin, xmax = bounds[xx]
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        acc = acc + (src[xmin:xmin + xmax] * kk[xx, :xmax, None, None]).sum(axis=0)
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def pil_bicubic_resize(img_hwc: np.ndarray, out_size: int) -> np.ndarray:
    """Image.resize((out, out), BICUBIC) of an RGB uint8 image: horizontal pass, then vertical (ImagingResample)."""
    h = img_hwc
    if img_hwc.shape[1] != out_size:
        h = _resample_axis0(img_hwc.transpose(1, 0, 2), out_size).transpose(1, 0, 2)
    if h.shape[0] != out_size:
        h = _resample_axis0(h, out_size)
    return h


def expand2square(img_hwc: np.ndarray, background) -> np.ndarray:
    """rgb_preprocess.py:32-44."""
    H, W, _ = img_hwc.shape
    if W == H:
        return img_hwc
    L = max(H, W)
    out = np.empty((L, L, 3), dtype=np.uint8)
    out[:] = np.asarray(background, dtype=np.uint8)
    if W > H:
        out[(W - H) // 2:(W - H) // 2 + H] = img_hwc
    else:
        out[:, (H - W) // 2:(H - W) // 2 + W] = img_hwc
    return out


def preprocess_rgb(img_hwc: np.ndarray, size: int, image_mean, image_std, rescale_factor: float = 1 / 255,
                   image_pad_mode: str = "mean"):
    """Returns (pixel_values float32 [3, size, size], resized uint8 [size, size, 3])."""
    bg = (0, 0, 0) if image_pad_mode == "zero" else tuple(int(x * 255) for x in image_mean)
    sq = expand2square(img_hwc, bg)
    u8 = pil_bicubic_resize(sq, size)
    x = (u8.astype(np.float64) * rescale_factor).astype(np.float32)                      # image_transforms.rescale
    x = (x - np.asarray(image_mean, dtype=np.float32)) / np.asarray(image_std, dtype=np.float32)   # normalize
    return x.transpose(2, 0, 1).astype(np.float32), u8


def action_normalize(data: np.ndarray, stats: dict, use_quantiles: bool) -> np.ndarray:
    """ActionNorm._normalize (action.py:268-275), verbatim arithmetic."""
    data = np.asarray(data)
    if use_quantiles:
        return ((data - np.asarray(stats["min"])) / (np.asarray(stats["max"]) - np.asarray(stats["min"]) + 1e-6) * 2.0
                - 1.0).astype(np.float32)
    return ((data - np.asarray(stats["mean"])) / (np.asarray(stats["std"]) + 1e-6)).astype(np.float32)
