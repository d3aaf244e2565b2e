"""Device-side input pipeline: the per-sample CPU work of the reference's dataloader workers, batched on
the GPU.  `ImagePreprocessor` = PreprocessRGB with image_aspect_ratio='pad' (dexbotic/data/dataset/rgb_preprocess.py:
13-44) + the HF CLIP / SigLIP image processor's resize (PIL bicubic) / rescale / normalize; `ActionNormalizer` =
ActionNorm (dexbotic/data/dataset/transform/action.py:229-275).  The random pixel augmentations of
data/dataset/augmentations.py are albumentations policies (un-vendored, absent offline): not restated.

The host side only builds small tables (Pillow's fixed-point resize coefficients, the processor's 3 x 256 value map);
every pixel is touched on the device (csrc/input_ops.cu)."""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from .ops import _cuda, _stream

PRECISION_BITS = 32 - 8 - 2          # Pillow src/libImaging/Resample.c


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_coefficients(in_size: int, out_size: int):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for the bicubic filter (support 2) over the whole input
    range: (kk int32 [out, ksize], bounds int32 [out, 2] = (first tap, tap count), ksize)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(w)
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return kk, bounds, ksize


class ImagePreprocessor:
    """uint8 camera frames [B, H, W, 3] on the device -> normalised tensor [B, 3, size, size]."""

    def __init__(self, size: int = 224, image_mean: Sequence[float] = (0.48145466, 0.4578275, 0.40821073),
                 image_std: Sequence[float] = (0.26862954, 0.26130258, 0.27577711), rescale_factor: float = 1 / 255,
                 image_pad_mode: str = "mean", device="cuda"):
        self.size, self.device = size, torch.device(device)
        self.mean, self.std, self.rescale = tuple(image_mean), tuple(image_std), rescale_factor
        # expand2square's background: int(x * 255) of the processor mean, or black (rgb_preprocess.py:22-25)
        self.background = (0, 0, 0) if image_pad_mode == "zero" else tuple(int(x * 255) for x in image_mean)
        # the processor's float stage, value by value, in its own arithmetic: rescale = float64 product cast to
        # float32 (image_transforms.rescale), normalize = (image - mean) / std in float32
        v = (np.arange(256, dtype=np.float64) * rescale_factor).astype(np.float32)
        lut = np.stack([(v - np.float32(m)) / np.float32(s) for m, s in zip(self.mean, self.std)]).astype(np.float32)
        self.lut = torch.from_numpy(lut).to(self.device).contiguous()
        self._tables = {}

    def _table(self, side: int):
        if side not in self._tables:
            kk, bounds, ksize = resample_coefficients(side, self.size)
            self._tables[side] = (torch.from_numpy(kk).to(self.device).contiguous(),
                                  torch.from_numpy(bounds).to(self.device).contiguous(), ksize)
        return self._tables[side]

    def __call__(self, frames_u8: torch.Tensor, dtype: torch.dtype = torch.float32, return_uint8: bool = False):
        _cuda(frames_u8)
        assert frames_u8.dtype == torch.uint8 and frames_u8.dim() == 4 and frames_u8.shape[-1] == 3 and frames_u8.is_contiguous()
        B, H, W, _ = frames_u8.shape
        L = max(H, W)
        kk, bounds, ksize = self._table(L)          # the padded image is square: one table serves both passes
        tmp = torch.empty((B, L, self.size, 3), device=frames_u8.device, dtype=torch.uint8)
        out = torch.empty((B, 3, self.size, self.size), device=frames_u8.device, dtype=dtype)
        u8 = torch.empty((B, self.size, self.size, 3), device=frames_u8.device, dtype=torch.uint8) if return_uint8 else None
        if L == self.size:
            raise NotImplementedError("frames already at the target size skip Pillow's resize passes: not wired")
        rc = _lib.load().b200_image_preprocess(
            frames_u8.data_ptr(), B, H, W, self.size, kk.data_ptr(), bounds.data_ptr(), ksize, kk.data_ptr(),
            bounds.data_ptr(), ksize, self.background[0], self.background[1], self.background[2], self.lut.data_ptr(),
            tmp.data_ptr(), out.data_ptr(), None if u8 is None else u8.data_ptr(),
            _lib.F32 if dtype == torch.float32 else _lib.BF16, _stream())
        _lib.check(rc, "image_preprocess")
        return (out, u8) if return_uint8 else out


class ActionNormalizer:
    """ActionNorm._normalize (action.py:268-275) on device: float64 arithmetic, fp32 result."""

    def __init__(self, stats: dict, use_quantiles: bool = False, device="cuda"):
        keys = ("min", "max") if use_quantiles else ("mean", "std")
        self.quantile = use_quantiles
        self.a = torch.tensor(np.asarray(stats[keys[0]], dtype=np.float64), device=device).contiguous()
        self.b = torch.tensor(np.asarray(stats[keys[1]], dtype=np.float64), device=device).contiguous()

    def __call__(self, actions: torch.Tensor) -> torch.Tensor:
        _cuda(actions)
        x = actions.to(torch.float64).contiguous()
        D = x.shape[-1]
        assert D == self.a.numel()
        out = torch.empty(x.shape, device=x.device, dtype=torch.float32)
        _lib.check(_lib.load().b200_action_normalize(x.data_ptr(), self.a.data_ptr(), self.b.data_ptr(), out.data_ptr(),
                                                     x.numel() // D, D, int(self.quantile), _stream()),
                   "action_normalize")
        return out
