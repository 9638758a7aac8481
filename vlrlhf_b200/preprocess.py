"""GPU image preprocessing for the DPO collator.

Mirror of what `LlavaDPODataCollatorWithPadding.__call__` (models/Llava/__init__.py:435-443) does after
`Image.open(...).convert("RGB")`: transformers-4.41 `CLIPImageProcessor.preprocess` = Pillow bicubic resize to the
shortest edge, center crop, rescale by 1/255, normalize with the OpenAI CLIP mean/std -> float32 [B, 3, 336, 336].

The host side here is integer/table logic only: the output geometry (image_transforms.get_resize_output_image_size,
center_crop) and Pillow's fixed-point resampling tables (Resample.c precompute_coeffs / normalize_coeffs_8bpc),
vectorised and cached per (input size, output size).  Every pixel is computed by libvlb200
(`vlb200_clip_preprocess_u8`); the result is bit-exact with Pillow + numpy for uint8 input.
"""
from __future__ import annotations

import functools
import math
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import ops

OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
PRECISION_BITS = 22  # Resample.c: 32 - 8 - 2


def _bicubic(x: np.ndarray) -> np.ndarray:
    a = -0.5
    x = np.abs(x)
    near = ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    far = (((x - 5) * x + 8) * x - 4) * a
    return np.where(x < 1.0, near, np.where(x < 2.0, far, 0.0))


@functools.lru_cache(maxsize=256)
def resample_tables(in_size: int, out_size: int) -> Tuple[int, np.ndarray, np.ndarray]:
    """Pillow's 8-bit bicubic tables for resizing one axis from in_size to out_size:
    -> (ksize, bounds int32 [out, 2] = (first tap, tap count), coefficients int32 [out, ksize])."""
    scale = in_size / out_size
    fscale = max(scale, 1.0)
    support = 2.0 * fscale
    ksize = int(math.ceil(support)) * 2 + 1
    centers = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    xmin = np.maximum(np.trunc(centers - support + 0.5).astype(np.int64), 0)
    count = np.minimum(np.trunc(centers + support + 0.5).astype(np.int64), in_size) - xmin
    taps = np.arange(ksize, dtype=np.int64)[None, :]
    live = taps < count[:, None]
    w = _bicubic(((taps + xmin[:, None]) - centers[:, None] + 0.5) * (1.0 / fscale))
    w = np.where(live, w, 0.0)
    total = np.cumsum(w, axis=1)[:, -1:]  # the C loop accumulates left to right
    w = np.where(total != 0.0, w / np.where(total != 0.0, total, 1.0), w)
    fixed = w * float(1 << PRECISION_BITS)
    coef = np.where(w < 0, np.trunc(fixed - 0.5), np.trunc(fixed + 0.5)).astype(np.int32)
    bounds = np.stack([xmin, count], axis=1).astype(np.int32)
    return ksize, bounds, coef


def resize_geometry(h: int, w: int, size: int, crop: int) -> Tuple[int, int, int, int]:
    """-> (new_h, new_w, top, left): shortest edge to `size` (the long edge truncated, as
    get_resize_output_image_size does), then the centered crop x crop window."""
    short, long = (w, h) if w <= h else (h, w)
    new_long = int(size * long / short)
    new_h, new_w = (new_long, size) if w <= h else (size, new_long)
    top, left = (new_h - crop) // 2, (new_w - crop) // 2
    if top < 0 or left < 0:
        raise ValueError(f"image resized to {new_h}x{new_w} is smaller than the {crop}x{crop} crop")
    return new_h, new_w, top, left


class ClipPreprocessor:
    """`CLIPImageProcessor(size={"shortest_edge": s}, crop_size=c)` for decoded RGB uint8 images, on the GPU.

    `square=True` is the Qwen-VL / InternLM-XC2 transform instead (visual.py:354-362): torchvision
    `Resize((s, s), BICUBIC)` without keeping the aspect ratio, `ToTensor`, `Normalize` -- the same two kernels with the
    crop covering the whole resized image (uint8 / 255 in float32 equals the float64-multiply rescale for all 256 values)."""

    def __init__(self, size: int = 336, crop: int = 336, image_mean: Sequence[float] = OPENAI_CLIP_MEAN,
                 image_std: Sequence[float] = OPENAI_CLIP_STD, rescale_factor: float = 1 / 255, device: str = "cuda",
                 out_dtype: torch.dtype = torch.float32, square: bool = False):
        self.size, self.crop = int(size), int(size if square else crop)
        self.square = bool(square)
        self.rescale = float(rescale_factor)
        self.device = torch.device(device)
        self.out_dtype = out_dtype
        self._mean_std = np.ascontiguousarray(np.array(list(image_mean) + list(image_std), dtype=np.float32))
        self._dev_tables = {}
        self._ws: Optional[torch.Tensor] = None

    def _tables(self, in_size: int, out_size: int):
        key = (in_size, out_size)
        t = self._dev_tables.get(key)
        if t is None:
            ksize, bounds, coef = resample_tables(in_size, out_size)
            t = (ksize, torch.from_numpy(bounds).to(self.device), torch.from_numpy(coef).to(self.device), bounds)
            if len(self._dev_tables) > 512:
                self._dev_tables.clear()
            self._dev_tables[key] = t
        return t

    def one(self, image: Union[np.ndarray, torch.Tensor], out: torch.Tensor):
        """image: [H, W, 3] uint8 (numpy / CPU tensor / CUDA tensor) -> writes out [3, crop, crop]."""
        if isinstance(image, np.ndarray):
            image = torch.from_numpy(np.ascontiguousarray(image))
        if image.dtype != torch.uint8 or image.dim() != 3 or image.shape[2] != 3:
            raise ValueError(f"expected an [H, W, 3] uint8 RGB image, got {tuple(image.shape)} {image.dtype}")
        h, w = int(image.shape[0]), int(image.shape[1])
        new_h, new_w, top, left = (self.size, self.size, 0, 0) if self.square else resize_geometry(h, w, self.size, self.crop)
        kh, bh, ch, _ = self._tables(w, new_w)
        kv, bv, cv, bv_host = self._tables(h, new_h)
        row0 = int(bv_host[top, 0])
        last = bv_host[top + self.crop - 1]
        rows = int(last[0] + last[1]) - row0
        img = image.to(self.device, non_blocking=True).contiguous()
        need = rows * self.crop * 3
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        ops.clip_preprocess_u8(img, ch, bh, kh, cv, bv, kv, new_h, new_w, top, left, self.crop, self.crop, row0, rows,
                               self._ws, self.rescale, self._mean_std, out)
        return out

    def __call__(self, images: List[Union[np.ndarray, torch.Tensor]]) -> torch.Tensor:
        out = torch.empty(len(images), 3, self.crop, self.crop, dtype=self.out_dtype, device=self.device)
        for i, im in enumerate(images):
            self.one(im, out[i])
        return out


class AnyresPreprocessor:
    """`LlavaNextImageProcessor` (transformers 4.41) for decoded RGB uint8 images on the GPU: the base view (whole image
    resized to crop x crop) plus the crop x crop cells of the aspect-preserving resize pasted on the zero canvas of the best
    pinpoint.  Each view is one `vlb200_clip_preprocess_u8` call with the view's window in resized-image coordinates
    (negative / overshooting offsets read the canvas' zero padding).  -> (pixel_values [B, max_views, 3, c, c] zero padded
    like `_pad_for_batching`, image_sizes [B, 2])."""

    def __init__(self, pinpoints, crop: int = 336, image_mean: Sequence[float] = OPENAI_CLIP_MEAN,
                 image_std: Sequence[float] = OPENAI_CLIP_STD, rescale_factor: float = 1 / 255, device: str = "cuda",
                 out_dtype: torch.dtype = torch.float32):
        self.pinpoints = [tuple(int(x) for x in p) for p in pinpoints]
        self.inner = ClipPreprocessor(size=crop, crop=crop, image_mean=image_mean, image_std=image_std,
                                      rescale_factor=rescale_factor, device=device, out_dtype=out_dtype, square=True)
        self.crop = int(crop)

    def geometry(self, h: int, w: int):
        """-> (th, tw, nh, nw, py, px): pinpoint, aspect-preserving size inside it, paste offset."""
        from .host import best_resolution
        th, tw = best_resolution((h, w), self.pinpoints)
        sw, sh = tw / w, th / h
        if sw < sh:
            nw, nh = tw, min(math.ceil(h * sw), th)
        else:
            nh, nw = th, min(math.ceil(w * sh), tw)
        return th, tw, nh, nw, (th - nh) // 2, (tw - nw) // 2

    def _window(self, img: torch.Tensor, nh: int, nw: int, top: int, left: int, out: torch.Tensor):
        p, c = self.inner, self.crop
        h, w = int(img.shape[0]), int(img.shape[1])
        kh, bh, ch, _ = p._tables(w, nw)
        kv, bv, cv, bv_host = p._tables(h, nh)
        y0, y1 = max(top, 0), min(top + c, nh)  # resized rows the window touches
        if y1 > y0:
            row0 = int(bv_host[y0, 0])
            rows = int(bv_host[y1 - 1, 0] + bv_host[y1 - 1, 1]) - row0
        else:
            row0, rows = 0, 0
        need = max(rows * c * 3, 16)
        if p._ws is None or p._ws.numel() < need:
            p._ws = torch.empty(need, dtype=torch.uint8, device=p.device)
        ops.clip_preprocess_u8(img, ch, bh, kh, cv, bv, kv, nh, nw, top, left, c, c, row0, rows, p._ws, p.rescale, p._mean_std, out)

    def __call__(self, images: List[Union[np.ndarray, torch.Tensor]]):
        c, dev = self.crop, self.inner.device
        geo, views = [], []
        for im in images:
            h, w = int(im.shape[0]), int(im.shape[1])
            g = self.geometry(h, w)
            geo.append(g)
            views.append(1 + (g[0] // c) * (g[1] // c))
        out = torch.zeros(len(images), max(views), 3, c, c, dtype=self.inner.out_dtype, device=dev)
        for i, im in enumerate(images):
            if isinstance(im, np.ndarray):
                im = torch.from_numpy(np.ascontiguousarray(im))
            img = im.to(dev, non_blocking=True).contiguous()
            th, tw, nh, nw, py, px = geo[i]
            self._window(img, c, c, 0, 0, out[i, 0])  # base view: the whole image resized to c x c
            k = 1
            for r in range(0, th, c):
                for col in range(0, tw, c):
                    self._window(img, nh, nw, r - py, col - px, out[i, k])
                    k += 1
        sizes = torch.tensor([[int(im.shape[0]), int(im.shape[1])] for im in images], dtype=torch.int64)
        return out, sizes
