"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- CPU restatement of the image side of the reference's DPO
collator: `LlavaDPODataCollatorWithPadding.__call__` (models/Llava/__init__.py:435-443) ->
`processor.image_processor(images=imgs, return_tensors="pt")`, i.e. transformers-4.41 `CLIPImageProcessor.preprocess`
(convert_rgb -> resize shortest edge, PIL bicubic -> center crop -> rescale 1/255 -> normalize).

Two third-party pieces are NOT under /root/reference and are restated from their published algorithms:
  * Pillow (pyproject.toml:20 "pillow", unpinned; the resampling code has been unchanged since Pillow 7.0):
    `Image.resize(size, BICUBIC)` = src/libImaging/Resample.c `ImagingResample` for 8-bit images:
    `precompute_coeffs` (double), `normalize_coeffs_8bpc` (fixed point, PRECISION_BITS = 22),
    `ImagingResampleHorizontal_8bpc` then `ImagingResampleVertical_8bpc` (int32 accumulate, +half, >>22, clip8).
  * transformers 4.41 image_transforms.py `get_resize_output_image_size`, `center_crop`, `rescale` (float64 multiply,
    float32 store), `normalize` ((x - mean) / std in float32).
Pinned (tests/test_preprocess_golden.py) against Pillow's own `Image.resize` and transformers' PIL-backend CLIP
processor run in this image, and against the committed fixture tests/golden/g7_clip_preprocess.npz made from them
(oracle/make_fixtures.py --preprocess).
"""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2  # Resample.c
OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def bicubic_filter(x: float) -> float:
    """Resample.c bicubic_filter, a = -0.5 (support 2.0)."""
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, in0: float, in1: float, out_size: int, support: float = 2.0):
    """Resample.c precompute_coeffs -> (ksize, bounds int32 [out,2] = (xmin, count), kk float64 [out, ksize])."""
    scale = filterscale = (in1 - in0) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = support * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), dtype=np.float64)
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        ww = 0.0
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = kk[xx]
        for x in range(xmax):
            w = bicubic_filter((x + xmin - center + 0.5) * ss)
            k[x] = w
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def normalize_coeffs_8bpc(kk: np.ndarray) -> np.ndarray:
    """Resample.c normalize_coeffs_8bpc: round-half-away fixed point with 22 fractional bits (C int truncation)."""
    scaled = kk * float(1 << PRECISION_BITS)
    return np.where(kk < 0, np.trunc(-0.5 + scaled), np.trunc(0.5 + scaled)).astype(np.int32)


def _clip8(v: np.ndarray) -> np.ndarray:
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resample_axis_u8(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One 8bpc pass along `axis` (1 = horizontal, 0 = vertical) of an [H, W, C] uint8 image."""
    in_size = img.shape[axis]
    ksize, bounds, kk = precompute_coeffs(in_size, 0.0, float(in_size), out_size)
    ki = normalize_coeffs_8bpc(kk)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx in range(out_size):
        xmin, n = bounds[xx]
        acc = np.tensordot(ki[xx, :n].astype(np.int64), src[xmin:xmin + n], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        out[xx] = _clip8(acc)
    return np.moveaxis(out, 0, axis)


def pil_bicubic_resize_u8(img: np.ndarray, out_hw: Tuple[int, int]) -> np.ndarray:
    """`PIL.Image.fromarray(img).resize((w, h), BICUBIC)` for an [H, W, C] uint8 array: the horizontal pass runs
    first (when the width changes), then the vertical pass on its uint8 result (when the height changes)."""
    oh, ow = out_hw
    if img.shape[1] != ow:
        img = resample_axis_u8(img, ow, axis=1)
    if img.shape[0] != oh:
        img = resample_axis_u8(img, oh, axis=0)
    return img


def shortest_edge_size(h: int, w: int, size: int) -> Tuple[int, int]:
    """image_transforms.get_resize_output_image_size(default_to_square=False): -> (new_h, new_w)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def clip_preprocess(img: np.ndarray, size: int = 336, crop: int = 336, mean: Sequence[float] = OPENAI_CLIP_MEAN,
                    std: Sequence[float] = OPENAI_CLIP_STD) -> np.ndarray:
    """[H, W, 3] uint8 RGB -> [3, crop, crop] float32, as CLIPImageProcessor (4.41 slow / 5.x PIL backend)."""
    h, w = img.shape[:2]
    nh, nw = shortest_edge_size(h, w, size)
    r = pil_bicubic_resize_u8(img, (nh, nw))
    top, left = (nh - crop) // 2, (nw - crop) // 2
    if top < 0 or left < 0:
        raise ValueError("center_crop padding branch (image smaller than the crop) is not on the CLIP-336 path")
    r = r[top:top + crop, left:left + crop]
    x = (r.astype(np.float64) * (1 / 255)).astype(np.float32)          # rescale: float64 multiply, float32 store
    x = (x - np.array(mean, dtype=np.float32)) / np.array(std, dtype=np.float32)  # normalize in float32
    return np.ascontiguousarray(x.transpose(2, 0, 1))


def square_preprocess(img: np.ndarray, size: int = 448, mean: Sequence[float] = OPENAI_CLIP_MEAN,
                      std: Sequence[float] = OPENAI_CLIP_STD) -> np.ndarray:
    """Qwen-VL / InternLM-XC2 image transform (models/QwenVL/visual.py:354-362, InternLMXC2/modeling_internlm_xcomposer2.py
    :67-73): torchvision Resize((size, size), BICUBIC) on the PIL image (= Pillow resize, aspect ratio NOT kept) ->
    ToTensor (uint8 / 255 in float32) -> Normalize ((x - mean) / std in float32).  [H, W, 3] uint8 -> [3, size, size]."""
    r = pil_bicubic_resize_u8(img, (size, size))
    x = r.astype(np.float32) / np.float32(255)
    x = (x - np.array(mean, dtype=np.float32)) / np.array(std, dtype=np.float32)
    return np.ascontiguousarray(x.transpose(2, 0, 1))


def anyres_preprocess(img: np.ndarray, pinpoints, crop: int = 336, mean: Sequence[float] = OPENAI_CLIP_MEAN,
                      std: Sequence[float] = OPENAI_CLIP_STD) -> np.ndarray:
    """transformers-4.41 LlavaNextImageProcessor.preprocess for one image (what LlavaNextProcessor hands the reference's
    collator, models/LlavaNext/__init__.py): select_best_resolution -> resize keeping the aspect ratio
    (get_patch_output_size) -> zero-pad the uint8 canvas to the chosen pinpoint (centered) -> cut crop x crop cells
    row-major; the base view is the whole image resized to crop x crop; every view is then rescaled and normalised.
    [H, W, 3] uint8 -> [1 + cells, 3, crop, crop] float32."""
    from .restate import select_best_resolution
    h, w = img.shape[:2]
    th, tw = select_best_resolution((h, w), [tuple(p) for p in pinpoints])
    sw, sh = tw / w, th / h
    if sw < sh:
        nw, nh = tw, min(math.ceil(h * sw), th)
    else:
        nh, nw = th, min(math.ceil(w * sh), tw)
    resized = pil_bicubic_resize_u8(img, (nh, nw))
    px, rx = divmod(tw - nw, 2)
    py, ry = divmod(th - nh, 2)
    canvas = np.pad(resized, ((py, py + ry), (px, px + rx), (0, 0)), mode="constant", constant_values=0)
    views = [pil_bicubic_resize_u8(img, (crop, crop))]
    for r in range(0, th, crop):
        for c in range(0, tw, crop):
            views.append(canvas[r:r + crop, c:c + crop])
    out = []
    for v in views:
        x = (v.astype(np.float64) * (1 / 255)).astype(np.float32)
        x = (x - np.array(mean, dtype=np.float32)) / np.array(std, dtype=np.float32)
        out.append(x.transpose(2, 0, 1))
    return np.ascontiguousarray(np.stack(out))


# ---- seeded synthetic RGB images shared by the fixture generator (make_fixtures.py --preprocess) and the tests
def synthetic_image(h: int, w: int, seed: int) -> np.ndarray:
    """[h, w, 3] uint8: smooth colour waves + noise + hard edges (exercises the negative bicubic lobes / clipping)."""
    rs = np.random.RandomState(seed)
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([127 + 120 * np.sin(x / (17.0 + 5 * c) + c) * np.cos(y / (23.0 - 3 * c) - c) for c in range(3)], -1)
    img = img + rs.randn(h, w, 3) * 20
    img[(x // 16 + y // 16) % 7 == 0] = 255.0 * (seed % 2)
    img[h // 3:h // 3 + 2] = 0
    return np.clip(img, 0, 255).astype(np.uint8)


G7_SIZES = [(480, 640), (640, 480), (336, 336), (500, 333), (97, 211), (768, 1024), (337, 336)]
