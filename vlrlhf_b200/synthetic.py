"""Synthetic preference batches in the exact format VLDPODataCollatorWithPadding emits
(base/collator.py:26-68): right-padded chosen_/rejected_ input_ids / attention_mask / labels (int64) and
img_input_dict.pixel_values (fp32 [B,3,H,W]).  One <image> placeholder after BOS; labels = -100 on prompt
and padding.  (SURVEY.md §8d "Synthetic inputs".)"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from .config import ModelConfig


def make_batch(cfg: ModelConfig, n_pairs: int, text_len: int, prompt_len: int, seed: int, pin: bool = False,
               image_sizes=None) -> Dict:
    """LLaVA-Next (`cfg.family == "llava_next"`): pixel_values is the processor's [B, max_crops, 3, H, W] stack and
    `image_sizes` [B, 2] (height, width; default: square crop-sized images) rides along (LlavaNext/__init__.py:211-225)."""
    g = np.random.RandomState(seed)
    lo, hi = 3, min(cfg.image_token_index, cfg.vocab) - 1
    B, L = n_pairs, text_len
    prompt = g.randint(lo, hi, size=(B, prompt_len))
    prompt[:, 0] = 1
    prompt[:, 1] = cfg.image_token_index
    long_len = np.full(B, L)
    short_len = g.randint(int(0.75 * L), L + 1, size=B)
    swap = g.rand(B) < 0.5
    lens = {"chosen": np.where(swap, short_len, long_len), "rejected": np.where(swap, long_len, short_len)}
    out: Dict = {}
    for key in ("chosen", "rejected"):
        ids = np.full((B, L), cfg.pad_token_id, dtype=np.int64)
        mask = np.zeros((B, L), dtype=np.int64)
        labels = np.full((B, L), -100, dtype=np.int64)
        for b in range(B):
            n = int(lens[key][b])
            ids[b, :prompt_len] = prompt[b]
            ids[b, prompt_len:n] = g.randint(lo, hi, size=n - prompt_len)
            mask[b, :n] = 1
            labels[b, prompt_len:n] = ids[b, prompt_len:n]
        out[f"{key}_input_ids"] = torch.from_numpy(ids)
        out[f"{key}_attention_mask"] = torch.from_numpy(mask)
        out[f"{key}_labels"] = torch.from_numpy(labels)
    gen = torch.Generator().manual_seed(seed)
    if cfg.family == "llava_next":
        from .host import anyres_num_crops
        sizes = list(image_sizes) if image_sizes is not None else [(cfg.image_size, cfg.image_size)] * B
        crops = [anyres_num_crops(sz, cfg.image_grid_pinpoints, cfg.image_size) for sz in sizes]
        px = torch.randn(B, max(crops), 3, cfg.image_size, cfg.image_size, generator=gen)
        for b, c in enumerate(crops):
            px[b, c:] = 0
        out["img_input_dict"] = {"pixel_values": px, "image_sizes": torch.tensor(sizes, dtype=torch.int64)}
    else:
        out["img_input_dict"] = {"pixel_values": torch.randn(B, 3, cfg.image_size, cfg.image_size, generator=gen)}
    if pin:
        for k, v in list(out.items()):
            if isinstance(v, torch.Tensor):
                out[k] = v.pin_memory()
        out["img_input_dict"]["pixel_values"] = out["img_input_dict"]["pixel_values"].pin_memory()
    return out


def make_qwen_batch(cfg, n_pairs: int, text_len: int, prompt_len: int, seed: int, pin: bool = False) -> Dict:
    """Qwen-VL format: the prompt carries <img> + n_queries placeholder tokens + </img> (the reference spells the image
    path into the first placeholders, modeling_qwen.py:524-534); pixel_values [B, 3, 448, 448] ride in img_input_dict."""
    g = np.random.RandomState(seed)
    lo, hi = 3, min(cfg.image_start_id, cfg.pad_token_id, cfg.vocab) - 1
    B, L, Q = n_pairs, text_len, cfg.n_queries
    assert prompt_len >= Q + 3 and L > prompt_len + 4
    prompt = g.randint(lo, hi, size=(B, prompt_len))
    prompt[:, 0] = 1
    prompt[:, 1] = cfg.image_start_id
    prompt[:, 2:2 + Q] = cfg.image_start_id + 2
    prompt[:, 2 + Q] = cfg.image_start_id + 1
    long_len = np.full(B, L)
    short_len = g.randint(prompt_len + (L - prompt_len) * 3 // 4, L + 1, size=B)
    swap = g.rand(B) < 0.5
    lens = {"chosen": np.where(swap, short_len, long_len), "rejected": np.where(swap, long_len, short_len)}
    out: Dict = {}
    for key in ("chosen", "rejected"):
        ids = np.full((B, L), cfg.pad_token_id, dtype=np.int64)
        mask = np.zeros((B, L), dtype=np.int64)
        labels = np.full((B, L), -100, dtype=np.int64)
        for b in range(B):
            n = int(lens[key][b])
            ids[b, :prompt_len] = prompt[b]
            ids[b, prompt_len:n] = g.randint(lo, hi, size=n - prompt_len)
            mask[b, :n] = 1
            labels[b, prompt_len:n] = ids[b, prompt_len:n]
        out[f"{key}_input_ids"] = torch.from_numpy(ids)
        out[f"{key}_attention_mask"] = torch.from_numpy(mask)
        out[f"{key}_labels"] = torch.from_numpy(labels)
    gen = torch.Generator().manual_seed(seed)
    out["img_input_dict"] = {"pixel_values": torch.randn(B, 3, cfg.image_size, cfg.image_size, generator=gen)}
    if pin:
        for k, v in list(out.items()):
            if isinstance(v, torch.Tensor):
                out[k] = v.pin_memory()
        out["img_input_dict"]["pixel_values"] = out["img_input_dict"]["pixel_values"].pin_memory()
    return out
