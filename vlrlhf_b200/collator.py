"""DPO data collator of the B200 path.

Mirror of `VLDPODataCollatorWithPadding.__call__` (base/collator.py:26-68) and of the image step
`LlavaDPODataCollatorWithPadding.__call__` adds (models/Llava/__init__.py:435-443): same keys, same padding
(chosen/rejected right-padded, prompt left-padded, ids with pad_token_id, labels with label_pad_token_id, masks with
0), cached `*_logps` tensors, everything else passed through as lists.  The difference is where the pixels are made:
the reference runs CLIPImageProcessor on the CPU inside the collator; here the decoded RGB bytes are shipped as uint8
and `preprocess.ClipPreprocessor` (libvlb200) produces `img_input_dict["pixel_values"]` directly in HBM, bit-identical
to the CPU result.  Decoding the file (`PIL.Image.open(...).convert("RGB")`) stays on the host, as in the reference.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional

import numpy as np
import torch


def pad_rows(rows: List[List[int]], padding_value: int, left: bool = False) -> torch.Tensor:
    """pad_sequence(batch_first=True); `left` puts the padding in front (the reference reverses, pads, flips back)."""
    n = max(len(r) for r in rows)
    out = torch.full((len(rows), n), padding_value, dtype=torch.long)
    for i, r in enumerate(rows):
        if len(r) == 0:
            continue
        t = torch.as_tensor(r, dtype=torch.long)
        if left:
            out[i, n - len(r):] = t
        else:
            out[i, :len(r)] = t
    return out


def load_rgb(path: str) -> np.ndarray:
    """`Image.open(path).convert("RGB")` as an [H, W, 3] uint8 array (host decode, Llava/__init__.py:439)."""
    from PIL import Image
    with Image.open(path) as im:
        return np.array(im.convert("RGB"))  # writable copy


class B200DPODataCollatorWithPadding:
    def __init__(self, pad_token_id: int = 0, label_pad_token_id: int = -100, is_encoder_decoder: bool = False,
                 preprocessor: Optional[Callable[[List[np.ndarray]], torch.Tensor]] = None,
                 image_loader: Callable[[str], np.ndarray] = load_rgb):
        if is_encoder_decoder:
            raise ValueError("encoder-decoder models are not supported by the B200 path")
        self.pad_token_id = pad_token_id
        self.label_pad_token_id = label_pad_token_id
        self.is_encoder_decoder = is_encoder_decoder
        self.preprocessor = preprocessor
        self.image_loader = image_loader

    def pad(self, features: List[Dict[str, Any]]) -> Dict[str, Any]:
        batch: Dict[str, Any] = {}
        for k in features[0].keys():
            if k.endswith("_input_ids") or k.endswith("_attention_mask") or k.endswith("_labels"):
                if k.endswith("_input_ids"):
                    value = self.pad_token_id
                elif k.endswith("_labels"):
                    value = self.label_pad_token_id
                else:
                    value = 0
                batch[k] = pad_rows([list(ex[k]) for ex in features], value, left="prompt" in k)
            elif k.endswith("_logps"):
                batch[k] = torch.tensor([ex[k] for ex in features])
            else:
                batch[k] = [ex[k] for ex in features]
        return batch

    def __call__(self, features: List[Dict[str, Any]]) -> Dict[str, Any]:
        batch = self.pad(features)
        if "img_path" in batch:
            if self.preprocessor is None:
                raise RuntimeError("a preprocessor (vlrlhf_b200.preprocess.ClipPreprocessor) is required for images")
            images = [self.image_loader(p) for p in batch["img_path"]]
            pv = self.preprocessor(images)
            if isinstance(pv, tuple):  # LLaVA-Next: (pixel_values [B, views, 3, c, c], image_sizes [B, 2])
                batch["img_input_dict"] = dict(pixel_values=pv[0], image_sizes=pv[1])
            else:
                batch["img_input_dict"] = dict(pixel_values=pv)
        return batch
