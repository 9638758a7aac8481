"""Host-side (CPU, integer) logic of the hot path -- mirrors of the reference's own Python:

  * `concatenated_inputs`  <- VLDPOTrainer.concatenated_inputs (base/trainer.py:124-146) + the trl-0.8.1
    parent it calls (pad chosen/rejected to a common length, concatenate chosen-then-rejected).
  * `ddpo_row_weights`     <- the mask_shared_tokens branch of get_batch_logps (base/trainer.py:169-184) over
    utils/diff_lib.get_diff_ids (difflib.SequenceMatcher, autojunk ON).  The reference runs this inside the
    step with a device sync (`.tolist()`); here it runs on the host batch before the step (collator side).
"""
from __future__ import annotations

import difflib
from typing import Dict, List, Optional, Tuple

import torch


def pad_to_length(t: torch.Tensor, length: int, pad_value, dim: int = -1) -> torch.Tensor:
    """utils/common.py:58-87 (right padding)."""
    if t.size(dim) >= length:
        return t
    pad_size = list(t.shape)
    pad_size[dim] = length - t.size(dim)
    return torch.cat([t, torch.full(pad_size, pad_value, dtype=t.dtype, device=t.device)], dim=dim)


def concatenated_inputs(batch: Dict, is_encoder_decoder: bool = False, label_pad_token_id: int = -100,
                        padding_value: int = 0, device=None) -> Dict:
    if is_encoder_decoder:
        raise ValueError("encoder-decoder models are not supported by the B200 path")
    max_length = max(batch["chosen_input_ids"].shape[1], batch["rejected_input_ids"].shape[1])
    out: Dict = {}
    for side in ("chosen", "rejected"):
        for k in ("input_ids", "attention_mask", "labels"):
            pad = label_pad_token_id if k == "labels" else (padding_value if k == "input_ids" else 0)
            t = pad_to_length(batch[f"{side}_{k}"], max_length, pad)
            ck = f"concatenated_{k}"
            out[ck] = t if side == "chosen" else torch.cat([out[ck], t], dim=0)
    if "img_input_dict" in batch:
        cat = {}
        for k, v in batch["img_input_dict"].items():
            if isinstance(v, torch.Tensor):
                cat[k] = torch.cat([v, v], dim=0)
            elif isinstance(v, list):
                cat[k] = v + v
            else:
                raise ValueError(f"Unsupported type {type(v)} for concatenation.")
        out["concatenated_img_input_dict"] = cat
    return out


def get_diff_ids(a_seq: List[int], b_seq: List[int], min_match_size: int = 3) -> Tuple[List[int], List[int]]:
    """utils/diff_lib.py:116-180: indices of tokens inside spans modified on BOTH sides."""
    mb = difflib.SequenceMatcher(None, a_seq, b_seq).get_matching_blocks()
    mb = [m for m in mb[:-1] if m[2] >= min_match_size] + [mb[-1]]
    a_ids: List[int] = []
    b_ids: List[int] = []
    ai = bi = 0
    for (i, j, n) in mb:  # the gap before each matching block (and before the sentinel) is a modification span
        if i > ai and j > bi:
            a_ids.extend(range(ai, i))
            b_ids.extend(range(bi, j))
        ai, bi = i + n, j + n
    return a_ids, b_ids


def ddpo_row_weights(input_ids: torch.Tensor, labels: torch.Tensor, image_token_index: int, n_patches: int,
                     label_pad_token_id: int = -100, min_match_size: int = 3) -> torch.Tensor:
    """uint8 weights [2B, L-1] for the text-level logits rows (row j-1 predicts text token j).

    The reference diffs the *merged* shifted label sequences (length S-1: image positions are -100 -> 0,
    trainer.py:161-166), so the merged sequences are rebuilt here exactly (autojunk depends on the length)."""
    ids = input_ids.cpu()
    lab = labels.cpu()
    n2, L = ids.shape
    assert n2 % 2 == 0
    n = n2 // 2
    out = torch.zeros(n2, L - 1, dtype=torch.uint8)

    def merged_shift(b):
        seq: List[int] = []
        row_pos: List[int] = []  # merged shifted index for text token j (>=1), -1 if none
        for j in range(L):
            t = int(ids[b, j])
            if t == image_token_index:
                start = len(seq)
                seq.extend([0] * n_patches)
                row_pos.append(start - 1)
            else:
                v = int(lab[b, j])
                row_pos.append(len(seq) - 1)
                seq.append(0 if v == label_pad_token_id else v)
        return seq[1:], row_pos

    for i in range(n):
        ca, pa = merged_shift(i)
        cb, pb = merged_shift(n + i)
        ia, ib = get_diff_ids(ca, cb, min_match_size)
        sa, sb = set(ia), set(ib)
        for j in range(1, L):
            if pa[j] in sa:
                out[i, j - 1] = 1
            if pb[j] in sb:
                out[n + i, j - 1] = 1
    return out
