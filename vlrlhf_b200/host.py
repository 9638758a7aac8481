"""Host-side (CPU, integer) logic of the hot path -- mirrors of the reference's own Python:

  * `concatenated_inputs`  <- VLDPOTrainer.concatenated_inputs (base/trainer.py:124-146) + the trl-0.8.1
    parent it calls (pad chosen/rejected to a common length, concatenate chosen-then-rejected).
  * `anyres_*`             <- the integer side of LLaVA-Next's "spatial_unpad" packing (transformers-4.41
    modeling_llava_next.py get_anyres_image_grid_shape / image_size_to_num_patches / unpad_image / pack_image_features,
    called from models/LlavaNext/__init__.py:211-249): which projector-output row lands on which packed row.
  * `ddpo_row_weights`     <- the mask_shared_tokens branch of get_batch_logps (base/trainer.py:169-184) over
    utils/diff_lib.get_diff_ids (difflib.SequenceMatcher, autojunk ON).  The reference runs this inside the
    step with a device sync (`.tolist()`); here it runs on the host batch before the step (collator side).
"""
from __future__ import annotations

import difflib
from typing import Dict, List, Optional, Sequence, Tuple

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


def best_resolution(original_size: Sequence[int], pinpoints) -> Tuple[int, int]:
    """(height, width) pinpoint that keeps the most of the image and then wastes the least canvas."""
    oh, ow = int(original_size[0]), int(original_size[1])
    key = None
    best = None
    for h, w in pinpoints:
        sc = min(w / ow, h / oh)
        eff = min(int(ow * sc) * int(oh * sc), ow * oh)
        k = (eff, -(w * h - eff))
        if key is None or k > key:  # strict: the first pinpoint wins ties, as in the reference loop
            key, best = k, (int(h), int(w))
    return best


def anyres_grid(original_size, pinpoints, crop: int) -> Tuple[int, int]:
    h, w = best_resolution(original_size, pinpoints)
    return h // crop, w // crop


def anyres_num_crops(original_size, pinpoints, crop: int) -> int:
    """crops the processor emits for one image: ceil-grid cells of the chosen pinpoint + the base crop."""
    h, w = best_resolution(original_size, pinpoints)
    return -(-h // crop) * -(-w // crop) + 1


class AnyresPlan:
    """Row bookkeeping of one batch of anyres images (all host integers; `to(device)` stages the index vectors)."""
    __slots__ = ("crops", "feature_lens", "feat_off", "pack_index", "scatter_index", "newline_rows", "n_crop_rows",
                 "total_feats", "merged_len")

    def to(self, device):
        for k in ("feat_off", "pack_index", "scatter_index", "newline_rows"):
            setattr(self, k, getattr(self, k).to(device, non_blocking=True))
        return self


def anyres_pack_index(image_sizes, pinpoints, image_size: int, patch_size: int) -> AnyresPlan:
    """pack_image_features as an index: packed row r of image i reads projector-output row pack_index[r]
    (crop-major, `g*g` rows per crop, images concatenated) or the image_newline row (index n_crop_rows).

    Per image: the base crop's g*g rows; then the grid crops seen as one (gh*g) x (gw*g) map, cropped to the
    original aspect ratio (unpad_image) and read row by row with one newline after every map row."""
    g = image_size // patch_size
    P = g * g
    sizes = [(int(s[0]), int(s[1])) for s in (image_sizes.tolist() if hasattr(image_sizes, "tolist") else image_sizes)]
    plan = AnyresPlan()
    plan.crops = [anyres_num_crops(sz, pinpoints, image_size) for sz in sizes]
    plan.n_crop_rows = sum(plan.crops) * P
    NL = plan.n_crop_rows
    idx: List[int] = []
    lens: List[int] = []
    row0 = 0
    for (oh, ow), n in zip(sizes, plan.crops):
        start = len(idx)
        idx.extend(range(row0, row0 + P))  # base crop
        if n > 1:
            gh, gw = anyres_grid((oh, ow), pinpoints, image_size)
            H, W = gh * g, gw * g
            y0, y1, x0, x1 = 0, H, 0, W
            if ow / oh > W / H:  # wider than the canvas: rows were padded
                new_h = int(round(oh * (W / ow), 7))
                pad = (H - new_h) // 2
                y0, y1 = pad, H - pad
            else:
                new_w = int(round(ow * (H / oh), 7))
                pad = (W - new_w) // 2
                x0, x1 = pad, W - pad
            for y in range(y0, y1):
                cy, ry = divmod(y, g)
                for x in range(x0, x1):
                    cx, rx = divmod(x, g)
                    idx.append(row0 + (1 + cy * gw + cx) * P + ry * g + rx)
                idx.append(NL)
        else:
            idx.append(NL)
        lens.append(len(idx) - start)
        row0 += n * P
    pk = torch.tensor(idx, dtype=torch.int32)
    plan.pack_index = pk
    plan.scatter_index = torch.where(pk == NL, torch.full_like(pk, -1), pk)
    plan.newline_rows = torch.nonzero(pk == NL).flatten().to(torch.int32)
    plan.feature_lens = lens
    off = [0]
    for n in lens:
        off.append(off[-1] + n)
    plan.feat_off = torch.tensor(off, dtype=torch.int32)
    plan.total_feats = off[-1]
    return plan


def next_merged_len(input_ids: torch.Tensor, attention_mask: torch.Tensor, feature_lens: Sequence[int],
                    image_token_index: int, imgs_per_seq: int = 1) -> int:
    """max_embed_dim of LlavaNext/__init__.py:81-87: the longest valid merged sequence.  `feature_lens` holds one
    entry per image of the image batch; sequence b uses images (b % n_img_batch) * imgs_per_seq + slot."""
    n_img_batch = len(feature_lens) // imgs_per_seq
    per_img_seq = torch.tensor([sum(feature_lens[i * imgs_per_seq:(i + 1) * imgs_per_seq]) for i in range(n_img_batch)])
    n_seq = input_ids.shape[0]
    feat = per_img_seq[torch.arange(n_seq) % n_img_batch].to(input_ids.device)
    n_special = (input_ids == image_token_index).sum(-1)
    return int(((attention_mask == 1).sum(-1) - n_special + feat).max())


def right_pad_valid_tokens(input_ids: torch.Tensor, attention_mask: torch.Tensor, labels: torch.Tensor,
                           padding_value: int = 0, label_pad_token_id: int = -100, loss_type: str = "sigmoid"):
    """Left-padded (or otherwise interleaved) batches, SURVEY.md f-2: move every sequence's attended tokens to the front,
    order kept, and pad on the right.  Nothing the hot path returns depends on WHERE the padding sits: the reference derives
    the rotary positions from the mask (`position_ids = cumsum(mask) - 1`, Llava/__init__.py:98), masked keys are invisible
    to every attended query, and get_batch_logps sums over the labelled tokens only (base/trainer.py:185-188) -- so the
    per-sequence log-probs of the right-padded form are those of the original (checked against the reference's own
    LlavaForRL on a left-padded batch, tests/test_oracle_vs_reference.py).  Only the padding positions' own (unused) logits
    move.  Inputs whose masks are already prefixes are returned as they are.

    DDPO is the exception and is refused: the reference diffs the chosen and rejected LABEL sequences padding included
    (masked positions become token 0, base/trainer.py:166,177-180), so its shared-token mask -- and with it the log-probs --
    depends on the padding side (measured on the reference itself: up to 130 nats apart on a tiny batch)."""
    att = attention_mask == 1
    if not bool((att[:, 1:] & ~att[:, :-1]).any()):
        return input_ids, attention_mask, labels
    if loss_type == "ddpo":
        raise ValueError("loss_type='ddpo' needs right-padded batches (what VLDPODataCollatorWithPadding emits): the reference's "
                         "token diff runs over the padded label sequences and changes with the padding side")
    L = input_ids.shape[1]
    order = torch.argsort((~att).to(torch.int8), dim=1, stable=True)   # attended tokens first, original order kept
    tail = torch.arange(L, device=input_ids.device)[None, :] >= att.sum(1, keepdim=True)
    ids = torch.gather(input_ids, 1, order).masked_fill(tail, padding_value)
    lab = torch.gather(labels, 1, order).masked_fill(tail, label_pad_token_id)
    am = torch.gather(attention_mask, 1, order).masked_fill(tail, 0)
    return ids, am, lab


def validate_token_batch(input_ids: torch.Tensor, labels: Optional[torch.Tensor], vocab: int, image_token_index: Optional[int] = None,
                         imgs_per_seq: Optional[int] = None, label_pad_token_id: int = -100) -> None:
    """Host-side checks of one concatenated batch BEFORE anything is indexed on the device (no device sync: only CPU tensors
    are inspected; device-resident batches are the caller's responsibility).  Mirrors what the reference raises on:
    a wrong number of <image> placeholders per sequence (Llava/__init__.py:90-94 ValueError -- prompt truncation can cut
    the placeholder away) and token / label ids outside the embedding table (torch's embedding IndexError)."""
    if input_ids.device.type != "cpu":
        return
    if input_ids.numel() and (int(input_ids.min()) < 0 or int(input_ids.max()) >= vocab):
        raise IndexError(f"input_ids outside [0, {vocab}): min {int(input_ids.min())}, max {int(input_ids.max())}")
    if labels is not None and labels.device.type == "cpu" and labels.numel():
        lb = labels[labels != label_pad_token_id]
        if lb.numel() and (int(lb.min()) < 0 or int(lb.max()) >= vocab):
            raise IndexError(f"labels outside [0, {vocab}): min {int(lb.min())}, max {int(lb.max())}")
    if image_token_index is not None and imgs_per_seq is not None:
        cnt = (input_ids == image_token_index).sum(-1)
        if bool((cnt != imgs_per_seq).any()):
            raise ValueError(f"The input provided to the model are wrong. The number of image tokens is {int(cnt.sum())} while "
                             f"the number of image given to the model is {int(imgs_per_seq) * input_ids.shape[0]}. This prevents "
                             f"correct indexing and breaks batch generation. (per-sequence <image> counts {cnt.tolist()}, "
                             f"expected {int(imgs_per_seq)} each)")


def merged_seq_lens(input_ids: torch.Tensor, attention_mask: torch.Tensor, image_token_index: int, feat_rows) -> List[int]:
    """Merged length of every sequence's attended prefix -- what the merge kernels report as `seqlens` -- computed on the
    host so that a packed step (TrainConfig.pack_sequences) needs no device read-back: attended text tokens minus the
    <image> placeholders plus the image feature rows (Llava/__init__.py:44-47, LlavaNext/__init__.py:81-87).  `feat_rows`:
    feature rows per sequence (an int, or one entry per sequence).  Right padding only (the DPO collator's,
    base/collator.py:44-60); anything else raises the engine's left-padding ValueError."""
    ids, am = input_ids.cpu(), attention_mask.cpu()
    n_seq = ids.shape[0]
    att = am == 1
    if bool((att[:, 1:] & ~att[:, :-1]).any()):
        raise ValueError("attention_mask must be a right-padded prefix mask (left padding is not supported yet)")
    rows = torch.as_tensor(feat_rows, dtype=torch.int64).reshape(-1)
    rows = rows.expand(n_seq) if rows.numel() == 1 else rows
    if rows.numel() != n_seq:
        raise ValueError(f"feat_rows holds {rows.numel()} entries for {n_seq} sequences")
    is_img = ids == image_token_index
    if bool((is_img & ~att).any()):
        raise ValueError("an <image> placeholder lies outside the attended prefix of its sequence")
    return [int(v) for v in (att.sum(-1) - is_img.sum(-1) + rows)]


def shared_prefix_rows(input_ids: torch.Tensor, attention_mask: torch.Tensor, image_token_index: int, n_patches: int,
                       min_suffix: int = 1) -> List[int]:
    """Merged rows the chosen and the rejected sequence of every pair have in common (TrainConfig.share_prefix), from one
    concatenated HOST batch ([2B, L], chosen first -- base/trainer.py:124-146): the longest common token prefix of the two
    sequences, capped so that each keeps at least `min_suffix` rows of its own, expanded to merged rows (every <image>
    placeholder inside it stands for n_patches rows, Llava/__init__.py:44-52).  0 when nothing can be shared -- including
    the case where an <image> placeholder lies beyond the common prefix (the image rows then stay per sequence)."""
    ids, am = input_ids.cpu(), attention_mask.cpu()
    n_seq, L = ids.shape
    if n_seq % 2:
        raise ValueError("shared_prefix_rows: a concatenated batch holds an even number of sequences")
    B = n_seq // 2
    lens = (am == 1).sum(-1)
    same = (ids[:B] == ids[B:]) & (am[:B] == 1) & (am[B:] == 1)
    lcp = torch.cumprod(same.to(torch.int64), dim=-1).sum(-1)              # first mismatch / end of the shorter sequence
    lcp = torch.minimum(lcp, torch.minimum(lens[:B], lens[B:]) - int(min_suffix)).clamp(min=0)
    is_img = ids[:B] == image_token_index
    inside = torch.arange(L)[None, :] < lcp[:, None]
    n_in = (is_img & inside).sum(-1)
    n_all = (is_img & (am[:B] == 1)).sum(-1)
    rows = lcp + n_in * (int(n_patches) - 1)
    rows = torch.where(n_in == n_all, rows, torch.zeros_like(rows))       # every image of the pair inside the prefix, or no sharing
    return [int(v) for v in rows]


def ddpo_row_weights(input_ids: torch.Tensor, labels: torch.Tensor, image_token_index: int, n_patches,
                     label_pad_token_id: int = -100, min_match_size: int = 3,
                     attention_mask: Optional[torch.Tensor] = None, merged_len: Optional[int] = None) -> torch.Tensor:
    """uint8 weights [2B, L-1] for the text-level logits rows (row j-1 predicts text token j).

    The reference diffs the *merged* shifted label sequences (length S-1: image positions are -100 -> 0,
    trainer.py:161-166), so the merged sequences are rebuilt here exactly (autojunk depends on the length).
    `n_patches` is an int (LLaVA-1.5) or one packed feature length per sequence (LLaVA-Next).  With
    `attention_mask`/`merged_len` given the LLaVA-Next merge is mirrored: masked tokens are dropped and every
    sequence is padded with ignore labels to `merged_len` (LlavaNext/__init__.py:96-127)."""
    ids = input_ids.cpu()
    lab = labels.cpu()
    am = attention_mask.cpu() if attention_mask is not None else None
    n2, L = ids.shape
    assert n2 % 2 == 0
    n = n2 // 2
    out = torch.zeros(n2, L - 1, dtype=torch.uint8)
    per_seq = [int(n_patches)] * n2 if isinstance(n_patches, int) else [int(x) for x in n_patches]
    assert len(per_seq) == n2

    def merged_shift(b):
        seq: List[int] = []
        row_pos: List[int] = []  # merged shifted index for text token j (>=1), -1 if none
        for j in range(L):
            t = int(ids[b, j])
            if am is not None and int(am[b, j]) == 0:
                row_pos.append(-1)
            elif t == image_token_index:
                start = len(seq)
                seq.extend([0] * per_seq[b])
                row_pos.append(start - 1)
            else:
                v = int(lab[b, j])
                row_pos.append(len(seq) - 1)
                seq.append(0 if v == label_pad_token_id else v)
        if merged_len is not None:
            seq.extend([0] * (merged_len - len(seq)))
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


def ddpo_row_weights_native(input_ids: torch.Tensor, labels: torch.Tensor, image_token_index: int, n_patches,
                            label_pad_token_id: int = -100, min_match_size: int = 3,
                            attention_mask: Optional[torch.Tensor] = None, merged_len: Optional[int] = None) -> torch.Tensor:
    """`ddpo_row_weights` through the native host routine `vlb200_host_ddpo_row_weights` (a C++ restatement of
    difflib's matcher, ~50x faster than the pure-Python mirror above; same result bit for bit)."""
    from . import _lib
    L_ = _lib.load()
    ids = input_ids.cpu().contiguous()
    lab = labels.cpu().contiguous()
    am = attention_mask.cpu().contiguous() if attention_mask is not None else None
    assert ids.dtype == torch.int64 and lab.dtype == torch.int64 and (am is None or am.dtype == torch.int64)
    n2, L = ids.shape
    per_seq = torch.tensor([int(n_patches)] * n2 if isinstance(n_patches, int) else [int(x) for x in n_patches],
                           dtype=torch.int32)
    assert per_seq.numel() == n2
    out = torch.zeros(n2, L - 1, dtype=torch.uint8)
    _lib.check(L_.vlb200_host_ddpo_row_weights(ids.data_ptr(), am.data_ptr() if am is not None else None, lab.data_ptr(), n2,
                                               L, image_token_index, per_seq.data_ptr(),
                                               -1 if merged_len is None else int(merged_len), label_pad_token_id,
                                               min_match_size, out.data_ptr()))
    return out


def matching_blocks_native(a_seq: List[int], b_seq: List[int]) -> List[Tuple[int, int, int]]:
    """difflib.SequenceMatcher(None, a, b).get_matching_blocks() through `vlb200_host_matching_blocks`."""
    from . import _lib
    L_ = _lib.load()
    a = torch.tensor(a_seq, dtype=torch.int64)
    b = torch.tensor(b_seq, dtype=torch.int64)
    cap = min(len(a_seq), len(b_seq)) + 2
    out = torch.zeros(cap, 3, dtype=torch.int32)
    n = L_.vlb200_host_matching_blocks(a.data_ptr(), len(a_seq), b.data_ptr(), len(b_seq), out.data_ptr(), cap)
    if n < 0:
        raise RuntimeError(L_.vlb200_last_error().decode())
    return [tuple(r) for r in out[:n].tolist()]


# ------------------------------------------------------------------------------------------
# Qwen-VL position tables (models/QwenVL/visual.py:24-96): weight transforms done once at load time
# ------------------------------------------------------------------------------------------
def sincos_2d(embed_dim: int, grid: int) -> torch.Tensor:
    """get_2d_sincos_pos_embed: float32 [grid*grid, embed_dim] (w-coordinate first, as the reference's meshgrid)."""
    import numpy as np
    gh = np.arange(grid, dtype=np.float32)
    gw = np.arange(grid, dtype=np.float32)
    g = np.stack(np.meshgrid(gw, gh), axis=0).reshape([2, 1, grid, grid])

    def one(dim, pos):
        omega = np.arange(dim // 2, dtype=np.float32)
        omega /= dim / 2.0
        omega = 1.0 / 10000 ** omega
        out = np.einsum("m,d->md", pos.reshape(-1), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)

    return torch.from_numpy(np.concatenate([one(embed_dim // 2, g[0]), one(embed_dim // 2, g[1])], axis=1)).float()


def interpolate_pos_table(table: torch.Tensor, n_target: int) -> torch.Tensor:
    """get_abs_pos: bicubic interpolation (align_corners=False) of a square [g*g, C] table to n_target positions."""
    import math
    src, tgt = int(math.sqrt(table.shape[0])), int(math.sqrt(n_target))
    if src == tgt:
        return table
    return torch.nn.functional.interpolate(table.float().reshape(1, src, src, -1).permute(0, 3, 1, 2), size=(tgt, tgt),
                                           mode="bicubic", align_corners=False).permute(0, 2, 3, 1).flatten(0, 2)
