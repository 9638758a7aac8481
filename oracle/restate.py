"""CPU restatement (torch fp32 / numpy) of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Every function cites the reference file:line (relative to /root/reference/) or the
installed-transformers source it follows.  Pinned against the reference's own functions by
`tests/test_oracle_vs_reference.py` (build container) and the vectors in `tests/golden/`
(minted by `oracle/make_fixtures.py` from the *reference* functions, not from this file).

Third-party pieces that are NOT on disk and are restated from their published algorithm
("parity unpinned" for these two only): trl==0.8.1 `DPOTrainer.concatenated_inputs` and
`DPOTrainer.get_batch_loss_metrics`.
"""
from __future__ import annotations

import difflib
import math
import zlib
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# configuration of the LLaVA-1.5 family
# --------------------------------------------------------------------------------------


@dataclass
class LlavaCfg:
    # vision tower: CLIP ViT (transformers modeling_clip.py)
    image_size: int = 336
    patch_size: int = 14
    v_hidden: int = 1024
    v_layers: int = 24
    v_heads: int = 16
    v_ff: int = 4096
    v_eps: float = 1e-5
    vision_feature_layer: int = -2  # Llava/__init__.py:163-165,180
    # decoder: Llama (transformers modeling_llama.py)
    hidden: int = 4096
    layers: int = 32
    heads: int = 32
    kv_heads: int = 32
    ff: int = 11008
    vocab: int = 32064
    rms_eps: float = 1e-5
    rope_theta: float = 10000.0
    image_token_index: int = 32000
    pad_token_id: int = 32001
    ignore_index: int = -100
    # "llava" (models/Llava) | "llava_next" (models/LlavaNext: anyres crops, image_newline, Mistral/Vicuna decoder)
    family: str = "llava"
    image_grid_pinpoints: Tuple[Tuple[int, int], ...] = ()

    @property
    def n_patches(self) -> int:
        return (self.image_size // self.patch_size) ** 2

    @property
    def v_used_layers(self) -> int:
        # hidden_states[vision_feature_layer]: index into [emb, l0, ..., l_{n-1}]
        n = self.v_layers
        idx = self.vision_feature_layer if self.vision_feature_layer >= 0 else n + 1 + self.vision_feature_layer
        return idx  # number of encoder layers whose output is needed

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads

    @property
    def v_head_dim(self) -> int:
        return self.v_hidden // self.v_heads


LLAVA15_7B = LlavaCfg()

TINY = LlavaCfg(image_size=28, patch_size=14, v_hidden=128, v_layers=3, v_heads=2, v_ff=256,
                hidden=128, layers=2, heads=2, kv_heads=2, ff=256, vocab=320,
                image_token_index=300, pad_token_id=301)

# a config that exercises the real tile shapes (head dims 64 / 128, several tiles per GEMM)
SMALL = LlavaCfg(image_size=112, patch_size=14, v_hidden=256, v_layers=3, v_heads=4, v_ff=512,
                 hidden=512, layers=2, heads=4, kv_heads=4, ff=1024, vocab=2048,
                 image_token_index=2000, pad_token_id=2001)


# LLaVA-Next (llava-v1.6-mistral-7b-hf config.json): CLIP-L/336 tower, Mistral-7B decoder (GQA 32/8, ff 14336,
# rope theta 1e6, no sliding window in -Instruct-v0.2), anyres pinpoints
LLAVA_NEXT_PINPOINTS = ((336, 672), (672, 336), (672, 672), (1008, 336), (336, 1008))
LLAVANEXT_MISTRAL_7B = LlavaCfg(hidden=4096, layers=32, heads=32, kv_heads=8, ff=14336, vocab=32064, rope_theta=1e6,
                                image_token_index=32000, pad_token_id=32001, family="llava_next",
                                image_grid_pinpoints=LLAVA_NEXT_PINPOINTS)
_TINY_PINS = ((28, 56), (56, 28), (56, 56), (84, 28), (28, 84))
TINY_NEXT = LlavaCfg(image_size=28, patch_size=14, v_hidden=128, v_layers=3, v_heads=2, v_ff=256,
                     hidden=256, layers=2, heads=4, kv_heads=2, ff=256, vocab=320, rope_theta=1e6,
                     image_token_index=300, pad_token_id=301, family="llava_next", image_grid_pinpoints=_TINY_PINS)
_SMALL_PINS = ((112, 224), (224, 112), (224, 224), (336, 112), (112, 336))
SMALL_NEXT = LlavaCfg(image_size=112, patch_size=14, v_hidden=256, v_layers=3, v_heads=4, v_ff=512,
                      hidden=512, layers=2, heads=4, kv_heads=2, ff=1024, vocab=2048, rope_theta=1e6,
                      image_token_index=2000, pad_token_id=2001, family="llava_next",
                      image_grid_pinpoints=_SMALL_PINS)


# --------------------------------------------------------------------------------------
# deterministic, device-independent weight / data generator (bit-exact twin of the CUDA
# kernel `vlb200_init_uniform` so 7B-shape weights never have to travel)
# --------------------------------------------------------------------------------------

def _lowbias32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint32, copy=True)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def hash_uniform(n: int, seed: int, scale: float, shift: float = 0.0, chunk: int = 1 << 24) -> torch.Tensor:
    """w[i] = shift + scale * (2 * (h(i ^ h(seed)) >> 8) * 2^-24 - 1), fp32, i < 2^32."""
    out = np.empty(n, dtype=np.float32)
    s = _lowbias32(np.array([seed & 0xFFFFFFFF], dtype=np.uint32))[0]
    with np.errstate(over="ignore"):
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            idx = np.arange(lo, hi, dtype=np.uint64).astype(np.uint32)
            h = _lowbias32(idx ^ s)
            u = (h >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
            u = u * np.float32(2.0) - np.float32(1.0)
            out[lo:hi] = u * np.float32(scale) + np.float32(shift)
    return torch.from_numpy(out)


def tensor_seed(name: str, base_seed: int) -> int:
    return (zlib.crc32(name.encode()) ^ (base_seed * 0x9E3779B1)) & 0xFFFFFFFF


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def weight_specs(cfg: LlavaCfg) -> List[Tuple[str, Tuple[int, ...], float, float]]:
    """(HF-4.41 parameter name, shape, uniform half-width, shift) for every tensor of the model."""
    a = 0.02 * math.sqrt(3.0)
    specs: List[Tuple[str, Tuple[int, ...], float, float]] = []
    vp = "vision_tower.vision_model."
    specs += [
        (vp + "embeddings.class_embedding", (cfg.v_hidden,), a, 0.0),
        (vp + "embeddings.patch_embedding.weight", (cfg.v_hidden, 3, cfg.patch_size, cfg.patch_size), a, 0.0),
        (vp + "embeddings.position_embedding.weight", (cfg.n_patches + 1, cfg.v_hidden), a, 0.0),
        (vp + "pre_layrnorm.weight", (cfg.v_hidden,), 0.1, 1.0),
        (vp + "pre_layrnorm.bias", (cfg.v_hidden,), 0.02, 0.0),
    ]
    for i in range(cfg.v_layers):
        p = f"{vp}encoder.layers.{i}."
        for ln in ("layer_norm1", "layer_norm2"):
            specs += [(p + ln + ".weight", (cfg.v_hidden,), 0.1, 1.0), (p + ln + ".bias", (cfg.v_hidden,), 0.02, 0.0)]
        for pr in ("q_proj", "k_proj", "v_proj", "out_proj"):
            specs += [(p + f"self_attn.{pr}.weight", (cfg.v_hidden, cfg.v_hidden), a, 0.0),
                      (p + f"self_attn.{pr}.bias", (cfg.v_hidden,), 0.02, 0.0)]
        specs += [(p + "mlp.fc1.weight", (cfg.v_ff, cfg.v_hidden), a, 0.0), (p + "mlp.fc1.bias", (cfg.v_ff,), 0.02, 0.0),
                  (p + "mlp.fc2.weight", (cfg.v_hidden, cfg.v_ff), a, 0.0), (p + "mlp.fc2.bias", (cfg.v_hidden,), 0.02, 0.0)]
    specs += [
        ("multi_modal_projector.linear_1.weight", (cfg.hidden, cfg.v_hidden), a, 0.0),
        ("multi_modal_projector.linear_1.bias", (cfg.hidden,), 0.02, 0.0),
        ("multi_modal_projector.linear_2.weight", (cfg.hidden, cfg.hidden), a, 0.0),
        ("multi_modal_projector.linear_2.bias", (cfg.hidden,), 0.02, 0.0),
        ("language_model.model.embed_tokens.weight", (cfg.vocab, cfg.hidden), a, 0.0),
    ]
    if cfg.family == "llava_next":  # nn.Parameter of LlavaNextForConditionalGeneration (modeling_llava_next.py)
        specs.append(("image_newline", (cfg.hidden,), a, 0.0))
    kv = cfg.kv_heads * cfg.head_dim
    for i in range(cfg.layers):
        p = f"language_model.model.layers.{i}."
        specs += [
            (p + "input_layernorm.weight", (cfg.hidden,), 0.1, 1.0),
            (p + "self_attn.q_proj.weight", (cfg.heads * cfg.head_dim, cfg.hidden), a, 0.0),
            (p + "self_attn.k_proj.weight", (kv, cfg.hidden), a, 0.0),
            (p + "self_attn.v_proj.weight", (kv, cfg.hidden), a, 0.0),
            (p + "self_attn.o_proj.weight", (cfg.hidden, cfg.heads * cfg.head_dim), a, 0.0),
            (p + "post_attention_layernorm.weight", (cfg.hidden,), 0.1, 1.0),
            (p + "mlp.gate_proj.weight", (cfg.ff, cfg.hidden), a, 0.0),
            (p + "mlp.up_proj.weight", (cfg.ff, cfg.hidden), a, 0.0),
            (p + "mlp.down_proj.weight", (cfg.hidden, cfg.ff), a, 0.0),
        ]
    specs += [("language_model.model.norm.weight", (cfg.hidden,), 0.1, 1.0),
              # wider lm_head so logits are O(1..10) and exercise the softmax range (SURVEY §8c v)
              ("language_model.lm_head.weight", (cfg.vocab, cfg.hidden), 3.0 * a, 0.0)]
    return specs


def make_weights(cfg: LlavaCfg, seed: int, names: Optional[List[str]] = None) -> Dict[str, torch.Tensor]:
    """bf16-representable fp32 weights, identical to what `Engine.init_synthetic(seed)` builds on the GPU."""
    out = {}
    for name, shape, scale, shift in weight_specs(cfg):
        if names is not None and name not in names:
            continue
        n = int(np.prod(shape))
        out[name] = bf16_round(hash_uniform(n, tensor_seed(name, seed), scale, shift)).reshape(shape)
    return out


# --------------------------------------------------------------------------------------
# synthetic preference batch (SURVEY.md §8d "Synthetic inputs")
# --------------------------------------------------------------------------------------

def make_batch(cfg: LlavaCfg, n_pairs: int, text_len: int, prompt_len: int, seed: int,
               ddpo_like: bool = False, image_sizes: Optional[List[Tuple[int, int]]] = None) -> Dict[str, torch.Tensor]:
    """Collated batch exactly as VLDPODataCollatorWithPadding emits it (base/collator.py:26-68):
    right-padded chosen_/rejected_ input_ids / attention_mask / labels + img_input_dict.pixel_values.
    One <image> placeholder at position 1 (after BOS); labels = -100 on prompt and padding."""
    g = np.random.RandomState(seed)
    lo, hi = 3, min(cfg.image_token_index, cfg.vocab) - 1
    B, L = n_pairs, text_len
    out = {}
    prompt = g.randint(lo, hi, size=(B, prompt_len))
    prompt[:, 0] = 1
    prompt[:, 1] = cfg.image_token_index
    chosen_len = np.full(B, L)
    rejected_len = g.randint(int(0.75 * L), L + 1, size=B)
    swap = g.rand(B) < 0.5
    chosen_len, rejected_len = np.where(swap, rejected_len, chosen_len), np.where(swap, chosen_len, rejected_len)
    base_resp = g.randint(lo, hi, size=(B, L))
    for key, lens in (("chosen", chosen_len), ("rejected", rejected_len)):
        ids = np.full((B, L), cfg.pad_token_id, dtype=np.int64)
        mask = np.zeros((B, L), dtype=np.int64)
        labels = np.full((B, L), -100, dtype=np.int64)
        for b in range(B):
            n = int(lens[b])
            resp = base_resp[b].copy() if ddpo_like else g.randint(lo, hi, size=L)
            if ddpo_like and key == "rejected":
                # a few span substitutions so the DDPO diff is non-trivial
                for _ in range(3):
                    s = g.randint(prompt_len, max(prompt_len + 1, n - 8))
                    w = g.randint(1, 6)
                    resp[s:s + w] = g.randint(lo, hi, size=len(resp[s:s + w]))
            ids[b, :prompt_len] = prompt[b]
            ids[b, prompt_len:n] = resp[prompt_len:n]
            mask[b, :n] = 1
            labels[b, prompt_len:n] = ids[b, prompt_len:n]
        out[f"{key}_input_ids"] = torch.from_numpy(ids)
        out[f"{key}_attention_mask"] = torch.from_numpy(mask)
        out[f"{key}_labels"] = torch.from_numpy(labels)
    if cfg.family == "llava_next":
        # LlavaNextProcessor output (models/LlavaNext/__init__.py:348-380 collate): pixel_values [B, max_crops, 3, H, W]
        # (base crop first, zero-padded to the largest crop count) + image_sizes [B, 2] = (height, width)
        sizes = image_sizes if image_sizes is not None else [(cfg.image_size, cfg.image_size)] * B
        assert len(sizes) == B
        crops = [image_size_to_num_patches(sz, cfg.image_grid_pinpoints, cfg.image_size) for sz in sizes]
        mc = max(crops)
        n_pix = B * mc * 3 * cfg.image_size * cfg.image_size
        pix = bf16_round(hash_uniform(n_pix, tensor_seed("pixel_values", seed), 1.7320508)).reshape(
            B, mc, 3, cfg.image_size, cfg.image_size).clone()
        for b in range(B):
            pix[b, crops[b]:] = 0
        out["img_input_dict"] = {"pixel_values": pix, "image_sizes": torch.tensor(sizes, dtype=torch.int64)}
        return out
    n_pix = B * 3 * cfg.image_size * cfg.image_size
    pix = bf16_round(hash_uniform(n_pix, tensor_seed("pixel_values", seed), 1.7320508)).reshape(
        B, 3, cfg.image_size, cfg.image_size)
    out["img_input_dict"] = {"pixel_values": pix}
    return out


# --------------------------------------------------------------------------------------
# trl 0.8.1 DPOTrainer.concatenated_inputs (NOT on disk; restated) + the reference override
# --------------------------------------------------------------------------------------

def pad_to_length(t: torch.Tensor, length: int, pad_value, dim: int = -1) -> torch.Tensor:
    """utils/common.py:58-87 (right padding branch)."""
    if t.size(dim) >= length:
        return t
    pad_size = list(t.shape)
    pad_size[dim] = length - t.size(dim)
    return torch.cat([t, pad_value * torch.ones(*pad_size, dtype=t.dtype)], dim=dim)


def concatenated_inputs(batch: Dict, label_pad_token_id: int = -100, padding_value: int = 0) -> Dict:
    """trl 0.8.1 DPOTrainer.concatenated_inputs (decoder-only branch) followed by the reference's
    image duplication (base/trainer.py:124-146): pad chosen/rejected to a common length with
    padding_value / label_pad_token_id / 0, concatenate chosen-then-rejected on dim 0, and
    duplicate every tensor of img_input_dict ([v, v])."""
    max_length = max(batch["chosen_input_ids"].shape[1], batch["rejected_input_ids"].shape[1])
    out = {}
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


# --------------------------------------------------------------------------------------
# utils/diff_lib.py:73-83,116-180 over CPython difflib (the reference's own dependency)
# --------------------------------------------------------------------------------------

def get_diff_ids(a_seq: List[int], b_seq: List[int], min_match_size: int = 3) -> Tuple[List[int], List[int]]:
    sm = difflib.SequenceMatcher(None, a_seq, b_seq)  # autojunk default ON (diff_lib.py:117)
    mb = sm.get_matching_blocks()
    mb = [m for m in mb[:-1] if m[2] >= min_match_size] + [mb[-1]]  # :121
    a_matches = [(x[0], x[0] + x[2]) for x in mb]
    b_matches = [(x[1], x[1] + x[2]) for x in mb]

    def complete(matches, length):  # :73-83
        i, j = 0, matches[0][0]
        out = []
        for idx in range(len(matches)):
            out.append((i, j))
            out.append(matches[idx])
            if idx + 1 < len(matches):
                i, j = matches[idx][1], matches[idx + 1][0]
            else:
                i, j = matches[idx][1], length
        return out

    a_spans, b_spans = complete(a_matches, len(a_seq)), complete(b_matches, len(b_seq))
    mod_map = {}
    for idx, (sa, sb) in enumerate(zip(a_spans, b_spans)):  # :136-155
        if idx % 2 == 1:
            continue
        if sa[0] != sa[1] and sb[0] != sb[1]:
            mod_map[sa] = sb
    a_ids = sorted({i for s in mod_map.keys() for i in range(s[0], s[1])})
    b_ids = sorted({i for s in mod_map.values() for i in range(s[0], s[1])})
    return a_ids, b_ids


def ddpo_shared_mask(shift_labels: torch.Tensor, min_match_size: int = 3) -> torch.Tensor:
    """base/trainer.py:169-184: bool mask [2B, S-1]; True where the token belongs to a modified span.
    `shift_labels` already has -100 rewritten to 0 (:166)."""
    n = shift_labels.shape[0] // 2
    assert n * 2 == shift_labels.shape[0]
    mask = torch.zeros_like(shift_labels, dtype=torch.bool)
    for i in range(n):
        c_mod, r_mod = get_diff_ids(shift_labels[i].tolist(), shift_labels[n + i].tolist(), min_match_size)
        mask[i, c_mod] = True
        mask[n + i, r_mod] = True
    return mask


# --------------------------------------------------------------------------------------
# base/trainer.py:148-188  get_batch_logps
# --------------------------------------------------------------------------------------

def get_batch_logps(logits: torch.Tensor, labels: torch.Tensor, average_log_prob: bool = False,
                    label_pad_token_id: int = -100, mask_shared_tokens: bool = False,
                    return_per_token: bool = False):
    if logits.shape[:-1] != labels.shape:
        raise ValueError("Logits (batch and sequence length dim) and labels must have the same shape.")
    labels = labels[:, 1:].clone()
    logits = logits[:, :-1, :]
    loss_mask = labels != label_pad_token_id
    labels[labels == label_pad_token_id] = 0
    per_token = torch.gather(logits.log_softmax(-1), dim=2, index=labels.unsqueeze(2)).squeeze(2)
    if mask_shared_tokens:
        loss_mask = loss_mask & ddpo_shared_mask(labels)
    if return_per_token:
        return per_token, loss_mask
    if average_log_prob:
        return (per_token * loss_mask).sum(-1) / loss_mask.sum(-1)
    return (per_token * loss_mask).sum(-1)


# --------------------------------------------------------------------------------------
# base/trainer.py:244-301  dpo_loss
# --------------------------------------------------------------------------------------

def dpo_loss(pc, pr, rc, rr, beta: float = 0.1, label_smoothing: float = 0.0, loss_type: str = "sigmoid",
             reference_free: bool = False):
    pi_logratios = pc - pr
    ref_logratios = torch.zeros(1, dtype=pi_logratios.dtype) if reference_free else rc - rr
    logits = pi_logratios - ref_logratios
    if loss_type in ("sigmoid", "ddpo"):
        losses = -F.logsigmoid(beta * logits) * (1 - label_smoothing) - F.logsigmoid(-beta * logits) * label_smoothing
    elif loss_type == "hinge":
        losses = torch.relu(1 - beta * logits)
    elif loss_type == "ipo":
        losses = (logits - 1 / (2 * beta)) ** 2
    elif loss_type == "kto_pair":
        chosen_KL = (pc - rc).mean().clamp(min=0)
        rejected_KL = (pr - rr).mean().clamp(min=0)
        chosen_logratios = pc - rc
        rejected_logratios = pr - rr
        losses = torch.cat((1 - torch.sigmoid(beta * (chosen_logratios - rejected_KL)),
                            1 - torch.sigmoid(beta * (chosen_KL - rejected_logratios))), 0)
    else:
        raise ValueError(f"Unknown loss type: {loss_type}. Should be one of ['sigmoid', 'hinge', 'ipo', 'kto_pair']")
    chosen_rewards = beta * (pc - rc).detach()
    rejected_rewards = beta * (pr - rr).detach()
    return losses, chosen_rewards, rejected_rewards


# --------------------------------------------------------------------------------------
# models/Llava/__init__.py:36-109  _merge_input_ids_with_image_features
# --------------------------------------------------------------------------------------

def merge_input_ids_with_image_features(cfg: LlavaCfg, image_features, inputs_embeds, input_ids, attention_mask,
                                        labels):
    num_images, num_image_patches, embed_dim = image_features.shape
    batch_size, sequence_length = input_ids.shape
    left_padding = not torch.sum(input_ids[:, -1] == torch.tensor(cfg.pad_token_id))
    special = input_ids == cfg.image_token_index
    num_special = torch.sum(special, dim=-1)
    max_embed_dim = int(num_special.max() * (num_image_patches - 1)) + sequence_length
    batch_indices, non_image_indices = torch.where(input_ids != cfg.image_token_index)
    new_token_positions = torch.cumsum((special * (num_image_patches - 1) + 1), -1) - 1
    nb_image_pad = max_embed_dim - 1 - new_token_positions[:, -1]
    if left_padding:
        new_token_positions = new_token_positions + nb_image_pad[:, None]
    text_to_overwrite = new_token_positions[batch_indices, non_image_indices]
    final_embedding = torch.zeros(batch_size, max_embed_dim, embed_dim, dtype=inputs_embeds.dtype)
    final_attention_mask = torch.zeros(batch_size, max_embed_dim, dtype=attention_mask.dtype)
    final_labels = torch.full((batch_size, max_embed_dim), cfg.ignore_index, dtype=input_ids.dtype)
    final_embedding[batch_indices, text_to_overwrite] = inputs_embeds[batch_indices, non_image_indices]
    final_attention_mask[batch_indices, text_to_overwrite] = attention_mask[batch_indices, non_image_indices]
    final_labels[batch_indices, text_to_overwrite] = labels[batch_indices, non_image_indices]
    image_to_overwrite = torch.all(final_embedding == 0, dim=-1)
    image_to_overwrite &= image_to_overwrite.cumsum(-1) - 1 >= nb_image_pad[:, None]
    if image_to_overwrite.sum() != image_features.shape[:-1].numel():
        raise ValueError("The input provided to the model are wrong. The number of image tokens is "
                         f"{torch.sum(special)} while the number of image given to the model is {num_images}.")
    final_embedding[image_to_overwrite] = image_features.contiguous().reshape(-1, embed_dim)
    final_attention_mask |= image_to_overwrite
    position_ids = (final_attention_mask.cumsum(-1) - 1).masked_fill_((final_attention_mask == 0), 1)
    batch_indices, pad_indices = torch.where(input_ids == cfg.pad_token_id)
    indices_to_mask = new_token_positions[batch_indices, pad_indices]
    final_embedding[batch_indices, indices_to_mask] = 0
    return final_embedding, final_attention_mask, final_labels, position_ids, image_to_overwrite


# --------------------------------------------------------------------------------------
# transformers CLIPVisionModel (modeling_clip.py:138-219 embeddings, 261-279 attention,
# 347-386 MLP/encoder layer, 647-692 vision transformer) -- hidden_states[vision_feature_layer]
# --------------------------------------------------------------------------------------

def clip_vision_features(cfg: LlavaCfg, w: Dict[str, torch.Tensor], pixel_values: torch.Tensor) -> torch.Tensor:
    p = "vision_tower.vision_model."
    B = pixel_values.shape[0]
    x = F.conv2d(pixel_values, w[p + "embeddings.patch_embedding.weight"], stride=cfg.patch_size)
    x = x.flatten(2).transpose(1, 2)  # [B, n_patches, d]
    cls = w[p + "embeddings.class_embedding"].expand(B, 1, -1)
    x = torch.cat([cls, x], dim=1) + w[p + "embeddings.position_embedding.weight"][None]
    x = F.layer_norm(x, (cfg.v_hidden,), w[p + "pre_layrnorm.weight"], w[p + "pre_layrnorm.bias"], cfg.v_eps)
    H, dh = cfg.v_heads, cfg.v_head_dim
    for i in range(cfg.v_used_layers):
        q = f"{p}encoder.layers.{i}."
        h = F.layer_norm(x, (cfg.v_hidden,), w[q + "layer_norm1.weight"], w[q + "layer_norm1.bias"], cfg.v_eps)
        S = h.shape[1]
        qq = F.linear(h, w[q + "self_attn.q_proj.weight"], w[q + "self_attn.q_proj.bias"]).view(B, S, H, dh).transpose(1, 2)
        kk = F.linear(h, w[q + "self_attn.k_proj.weight"], w[q + "self_attn.k_proj.bias"]).view(B, S, H, dh).transpose(1, 2)
        vv = F.linear(h, w[q + "self_attn.v_proj.weight"], w[q + "self_attn.v_proj.bias"]).view(B, S, H, dh).transpose(1, 2)
        att = torch.softmax(qq @ kk.transpose(-1, -2) * dh ** -0.5, dim=-1)
        o = (att @ vv).transpose(1, 2).reshape(B, S, cfg.v_hidden)
        x = x + F.linear(o, w[q + "self_attn.out_proj.weight"], w[q + "self_attn.out_proj.bias"])
        h = F.layer_norm(x, (cfg.v_hidden,), w[q + "layer_norm2.weight"], w[q + "layer_norm2.bias"], cfg.v_eps)
        h = F.linear(h, w[q + "mlp.fc1.weight"], w[q + "mlp.fc1.bias"])
        h = h * torch.sigmoid(1.702 * h)  # quick_gelu (activations.py QuickGELUActivation)
        x = x + F.linear(h, w[q + "mlp.fc2.weight"], w[q + "mlp.fc2.bias"])
    return x  # [B, 1+n_patches, d]; the caller drops CLS (Llava/__init__.py:182-183)


def projector(cfg: LlavaCfg, w: Dict[str, torch.Tensor], feats: torch.Tensor) -> torch.Tensor:
    """LlavaMultiModalProjector (modeling_llava.py:87-107): linear_1 -> GELU(erf) -> linear_2."""
    h = F.linear(feats, w["multi_modal_projector.linear_1.weight"], w["multi_modal_projector.linear_1.bias"])
    h = F.gelu(h)
    return F.linear(h, w["multi_modal_projector.linear_2.weight"], w["multi_modal_projector.linear_2.bias"])


# --------------------------------------------------------------------------------------
# transformers LlamaModel / LlamaForCausalLM (modeling_llama.py:53-67 RMSNorm, 138-168 RoPE,
# 171-184 MLP, 199-290 attention, 292-333 decoder layer)
# --------------------------------------------------------------------------------------

def rms_norm(x, weight, eps):
    var = x.float().pow(2).mean(-1, keepdim=True)
    return weight * (x.float() * torch.rsqrt(var + eps))


def rope_cos_sin(cfg: LlavaCfg, position_ids: torch.Tensor):
    dh = cfg.head_dim
    inv_freq = 1.0 / (cfg.rope_theta ** (torch.arange(0, dh, 2, dtype=torch.int64).float() / dh))
    freqs = position_ids[:, :, None].float() * inv_freq[None, None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos(), emb.sin()  # [B, S, dh]


def rotate_half(x):
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def lora_linear(x, w: Dict[str, torch.Tensor], module: str, lora: Optional[Dict[str, torch.Tensor]], scale: float):
    """nn.Linear, or peft.tuners.lora.Linear.forward over it when `lora` holds `<module>.lora_A|lora_B` (peft is not on
    disk; published algorithm): result = base(x) + lora_B(lora_A(dropout(x))) * scaling, dropout off (TRL's
    disable_dropout, base/trainer.py:58)."""
    y = F.linear(x, w[module + ".weight"])
    if lora is not None and module + ".lora_A" in lora:
        y = y + F.linear(F.linear(x, lora[module + ".lora_A"]), lora[module + ".lora_B"]) * scale
    return y


def llama_decoder(cfg: LlavaCfg, w: Dict[str, torch.Tensor], inputs_embeds, attention_mask, position_ids,
                  return_hidden: bool = False, lora: Optional[Dict[str, torch.Tensor]] = None, lora_scale: float = 1.0):
    B, S, _ = inputs_embeds.shape
    lin = lambda x, module: lora_linear(x, w, module, lora, lora_scale)  # noqa: E731
    H, KV, dh = cfg.heads, cfg.kv_heads, cfg.head_dim
    cos, sin = rope_cos_sin(cfg, position_ids)
    cos, sin = cos[:, None], sin[:, None]
    causal = torch.full((S, S), float("-inf")).triu(1)[None, None]
    keypad = torch.zeros(B, 1, 1, S).masked_fill(attention_mask[:, None, None, :] == 0, float("-inf"))
    bias = causal + keypad
    # rows that attend to nothing (cannot happen with right padding: the diagonal is always allowed
    # for ... padded queries? no: padded query rows see earlier valid keys) -- keep HF semantics.
    x = inputs_embeds
    for i in range(cfg.layers):
        p = f"language_model.model.layers.{i}."
        h = rms_norm(x, w[p + "input_layernorm.weight"], cfg.rms_eps)
        q = lin(h, p + "self_attn.q_proj").view(B, S, H, dh).transpose(1, 2)
        k = lin(h, p + "self_attn.k_proj").view(B, S, KV, dh).transpose(1, 2)
        v = lin(h, p + "self_attn.v_proj").view(B, S, KV, dh).transpose(1, 2)
        q = q * cos + rotate_half(q) * sin
        k = k * cos + rotate_half(k) * sin
        if KV != H:
            k = k.repeat_interleave(H // KV, dim=1)
            v = v.repeat_interleave(H // KV, dim=1)
        att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh) + bias, dim=-1)
        o = (att @ v).transpose(1, 2).reshape(B, S, H * dh)
        x = x + lin(o, p + "self_attn.o_proj")
        h = rms_norm(x, w[p + "post_attention_layernorm.weight"], cfg.rms_eps)
        h = F.silu(lin(h, p + "mlp.gate_proj")) * lin(h, p + "mlp.up_proj")
        x = x + lin(h, p + "mlp.down_proj")
    x = rms_norm(x, w["language_model.model.norm.weight"], cfg.rms_eps)
    if return_hidden:
        return x
    return F.linear(x, w["language_model.lm_head.weight"]).float()


# --------------------------------------------------------------------------------------
# models/Llava/__init__.py:111-271  LlavaForRL.forward (training branch: pixel_values given)
# --------------------------------------------------------------------------------------

def llava_forward(cfg: LlavaCfg, w: Dict[str, torch.Tensor], input_ids, attention_mask, labels, pixel_values,
                  lora: Optional[Dict[str, torch.Tensor]] = None, lora_scale: float = 1.0):
    """-> (logits fp32 [2B,S,V], labels [2B,S], image_position_map [2B,S])."""
    inputs_embeds = F.embedding(input_ids, w["language_model.model.embed_tokens.weight"])  # :174
    feats = clip_vision_features(cfg, w, pixel_values)[:, 1:]  # :178-183
    image_features = projector(cfg, w, feats)  # :191
    emb, mask, new_labels, pos, img_map = merge_input_ids_with_image_features(
        cfg, image_features, inputs_embeds, input_ids, attention_mask, labels)  # :192-196
    logits = llama_decoder(cfg, w, emb, mask, pos, lora=lora, lora_scale=lora_scale)  # :232-243
    return logits, new_labels, img_map


# --------------------------------------------------------------------------------------
# LLaVA-Next: transformers-4.41 modeling_llava_next.py helpers (select_best_resolution from
# image_processing_utils.py, get_anyres_image_grid_shape, image_size_to_num_patches, unpad_image,
# pack_image_features) + models/LlavaNext/__init__.py:38-171 merge and :173-345 forward
# --------------------------------------------------------------------------------------

def select_best_resolution(original_size, possible_resolutions):
    """(height, width) of the pinpoint with the largest effective and then smallest wasted resolution."""
    oh, ow = original_size
    best, max_eff, min_waste = None, 0, float("inf")
    for h, w in possible_resolutions:
        scale = min(w / ow, h / oh)
        dw, dh = int(ow * scale), int(oh * scale)
        eff = min(dw * dh, ow * oh)
        waste = w * h - eff
        if eff > max_eff or (eff == max_eff and waste < min_waste):
            max_eff, min_waste, best = eff, waste, (h, w)
    return best


def _as_pair(x):
    return tuple(int(v) for v in (x.tolist() if hasattr(x, "tolist") else x))


def get_anyres_image_grid_shape(image_size, grid_pinpoints, patch_size):
    h, w = select_best_resolution(_as_pair(image_size), [tuple(p) for p in grid_pinpoints])
    return h // patch_size, w // patch_size


def image_size_to_num_patches(image_size, grid_pinpoints, patch_size) -> int:
    """number of crops the processor emitted for this image: the grid cells + the base crop."""
    h, w = select_best_resolution(_as_pair(image_size), [tuple(p) for p in grid_pinpoints])
    return len(range(0, h, patch_size)) * len(range(0, w, patch_size)) + 1


def unpad_image(t: torch.Tensor, original_size) -> torch.Tensor:
    """[C, H, W] feature map of the padded+resized image -> the rows/cols that hold the original aspect."""
    oh, ow = _as_pair(original_size)
    ch, cw = t.shape[1:]
    if ow / oh > cw / ch:
        new_h = int(round(oh * (cw / ow), 7))
        pad = (ch - new_h) // 2
        return t[:, pad:ch - pad, :]
    new_w = int(round(ow * (ch / oh), 7))
    pad = (cw - new_w) // 2
    return t[:, :, pad:cw - pad]


def pack_image_features(cfg: LlavaCfg, image_features: List[torch.Tensor], image_sizes, image_newline: torch.Tensor):
    """"spatial_unpad": base crop features, then the grid crops stitched into one [gh*g, gw*g] map, unpadded,
    one image_newline appended per map row.  -> (concatenated [sum F, d], feature_lens [n_images])."""
    new, lens = [], []
    g = cfg.image_size // cfg.patch_size
    for i, feat in enumerate(image_features):
        if feat.shape[0] > 1:
            base, rest = feat[0], feat[1:]
            gh, gw = get_anyres_image_grid_shape(image_sizes[i], cfg.image_grid_pinpoints, cfg.image_size)
            rest = rest.view(gh, gw, g, g, -1).permute(4, 0, 2, 1, 3).contiguous().flatten(1, 2).flatten(2, 3)
            rest = unpad_image(rest, image_sizes[i])
            rest = torch.cat((rest, image_newline[:, None, None].expand(*rest.shape[:-1], 1).to(rest.dtype)), dim=-1)
            feat = torch.cat((base, rest.flatten(1, 2).transpose(0, 1)), dim=0)
        else:
            feat = torch.cat((feat[0], image_newline[None].to(feat.dtype)), dim=0)
        new.append(feat)
        lens.append(feat.shape[0])
    return torch.cat(new, dim=0), torch.tensor(lens, dtype=torch.long)


def next_merge_input_ids_with_image_features(cfg: LlavaCfg, image_features, feature_lens, inputs_embeds, input_ids,
                                             attention_mask, labels, padding_side: str = "right"):
    """models/LlavaNext/__init__.py:38-171.  Unlike the LLaVA-1.5 merge, tokens with attention_mask == 0 are not
    written at all and the merged length is the longest VALID merged sequence."""
    num_image_features, embed_dim = image_features.shape
    if int(feature_lens.sum()) != num_image_features:
        raise ValueError(f"{feature_lens=} / {feature_lens.sum()} != {image_features.shape=}")
    batch_size = input_ids.shape[0]
    _left = bool(torch.any(attention_mask[:, 0] == 0))
    _right = bool(torch.any(attention_mask[:, -1] == 0))
    left_padding = True
    if batch_size > 1:
        if _left and not _right:
            left_padding = True
        elif not _left and _right:
            left_padding = False
        elif not _left and not _right:
            left_padding = padding_side == "left"
        else:
            raise ValueError(f"both side of attention_mask has zero, invalid. {attention_mask}")
    special = input_ids == cfg.image_token_index
    num_special = torch.sum(special, dim=-1)
    if int(special.sum()) != feature_lens.shape[0]:
        raise ValueError(f"Number of image tokens in input_ids ({int(special.sum())}) different from num_images "
                         f"({feature_lens.shape[0]}).")
    per_seq = torch.tensor([int(x.sum()) for x in feature_lens.split(num_special.tolist(), dim=0)])
    embed_sequence_lengths = (attention_mask == 1).long().sum(-1) - num_special + per_seq
    max_embed_dim = int(embed_sequence_lengths.max())
    batch_indices, non_image_indices = torch.where((input_ids != cfg.image_token_index) & (attention_mask == 1))
    step = special.long()
    step[step == 1] = feature_lens - 1
    new_token_positions = torch.cumsum(step + 1, -1) - 1
    if left_padding:
        new_token_positions = new_token_positions + (max_embed_dim - 1 - new_token_positions[:, -1:])
    text_to_overwrite = new_token_positions[batch_indices, non_image_indices]
    final_embedding = torch.zeros(batch_size, max_embed_dim, embed_dim, dtype=inputs_embeds.dtype)
    final_attention_mask = torch.zeros(batch_size, max_embed_dim, dtype=attention_mask.dtype)
    final_labels = torch.full((batch_size, max_embed_dim), cfg.ignore_index, dtype=torch.long)
    final_embedding[batch_indices, text_to_overwrite] = inputs_embeds[batch_indices, non_image_indices]
    final_attention_mask[batch_indices, text_to_overwrite] = attention_mask[batch_indices, non_image_indices]
    final_labels[batch_indices, text_to_overwrite] = labels[batch_indices, non_image_indices]
    image_to_overwrite = torch.full((batch_size, max_embed_dim), True)
    image_to_overwrite[batch_indices, text_to_overwrite] = False
    idx = torch.arange(max_embed_dim)[None].expand(batch_size, max_embed_dim)
    lens = embed_sequence_lengths[:, None]
    image_to_overwrite &= ((max_embed_dim - idx) <= lens) if left_padding else (idx < lens)
    if int(image_to_overwrite.sum()) != num_image_features:
        raise ValueError(f"{image_to_overwrite.sum()=} != {num_image_features=} The input provided to the model are "
                         "wrong.")
    final_embedding[image_to_overwrite] = image_features.contiguous().reshape(-1, embed_dim)
    final_attention_mask |= image_to_overwrite
    position_ids = (final_attention_mask.cumsum(-1) - 1).masked_fill_((final_attention_mask == 0), 1)
    return final_embedding, final_attention_mask, final_labels, position_ids, image_to_overwrite


def llava_next_forward(cfg: LlavaCfg, w: Dict[str, torch.Tensor], input_ids, attention_mask, labels, pixel_values,
                       image_sizes, lora: Optional[Dict[str, torch.Tensor]] = None, lora_scale: float = 1.0):
    """models/LlavaNext/__init__.py:173-345, training branch -> (logits, labels, image_position_map)."""
    ids0 = input_ids.clone()
    ids0[input_ids == cfg.image_token_index] = 0  # :203-205
    inputs_embeds = F.embedding(ids0, w["language_model.model.embed_tokens.weight"])
    crops = [image_size_to_num_patches(sz, cfg.image_grid_pinpoints, cfg.image_size) for sz in image_sizes]  # :211-218
    if pixel_values.dim() == 5:
        pixel_values = torch.cat([pv[:n] for pv, n in zip(pixel_values, crops)], dim=0)  # :220-225
    elif pixel_values.dim() != 4:
        raise ValueError(f"pixel_values of shape {pixel_values.shape}, expect to be of 4 or 5 dimensions")
    feats = clip_vision_features(cfg, w, pixel_values)[:, 1:]  # :230-236
    image_features = projector(cfg, w, feats)  # :238
    image_features = torch.split(image_features, crops, dim=0)
    image_features, feature_lens = pack_image_features(cfg, list(image_features), image_sizes, w["image_newline"])
    emb, mask, new_labels, pos, img_map = next_merge_input_ids_with_image_features(
        cfg, image_features, feature_lens, inputs_embeds, input_ids, attention_mask, labels)  # :251-262
    logits = llama_decoder(cfg, w, emb, mask, pos, lora=lora, lora_scale=lora_scale)
    return logits, new_labels, img_map


def model_forward(cfg: LlavaCfg, w, input_ids, attention_mask, labels, lora=None, lora_scale: float = 1.0, **img):
    if cfg.family == "llava_next":
        return llava_next_forward(cfg, w, input_ids, attention_mask, labels, img["pixel_values"], img["image_sizes"],
                                  lora=lora, lora_scale=lora_scale)
    return llava_forward(cfg, w, input_ids, attention_mask, labels, img["pixel_values"], lora=lora, lora_scale=lora_scale)


# --------------------------------------------------------------------------------------
# base/trainer.py:190-242 concatenated_forward  +  trl 0.8.1 get_batch_loss_metrics (restated)
# --------------------------------------------------------------------------------------

def concatenated_forward(cfg: LlavaCfg, w, batch, loss_type: str = "sigmoid", label_pad_token_id: int = -100,
                         padding_value: int = 0):
    cb = concatenated_inputs(batch, label_pad_token_id, padding_value)
    n = batch["chosen_labels"].shape[0]
    logits, final_labels, _ = model_forward(cfg, w, cb["concatenated_input_ids"], cb["concatenated_attention_mask"],
                                            cb["concatenated_labels"], **cb["concatenated_img_input_dict"])
    all_logps = get_batch_logps(logits, final_labels, mask_shared_tokens=(loss_type == "ddpo"),
                                label_pad_token_id=label_pad_token_id)
    return all_logps[:n], all_logps[n:], logits[:n], logits[n:]


def get_batch_loss_metrics(cfg: LlavaCfg, w_policy, w_ref, batch, beta: float = 0.1, label_smoothing: float = 0.0,
                           loss_type: str = "sigmoid", reference_free: bool = False, padding_value: int = 0):
    """trl 0.8.1 DPOTrainer.get_batch_loss_metrics: policy pass (grad), no-grad reference pass,
    dpo_loss, reward accuracies / margins, losses.mean()."""
    pc, pr, pcl, prl = concatenated_forward(cfg, w_policy, batch, loss_type, padding_value=padding_value)
    with torch.no_grad():
        rc, rr, _, _ = concatenated_forward(cfg, w_ref, batch, loss_type, padding_value=padding_value)
    losses, cr, rj = dpo_loss(pc, pr, rc, rr, beta, label_smoothing, loss_type, reference_free)
    acc = (cr > rj).float()
    metrics = {
        "rewards/chosen": cr.mean(), "rewards/rejected": rj.mean(), "rewards/accuracies": acc.mean(),
        "rewards/margins": (cr - rj).mean(), "logps/rejected": pr.detach().mean(), "logps/chosen": pc.detach().mean(),
        "logits/rejected": prl.detach().mean(), "logits/chosen": pcl.detach().mean(),
    }
    return losses.mean(), metrics, dict(policy_chosen_logps=pc, policy_rejected_logps=pr, reference_chosen_logps=rc,
                                        reference_rejected_logps=rr, losses=losses, chosen_rewards=cr,
                                        rejected_rewards=rj)


def perturbed_reference_names(cfg: LlavaCfg) -> List[str]:
    return [n for n, *_ in weight_specs(cfg) if not n.startswith("vision_tower.")]


def make_policy_and_ref(cfg: LlavaCfg, seed: int):
    """Policy = seeded weights; reference = same vision tower, projector+LLM from seed+1 blended:
    w_ref = bf16(w + 0.05 * w') so margins are non-zero (SURVEY §8c iv) yet ref stays close."""
    wp = make_weights(cfg, seed)
    other = make_weights(cfg, seed + 1, perturbed_reference_names(cfg))
    wr = dict(wp)
    for n, t in other.items():
        wr[n] = bf16_round(wp[n] + 0.05 * (t - (1.0 if n.endswith("norm.weight") else 0.0)))
    return wp, wr
