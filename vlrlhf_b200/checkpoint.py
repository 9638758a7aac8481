"""HF checkpoint interchange for the flat arenas.

The reference loads the policy with `MyAutoModel.from_pretrained(model_name_or_path, config=config, torch_dtype=…)`
(utils/auto_load.py:522-535) and saves with HF `save_pretrained` / `safe_save_model_for_hf_trainer`
(dpo.py:89-95,147-149, utils/common.py:21-55) so that the eval harness and merge_peft_model.py can reload the result.
Here the same directory layout is read and written directly from/to the engine's storage:

  * `config.json` -> `ModelConfig` (`config_from_hf`), LLaVA-1.5 (`model_type: llava`) and LLaVA-Next (`llava_next`);
  * `model.safetensors` or `model.safetensors.index.json` + shards, streamed one tensor at a time into the HF-named
    views of the arenas (policy AND the frozen reference copy: TRL deep-copies the initial policy as the reference);
    both the transformers-4.41 names (`language_model.model.…`, what the reference's pinned version writes) and the
    5.x names (`model.language_model.…`) are accepted;
  * tensors the path never reads (CLIP layers above `vision_feature_layer`, `post_layernorm`) are kept on the host and
    written back unchanged, so a saved checkpoint is complete.
`save_hf_checkpoint` writes 4.41 names (what the reference's environment expects) in <= 5 GB shards + the index.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Iterator, Optional, Tuple

import torch

from .config import ModelConfig


def legacy_name(name: str) -> str:
    """transformers-5.x parameter name -> the 4.41 name the reference's checkpoints use (identity for 4.41 names)."""
    if name == "lm_head.weight":
        return "language_model.lm_head.weight"
    if name.startswith("model.language_model.model."):  # what transformers 5.x writes to disk (partial reverse mapping)
        return "language_model.model." + name[len("model.language_model.model."):]
    if name.startswith("model.language_model."):        # transformers 5.x in-memory names
        return "language_model.model." + name[len("model.language_model."):]
    for head in ("vision_tower.", "multi_modal_projector.", "image_newline"):
        if name.startswith("model." + head):
            return name[len("model."):]
    return name


def _get(obj, key, default=None):
    if isinstance(obj, dict):
        return obj.get(key, default)
    return getattr(obj, key, default)


def config_from_hf(hf_config) -> ModelConfig:
    """HF LlavaConfig / LlavaNextConfig (object or parsed config.json dict) -> ModelConfig."""
    mt = _get(hf_config, "model_type")
    if mt not in ("llava", "llava_next"):
        raise ValueError(f"model_type {mt!r}: only llava and llava_next checkpoints map onto this engine")
    t, v = _get(hf_config, "text_config"), _get(hf_config, "vision_config")
    if t is None or v is None:
        raise ValueError("config without text_config / vision_config")
    heads = _get(t, "num_attention_heads", 32)
    hidden = _get(t, "hidden_size", 4096)
    head_dim = _get(t, "head_dim") or hidden // heads
    if head_dim * heads != hidden:
        raise ValueError("head_dim * num_attention_heads != hidden_size is not supported")
    rope = _get(t, "rope_theta")
    if rope is None:
        rp = _get(t, "rope_parameters") or {}
        rope = _get(rp, "rope_theta", 10000.0)
    if _get(t, "sliding_window") not in (None, 0) and _get(t, "sliding_window") < _get(t, "max_position_embeddings", 1 << 30):
        raise ValueError("sliding-window attention is not implemented")
    pins = _get(hf_config, "image_grid_pinpoints") or ()
    tok = _get(hf_config, "image_token_index", _get(hf_config, "image_token_id", 32000))
    pad = _get(hf_config, "pad_token_id")
    if pad is None:
        pad = _get(t, "pad_token_id")
    return ModelConfig(
        image_size=_get(v, "image_size", 336), patch_size=_get(v, "patch_size", 14), v_hidden=_get(v, "hidden_size", 1024),
        v_layers=_get(v, "num_hidden_layers", 24), v_heads=_get(v, "num_attention_heads", 16),
        v_ff=_get(v, "intermediate_size", 4096), v_eps=_get(v, "layer_norm_eps", 1e-5),
        vision_feature_layer=_get(hf_config, "vision_feature_layer", -2),
        hidden=hidden, layers=_get(t, "num_hidden_layers", 32), heads=heads,
        kv_heads=_get(t, "num_key_value_heads", heads) or heads, ff=_get(t, "intermediate_size", 11008),
        vocab=_get(t, "vocab_size", 32064), rms_eps=_get(t, "rms_norm_eps", 1e-6), rope_theta=float(rope),
        image_token_index=int(tok), pad_token_id=int(pad) if pad is not None else -1,
        max_positions=max(4096, min(int(_get(t, "max_position_embeddings", 4096)), 8192)),
        family="llava_next" if mt == "llava_next" else "llava",
        image_grid_pinpoints=tuple(tuple(int(x) for x in p) for p in pins))


def hf_config_dict(cfg: ModelConfig) -> dict:
    """Inverse of `config_from_hf`: a config.json body (transformers-4.41 field names) for a ModelConfig."""
    text = dict(model_type="mistral" if cfg.family == "llava_next" and cfg.kv_heads != cfg.heads else "llama",
                hidden_size=cfg.hidden, intermediate_size=cfg.ff, num_hidden_layers=cfg.layers,
                num_attention_heads=cfg.heads, num_key_value_heads=cfg.kv_heads, vocab_size=cfg.vocab,
                rms_norm_eps=cfg.rms_eps, rope_theta=cfg.rope_theta, max_position_embeddings=cfg.max_positions,
                sliding_window=None, tie_word_embeddings=False, torch_dtype="bfloat16")
    vision = dict(model_type="clip_vision_model", hidden_size=cfg.v_hidden, intermediate_size=cfg.v_ff,
                  num_hidden_layers=cfg.v_layers, num_attention_heads=cfg.v_heads, image_size=cfg.image_size,
                  patch_size=cfg.patch_size, layer_norm_eps=cfg.v_eps, hidden_act="quick_gelu",
                  projection_dim=cfg.v_hidden)
    out = dict(model_type=cfg.family if cfg.family == "llava_next" else "llava",
               architectures=["LlavaNextForConditionalGeneration" if cfg.family == "llava_next" else
                              "LlavaForConditionalGeneration"],
               text_config=text, vision_config=vision, image_token_index=cfg.image_token_index, pad_token_id=cfg.pad_token_id,
               ignore_index=cfg.ignore_index, projector_hidden_act="gelu", vision_feature_layer=cfg.vision_feature_layer,
               vision_feature_select_strategy="default", tie_word_embeddings=False, torch_dtype="bfloat16")
    if cfg.family == "llava_next":
        out["image_grid_pinpoints"] = [list(p) for p in cfg.image_grid_pinpoints]
    return out


def iter_checkpoint(path: str) -> Iterator[Tuple[str, torch.Tensor]]:
    """Yield (name, CPU tensor) for every tensor of an HF safetensors checkpoint directory, one at a time."""
    from safetensors import safe_open
    index = os.path.join(path, "model.safetensors.index.json")
    if os.path.exists(index):
        with open(index) as f:
            files = sorted(set(json.load(f)["weight_map"].values()))
    elif os.path.exists(os.path.join(path, "model.safetensors")):
        files = ["model.safetensors"]
    else:
        raise FileNotFoundError(f"no model.safetensors[.index.json] under {path} (torch .bin checkpoints are not read)")
    for fn in files:
        with safe_open(os.path.join(path, fn), framework="pt", device="cpu") as f:
            for k in f.keys():
                yield k, f.get_tensor(k)


def load_hf_checkpoint(engine, path: str, strict: bool = True) -> Dict[str, torch.Tensor]:
    """Stream a checkpoint into the engine (policy + reference copy + vision tower).  Returns the tensors the path
    does not use (kept on the host for `save_hf_checkpoint`)."""
    pol, ref = engine.hf_state("policy"), engine.hf_state("ref")
    seen, extra = set(), {}
    for raw, t in iter_checkpoint(path):
        k = legacy_name(raw)
        if k in pol:
            if tuple(pol[k].shape) != tuple(t.shape) and pol[k].numel() != t.numel():
                raise ValueError(f"{raw}: checkpoint shape {tuple(t.shape)} vs model {tuple(pol[k].shape)}")
            src = t.to(device=engine.device, dtype=torch.bfloat16).reshape(pol[k].shape)
            pol[k].copy_(src)
            if k in ref and not k.startswith("vision_tower."):
                ref[k].copy_(src)
            seen.add(k)
        else:
            extra[k] = t
    # adapter tensors of a LoRA engine are not part of a base checkpoint (peft keeps them in adapter_model.safetensors)
    missing = [k for k in pol if k not in seen and not k.endswith((".lora_A", ".lora_B"))]
    if strict and missing:
        raise KeyError(f"checkpoint lacks {len(missing)} tensors, e.g. {missing[:4]}")
    engine.extra_state = extra
    engine.sync_master_from_params()
    return extra


def save_hf_checkpoint(engine, path: str, hf_config_dict: Optional[dict] = None, max_shard_bytes: int = 5 << 30,
                       which: str = "policy"):
    """Write `which` ("policy" | "ref") as an HF safetensors checkpoint with transformers-4.41 names."""
    from safetensors.torch import save_file
    os.makedirs(path, exist_ok=True)
    from .config import weight_specs
    hf_shape = {name: shape for name, shape, _, _ in weight_specs(engine.cfg)}
    # engine views may be fused / padded storage (the CLIP patch embedding is a [dv, 3*p*p] view): restore HF shapes
    state = {k: (v.reshape(hf_shape[k]) if k in hf_shape and tuple(v.shape) != tuple(hf_shape[k]) else v)
             for k, v in engine.hf_state(which).items()}
    state.update(getattr(engine, "extra_state", {}) or {})
    shards, cur, cur_bytes = [], {}, 0
    for k in sorted(state):
        t = state[k]
        nbytes = t.numel() * t.element_size()
        if cur and cur_bytes + nbytes > max_shard_bytes:
            shards.append(cur)
            cur, cur_bytes = {}, 0
        cur[k] = t
        cur_bytes += nbytes
    if cur:
        shards.append(cur)
    weight_map, total = {}, 0
    for i, shard in enumerate(shards):
        fn = "model.safetensors" if len(shards) == 1 else f"model-{i + 1:05d}-of-{len(shards):05d}.safetensors"
        # one shard at a time on the host (a 7B policy is 13.5 GB in bf16)
        host = {k: v.detach().to("cpu").contiguous() for k, v in shard.items()}
        save_file(host, os.path.join(path, fn), metadata={"format": "pt"})
        for k, v in host.items():
            weight_map[k] = fn
            total += v.numel() * v.element_size()
        del host
    if len(shards) > 1:
        with open(os.path.join(path, "model.safetensors.index.json"), "w") as f:
            json.dump({"metadata": {"total_size": total}, "weight_map": weight_map}, f, indent=2)
    if hf_config_dict is not None:
        with open(os.path.join(path, "config.json"), "w") as f:
            json.dump(hf_config_dict, f, indent=2)
    return sorted(set(weight_map.values()))
