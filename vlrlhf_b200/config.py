"""Model / training configuration of the hot path (LLaVA-1.5 family).

Parameter names follow transformers-4.41 `LlavaForConditionalGeneration` (what the reference's
`LlavaForRL` subclasses, models/Llava/__init__.py:35) so state dicts map 1:1.
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass
from typing import List, Tuple


@dataclass
class ModelConfig:
    # vision tower (CLIP ViT)
    image_size: int = 336
    patch_size: int = 14
    v_hidden: int = 1024
    v_layers: int = 24
    v_heads: int = 16
    v_ff: int = 4096
    v_eps: float = 1e-5
    vision_feature_layer: int = -2
    # decoder (Llama / Mistral style)
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
    max_positions: int = 4096
    # "llava" (models/Llava/__init__.py) | "llava_next" (models/LlavaNext/__init__.py: anyres crops + image_newline)
    family: str = "llava"
    image_grid_pinpoints: Tuple[Tuple[int, int], ...] = ()

    @property
    def n_patches(self) -> int:
        return (self.image_size // self.patch_size) ** 2

    @property
    def v_used_layers(self) -> int:
        n = self.v_layers
        return self.vision_feature_layer if self.vision_feature_layer >= 0 else n + 1 + self.vision_feature_layer

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads

    @property
    def v_head_dim(self) -> int:
        return self.v_hidden // self.v_heads

    @property
    def qkv_dim(self) -> int:
        return (self.heads + 2 * self.kv_heads) * self.head_dim

    @property
    def patch_k(self) -> int:
        return 3 * self.patch_size * self.patch_size

    @property
    def patch_k_padded(self) -> int:
        return (self.patch_k + 7) // 8 * 8


LLAVA15_7B = ModelConfig()
TINY = ModelConfig(image_size=28, patch_size=14, v_hidden=128, v_layers=3, v_heads=2, v_ff=256, hidden=128, layers=2,
                   heads=2, kv_heads=2, ff=256, vocab=320, image_token_index=300, pad_token_id=301)
# exercises the production tile shapes (head dims 64 / 128, several tiles per GEMM)
SMALL = ModelConfig(image_size=112, patch_size=14, v_hidden=256, v_layers=3, v_heads=4, v_ff=512, hidden=512, layers=2,
                    heads=4, kv_heads=4, ff=1024, vocab=2048, image_token_index=2000, pad_token_id=2001)


# llava-v1.6-mistral-7b-hf: CLIP-L/336 tower, Mistral-7B-Instruct-v0.2 decoder (GQA 32/8, ff 14336, theta 1e6, no
# sliding window), anyres pinpoints (BASELINE.json configs[3])
LLAVA_NEXT_PINPOINTS = ((336, 672), (672, 336), (672, 672), (1008, 336), (336, 1008))
LLAVANEXT_MISTRAL_7B = ModelConfig(hidden=4096, layers=32, heads=32, kv_heads=8, ff=14336, vocab=32064, rope_theta=1e6,
                                   family="llava_next", image_grid_pinpoints=LLAVA_NEXT_PINPOINTS, max_positions=8192)
TINY_NEXT = ModelConfig(image_size=28, patch_size=14, v_hidden=128, v_layers=3, v_heads=2, v_ff=256, hidden=256, layers=2,
                        heads=4, kv_heads=2, ff=256, vocab=320, rope_theta=1e6, image_token_index=300, pad_token_id=301,
                        family="llava_next",
                        image_grid_pinpoints=((28, 56), (56, 28), (56, 56), (84, 28), (28, 84)))
SMALL_NEXT = ModelConfig(image_size=112, patch_size=14, v_hidden=256, v_layers=3, v_heads=4, v_ff=512, hidden=512,
                         layers=2, heads=4, kv_heads=2, ff=1024, vocab=2048, rope_theta=1e6, image_token_index=2000,
                         pad_token_id=2001, family="llava_next",
                         image_grid_pinpoints=((112, 224), (224, 112), (224, 224), (336, 112), (112, 336)))


@dataclass
class TrainConfig:
    """Hot-path knobs of the reference's TrainingArguments / VLDPOTrainer (dpo.py:16-86, base/trainer.py:38-67)."""
    beta: float = 0.1
    label_smoothing: float = 0.0
    loss_type: str = "sigmoid"  # sigmoid | hinge | ipo | kto_pair | ddpo
    reference_free: bool = False
    label_pad_token_id: int = -100
    padding_value: int = 0
    # optimizer (scripts/dpo_llava.sh:35-42; HF Trainer defaults)
    learning_rate: float = 1e-6
    adam_beta1: float = 0.9
    adam_beta2: float = 0.98
    adam_eps: float = 1e-6
    weight_decay: float = 0.0
    max_grad_norm: float = 1.0
    # HF TrainingArguments.gradient_checkpointing (dpo.py:21 / scripts/*.sh --gradient_checkpointing True): keep only
    # each decoder layer's fp32 input, recompute the layer (minus down_proj) in backward
    activation_checkpointing: bool = False
    # drop the collator's padding rows (base/collator.py:44-60) from the merged batch -- sequences are stacked
    # back to back, attention runs var-len -- so padded positions cost nothing.  Log-probs, losses, rewards and gradients
    # are those of the padded batch; only TRL's `logits/chosen|rejected` metric changes meaning (it averages over the
    # attended positions instead of all positions).  All engines (LLaVA-1.5 / LLaVA-Next full fine-tune and LoRA, Qwen-VL, XC2).
    pack_sequences: bool = False
    # the reference concatenates chosen and rejected on the batch axis (base/trainer.py:124-146), so the
    # prompt + image prefix the two sequences of a pair have in common (identical tokens, identical image, causal attention
    # => identical hidden states) is computed twice per pass.  With share_prefix it is laid out, projected, normed and
    # attended ONCE per pair: rows = [chosen suffixes | prefixes | rejected suffixes], the attention kernels treat a suffix's
    # prefix as context keys (vlb200_attn_*_tc_ctx), the backward sums both suffixes' gradients into the prefix.  Implies
    # packed rows.  Results equal the padded step up to accumulation order (the prefix gradient is accumulated inside one
    # attention backward instead of two).  LLaVA-1.5 family (full fine-tune and LoRA).
    share_prefix: bool = False
    # HF TrainingArguments the reference recipes set (scripts/dpo_qwenvl.sh:4,35,42-43; scripts/dpo_llava.sh:40-41):
    # k micro-batches are accumulated in the gradient arena before ONE reduction + AdamW step; the learning rate follows
    # transformers.get_scheduler ("constant" | "linear" | "cosine", linear warm-up over warmup_steps or
    # ceil(warmup_ratio * max_steps) optimizer steps; max_steps = 0: constant after the warm-up)
    gradient_accumulation_steps: int = 1
    lr_scheduler_type: str = "constant"
    warmup_steps: int = 0
    warmup_ratio: float = 0.0
    max_steps: int = 0

    def lr_at(self, optimizer_step: int) -> float:
        """Learning rate of the optimizer step with this 0-based index == what transformers' LambdaLR schedules
        (optimization.py: get_constant_schedule_with_warmup / get_linear_schedule_with_warmup /
        get_cosine_schedule_with_warmup, num_cycles 0.5) hand to torch.optim.AdamW at that step."""
        import math as _m
        t = int(optimizer_step)
        warm = self.warmup_steps if self.warmup_steps > 0 else int(_m.ceil(self.warmup_ratio * self.max_steps))
        if t < warm:
            return self.learning_rate * t / max(1, warm)
        kind = self.lr_scheduler_type
        if kind in ("constant", "constant_with_warmup") or self.max_steps <= 0:
            return self.learning_rate
        if kind == "linear":
            return self.learning_rate * max(0.0, (self.max_steps - t) / max(1, self.max_steps - warm))
        if kind == "cosine":
            prog = (t - warm) / max(1, self.max_steps - warm)
            return self.learning_rate * max(0.0, 0.5 * (1.0 + _m.cos(_m.pi * 2.0 * 0.5 * prog)))
        raise ValueError(f"lr_scheduler_type {kind!r}: constant, linear or cosine")


def tensor_seed(name: str, base_seed: int) -> int:
    return (zlib.crc32(name.encode()) ^ (base_seed * 0x9E3779B1)) & 0xFFFFFFFF


def weight_specs(cfg: ModelConfig) -> List[Tuple[str, Tuple[int, ...], float, float]]:
    """(HF-4.41 name, shape, uniform half-width, shift) -- the synthetic-init recipe (random-init weights of
    the named architecture; there are no checkpoints on the box)."""
    a = 0.02 * math.sqrt(3.0)
    s: List[Tuple[str, Tuple[int, ...], float, float]] = []
    vp = "vision_tower.vision_model."
    s += [(vp + "embeddings.class_embedding", (cfg.v_hidden,), a, 0.0),
          (vp + "embeddings.patch_embedding.weight", (cfg.v_hidden, 3, cfg.patch_size, cfg.patch_size), a, 0.0),
          (vp + "embeddings.position_embedding.weight", (cfg.n_patches + 1, cfg.v_hidden), a, 0.0),
          (vp + "pre_layrnorm.weight", (cfg.v_hidden,), 0.1, 1.0),
          (vp + "pre_layrnorm.bias", (cfg.v_hidden,), 0.02, 0.0)]
    for i in range(cfg.v_layers):
        p = f"{vp}encoder.layers.{i}."
        for ln in ("layer_norm1", "layer_norm2"):
            s += [(p + ln + ".weight", (cfg.v_hidden,), 0.1, 1.0), (p + ln + ".bias", (cfg.v_hidden,), 0.02, 0.0)]
        for pr in ("q_proj", "k_proj", "v_proj", "out_proj"):
            s += [(p + f"self_attn.{pr}.weight", (cfg.v_hidden, cfg.v_hidden), a, 0.0),
                  (p + f"self_attn.{pr}.bias", (cfg.v_hidden,), 0.02, 0.0)]
        s += [(p + "mlp.fc1.weight", (cfg.v_ff, cfg.v_hidden), a, 0.0), (p + "mlp.fc1.bias", (cfg.v_ff,), 0.02, 0.0),
              (p + "mlp.fc2.weight", (cfg.v_hidden, cfg.v_ff), a, 0.0), (p + "mlp.fc2.bias", (cfg.v_hidden,), 0.02, 0.0)]
    s += [("multi_modal_projector.linear_1.weight", (cfg.hidden, cfg.v_hidden), a, 0.0),
          ("multi_modal_projector.linear_1.bias", (cfg.hidden,), 0.02, 0.0),
          ("multi_modal_projector.linear_2.weight", (cfg.hidden, cfg.hidden), a, 0.0),
          ("multi_modal_projector.linear_2.bias", (cfg.hidden,), 0.02, 0.0),
          ("language_model.model.embed_tokens.weight", (cfg.vocab, cfg.hidden), a, 0.0)]
    if cfg.family == "llava_next":
        s.append(("image_newline", (cfg.hidden,), a, 0.0))
    kv = cfg.kv_heads * cfg.head_dim
    for i in range(cfg.layers):
        p = f"language_model.model.layers.{i}."
        s += [(p + "input_layernorm.weight", (cfg.hidden,), 0.1, 1.0),
              (p + "self_attn.q_proj.weight", (cfg.heads * cfg.head_dim, cfg.hidden), a, 0.0),
              (p + "self_attn.k_proj.weight", (kv, cfg.hidden), a, 0.0),
              (p + "self_attn.v_proj.weight", (kv, cfg.hidden), a, 0.0),
              (p + "self_attn.o_proj.weight", (cfg.hidden, cfg.heads * cfg.head_dim), a, 0.0),
              (p + "post_attention_layernorm.weight", (cfg.hidden,), 0.1, 1.0),
              (p + "mlp.gate_proj.weight", (cfg.ff, cfg.hidden), a, 0.0),
              (p + "mlp.up_proj.weight", (cfg.ff, cfg.hidden), a, 0.0),
              (p + "mlp.down_proj.weight", (cfg.hidden, cfg.ff), a, 0.0)]
    s += [("language_model.model.norm.weight", (cfg.hidden,), 0.1, 1.0),
          ("language_model.lm_head.weight", (cfg.vocab, cfg.hidden), 3.0 * a, 0.0)]
    return s


# ------------------------------------------------------------------------------------------------
# Qwen-VL (BASELINE.json configs[2]): open_clip ViT-bigG/14 + resampler, Qwen-7B decoder, LoRA on the
# LM (scripts/dpo_qwenvl.sh: r 64, alpha 16, c_attn / attn.c_proj / w1 / w2), vision tower frozen.
# Parameter names follow the reference's vendored model (models/QwenVL/modeling_qwen.py, visual.py).
# ------------------------------------------------------------------------------------------------
@dataclass
class QwenModelConfig:
    image_size: int = 448
    patch_size: int = 14
    v_width: int = 1664
    v_layers: int = 48
    v_heads: int = 16
    v_mlp: int = 8192
    n_queries: int = 256
    v_eps: float = 1e-6
    pos_table: int = 256
    hidden: int = 4096
    layers: int = 32
    heads: int = 32
    ff: int = 11008
    vocab: int = 151936
    rms_eps: float = 1e-6
    rope_theta: float = 10000.0
    image_start_id: int = 151857
    pad_token_id: int = 151643
    ignore_index: int = -100
    lora_r: int = 64
    lora_alpha: float = 16.0
    max_positions: int = 4096
    family: str = "qwen_vl"

    @property
    def n_patches(self) -> int:
        return (self.image_size // self.patch_size) ** 2

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads

    @property
    def kv_heads(self) -> int:
        return self.heads

    @property
    def qkv_dim(self) -> int:
        return 3 * self.hidden

    @property
    def v_head_dim(self) -> int:
        return self.v_width // self.v_heads

    @property
    def v_head_pad(self) -> int:
        """ViT heads are laid out at the next supported attention head size (104 -> 128): zero-padded q/k/v dims leave
        q.k and the context unchanged."""
        return 64 if self.v_head_dim <= 64 else 128

    @property
    def r_heads(self) -> int:
        return self.hidden // 128

    @property
    def lora_scale(self) -> float:
        return self.lora_alpha / self.lora_r

    @property
    def patch_k(self) -> int:
        return 3 * self.patch_size * self.patch_size

    @property
    def patch_k_padded(self) -> int:
        return (self.patch_k + 7) // 8 * 8

    @property
    def image_token_index(self) -> int:  # the token DDPO's merged-label layout treats as "image start" (no expansion here)
        return -1


QWEN_VL_CHAT = QwenModelConfig()
TINY_QWEN = QwenModelConfig(image_size=112, patch_size=14, v_width=208, v_layers=2, v_heads=2, v_mlp=416, n_queries=16,
                            hidden=256, layers=2, heads=2, ff=256, vocab=512, image_start_id=500, pad_token_id=499,
                            lora_r=16, lora_alpha=8.0)
SMALL_QWEN = QwenModelConfig(image_size=224, patch_size=14, v_width=416, v_layers=2, v_heads=4, v_mlp=1024, n_queries=64,
                             hidden=512, layers=2, heads=4, ff=1024, vocab=2048, image_start_id=2000, pad_token_id=1999,
                             lora_r=16, lora_alpha=8.0)
QWEN_LORA_TARGETS = ("attn.c_attn", "attn.c_proj", "mlp.w1", "mlp.w2")


def qwen_weight_specs(cfg: QwenModelConfig) -> List[Tuple[str, Tuple[int, ...], float, float]]:
    """(reference state-dict name, shape, uniform half-width, shift) of the frozen base model (synthetic-init recipe)."""
    a = 0.02 * math.sqrt(3.0)
    d, w = cfg.hidden, cfg.v_width
    s: List[Tuple[str, Tuple[int, ...], float, float]] = []
    v = "transformer.visual."
    s += [(v + "positional_embedding", (cfg.pos_table, w), a, 0.0), (v + "proj", (d, d), a, 0.0),
          (v + "conv1.weight", (w, 3, cfg.patch_size, cfg.patch_size), a, 0.0),
          (v + "ln_pre.weight", (w,), 0.1, 1.0), (v + "ln_pre.bias", (w,), 0.02, 0.0)]
    for i in range(cfg.v_layers):
        p = f"{v}transformer.resblocks.{i}."
        s += [(p + "ln_1.weight", (w,), 0.1, 1.0), (p + "ln_1.bias", (w,), 0.02, 0.0),
              (p + "ln_2.weight", (w,), 0.1, 1.0), (p + "ln_2.bias", (w,), 0.02, 0.0),
              (p + "attn.in_proj.weight", (3 * w, w), a, 0.0), (p + "attn.in_proj.bias", (3 * w,), 0.02, 0.0),
              (p + "attn.out_proj.weight", (w, w), a, 0.0), (p + "attn.out_proj.bias", (w,), 0.02, 0.0),
              (p + "mlp.c_fc.weight", (cfg.v_mlp, w), a, 0.0), (p + "mlp.c_fc.bias", (cfg.v_mlp,), 0.02, 0.0),
              (p + "mlp.c_proj.weight", (w, cfg.v_mlp), a, 0.0), (p + "mlp.c_proj.bias", (w,), 0.02, 0.0)]
    p = v + "attn_pool."
    s += [(p + "query", (cfg.n_queries, d), a, 0.0), (p + "kv_proj.weight", (d, w), a, 0.0),
          (p + "attn.in_proj_weight", (3 * d, d), a, 0.0), (p + "attn.in_proj_bias", (3 * d,), 0.02, 0.0),
          (p + "attn.out_proj.weight", (d, d), a, 0.0), (p + "attn.out_proj.bias", (d,), 0.02, 0.0),
          (p + "ln_q.weight", (d,), 0.1, 1.0), (p + "ln_q.bias", (d,), 0.02, 0.0),
          (p + "ln_kv.weight", (d,), 0.1, 1.0), (p + "ln_kv.bias", (d,), 0.02, 0.0)]
    s += [(v + "ln_post.weight", (d,), 0.1, 1.0), (v + "ln_post.bias", (d,), 0.02, 0.0)]
    s += [("transformer.wte.weight", (cfg.vocab, d), a, 0.0)]
    for i in range(cfg.layers):
        p = f"transformer.h.{i}."
        s += [(p + "ln_1.weight", (d,), 0.1, 1.0),
              (p + "attn.c_attn.weight", (3 * d, d), a, 0.0), (p + "attn.c_attn.bias", (3 * d,), 0.02, 0.0),
              (p + "attn.c_proj.weight", (d, d), a, 0.0),
              (p + "ln_2.weight", (d,), 0.1, 1.0),
              (p + "mlp.w1.weight", (cfg.ff, d), a, 0.0), (p + "mlp.w2.weight", (cfg.ff, d), a, 0.0),
              (p + "mlp.c_proj.weight", (d, cfg.ff), a, 0.0)]
    s += [("transformer.ln_f.weight", (d,), 0.1, 1.0), ("lm_head.weight", (cfg.vocab, d), 3.0 * a, 0.0)]
    return s


def qwen_lora_specs(cfg: QwenModelConfig) -> List[Tuple[str, Tuple[int, ...], float, float]]:
    """`<module>.lora_A` [r, in] / `.lora_B` [out, r] per target module (peft layout).  peft starts B at 0; the
    synthetic recipe draws B too so the adapter path carries signal in parity tests."""
    a = 0.02 * math.sqrt(3.0)
    d, r = cfg.hidden, cfg.lora_r
    outs = {"attn.c_attn": 3 * d, "attn.c_proj": d, "mlp.w1": cfg.ff, "mlp.w2": cfg.ff}
    s = []
    for i in range(cfg.layers):
        for t in QWEN_LORA_TARGETS:
            s += [(f"transformer.h.{i}.{t}.lora_A", (r, d), a, 0.0), (f"transformer.h.{i}.{t}.lora_B", (outs[t], r), a, 0.0)]
    return s


# ------------------------------------------------------------------------------------------------
# InternLM-XComposer2-VL (BASELINE.json configs[4]): CLIP-L/14 at 490 px (35x35 patches, last layer),
# Linear-GELU-Linear projector, InternLM2-7B (GQA 32/8) whose every linear carries a frozen partial-LoRA (r 256) on the
# image rows; training adapts peft LoRA r 64 on attention.wqkv/wo and feed_forward.w1/w2/w3 (scripts/dpo_internlmxc2vl7b.sh).
# ------------------------------------------------------------------------------------------------
@dataclass
class XC2ModelConfig(ModelConfig):
    plora_r: int = 256
    plora_alpha: float = 256.0
    lora_r: int = 64
    lora_alpha: float = 64.0

    @property
    def plora_scale(self) -> float:
        return self.plora_alpha / self.plora_r

    @property
    def lora_scale(self) -> float:
        return self.lora_alpha / self.lora_r


XC2_VL_7B = XC2ModelConfig(image_size=490, vision_feature_layer=-1, hidden=4096, layers=32, heads=32, kv_heads=8, ff=14336,
                           vocab=92544, rms_eps=1e-5, rope_theta=1e6, image_token_index=92543, pad_token_id=2, family="xc2")
TINY_XC2 = XC2ModelConfig(image_size=70, patch_size=14, v_hidden=128, v_layers=2, v_heads=2, v_ff=256, vision_feature_layer=-1,
                          hidden=256, layers=2, heads=4, kv_heads=2, ff=512, vocab=512, rms_eps=1e-5, rope_theta=1e6,
                          image_token_index=500, pad_token_id=2, family="xc2", plora_r=32, plora_alpha=32.0, lora_r=16,
                          lora_alpha=16.0)
SMALL_XC2 = XC2ModelConfig(image_size=112, patch_size=14, v_hidden=256, v_layers=2, v_heads=4, v_ff=512,
                           vision_feature_layer=-1, hidden=512, layers=2, heads=4, kv_heads=2, ff=1024, vocab=2048,
                           rms_eps=1e-5, rope_theta=1e6, image_token_index=2000, pad_token_id=2, family="xc2", plora_r=64,
                           plora_alpha=64.0, lora_r=16, lora_alpha=16.0)
XC2_LINEARS = ("attention.wqkv", "attention.wo", "feed_forward.w1", "feed_forward.w3", "feed_forward.w2")


def xc2_linear_dims(cfg: XC2ModelConfig):
    d, dh = cfg.hidden, cfg.head_dim
    return {"attention.wqkv": ((cfg.heads + 2 * cfg.kv_heads) * dh, d), "attention.wo": (d, cfg.heads * dh),
            "feed_forward.w1": (cfg.ff, d), "feed_forward.w3": (cfg.ff, d), "feed_forward.w2": (d, cfg.ff)}


def xc2_weight_specs(cfg: XC2ModelConfig) -> List[Tuple[str, Tuple[int, ...], float, float]]:
    a = 0.02 * math.sqrt(3.0)
    s: List[Tuple[str, Tuple[int, ...], float, float]] = []
    for name, shape, scale, shift in weight_specs(cfg):
        if name.startswith("vision_tower.vision_model."):
            s.append(("vit." + name, shape, scale, shift))
    s += [("vision_proj.0.weight", (cfg.hidden, cfg.v_hidden), a, 0.0), ("vision_proj.0.bias", (cfg.hidden,), 0.02, 0.0),
          ("vision_proj.2.weight", (cfg.hidden, cfg.hidden), a, 0.0), ("vision_proj.2.bias", (cfg.hidden,), 0.02, 0.0),
          ("model.tok_embeddings.weight", (cfg.vocab, cfg.hidden), a, 0.0)]
    dims = xc2_linear_dims(cfg)
    for i in range(cfg.layers):
        p = f"model.layers.{i}."
        s += [(p + "attention_norm.weight", (cfg.hidden,), 0.1, 1.0), (p + "ffn_norm.weight", (cfg.hidden,), 0.1, 1.0)]
        for lin in XC2_LINEARS:
            out, inn = dims[lin]
            s += [(p + lin + ".weight", (out, inn), a, 0.0), (p + lin + ".Plora_A.weight", (cfg.plora_r, inn), a, 0.0),
                  (p + lin + ".Plora_B.weight", (out, cfg.plora_r), a, 0.0)]
    s += [("model.norm.weight", (cfg.hidden,), 0.1, 1.0), ("output.weight", (cfg.vocab, cfg.hidden), 3.0 * a, 0.0)]
    return s


def xc2_lora_specs(cfg: XC2ModelConfig) -> List[Tuple[str, Tuple[int, ...], float, float]]:
    a = 0.02 * math.sqrt(3.0)
    dims = xc2_linear_dims(cfg)
    s = []
    for i in range(cfg.layers):
        for lin in XC2_LINEARS:
            out, inn = dims[lin]
            s += [(f"model.layers.{i}.{lin}.lora_A", (cfg.lora_r, inn), a, 0.0),
                  (f"model.layers.{i}.{lin}.lora_B", (out, cfg.lora_r), a, 0.0)]
    return s


# ------------------------------------------------------------------------------------------------
# LLaVA-1.5 / LLaVA-Next with peft LoRA on the decoder linears -- what every reference launch script actually trains
# (scripts/dpo_llava.sh:24-30, scripts/dpo_llavanext.sh:24-30, scripts/kto_llava.sh, scripts/ddpo_llava.sh:
# `--use_lora True --lora_r 128 --lora_alpha 256 --lora_target_modules auto`).  "auto" resolves to the short names of
# the language model's nn.Linear modules minus lm_head (models/Llava/__init__.py:273-286): q_proj, k_proj, v_proj,
# o_proj, gate_proj, up_proj, down_proj.  (peft matches those suffixes wherever they occur, i.e. it would also wrap the
# CLIP tower's q/k/v_proj; this build adapts the decoder only -- equivalent to spelling `--lora_target_modules` with
# the language_model prefix -- and keeps the tower frozen like `--freeze_vision_tower True` intends.)
# ------------------------------------------------------------------------------------------------
@dataclass
class LlavaLoRAModelConfig(ModelConfig):
    lora_r: int = 128
    lora_alpha: float = 256.0

    @property
    def lora_scale(self) -> float:
        return self.lora_alpha / self.lora_r


LLAVA_LORA_LINEARS = ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj", "mlp.gate_proj",
                      "mlp.up_proj", "mlp.down_proj")


def with_lora(cfg: ModelConfig, r: int = 128, alpha: float = 256.0) -> LlavaLoRAModelConfig:
    """The same architecture with decoder LoRA adapters of rank r (a LLaVA-1.5 or LLaVA-Next ModelConfig in)."""
    import dataclasses
    return LlavaLoRAModelConfig(**{f.name: getattr(cfg, f.name) for f in dataclasses.fields(ModelConfig)}, lora_r=r,
                                lora_alpha=alpha)


LLAVA15_7B_LORA = with_lora(LLAVA15_7B)
LLAVANEXT_MISTRAL_7B_LORA = with_lora(LLAVANEXT_MISTRAL_7B)
TINY_LORA = with_lora(TINY, 16, 32.0)
SMALL_LORA = with_lora(SMALL, 16, 32.0)
TINY_NEXT_LORA = with_lora(TINY_NEXT, 16, 32.0)
SMALL_NEXT_LORA = with_lora(SMALL_NEXT, 16, 32.0)


def llava_linear_dims(cfg: ModelConfig):
    d, hd, kvd = cfg.hidden, cfg.heads * cfg.head_dim, cfg.kv_heads * cfg.head_dim
    return {"self_attn.q_proj": (hd, d), "self_attn.k_proj": (kvd, d), "self_attn.v_proj": (kvd, d),
            "self_attn.o_proj": (d, hd), "mlp.gate_proj": (cfg.ff, d), "mlp.up_proj": (cfg.ff, d), "mlp.down_proj": (d, cfg.ff)}


def llava_lora_specs(cfg: LlavaLoRAModelConfig) -> List[Tuple[str, Tuple[int, ...], float, float]]:
    """`language_model.model.layers.{i}.<linear>.lora_A` [r, in] / `.lora_B` [out, r].  peft starts B at 0; the synthetic
    recipe draws B too so the adapter path carries signal in the parity tests."""
    a = 0.02 * math.sqrt(3.0)
    dims = llava_linear_dims(cfg)
    s = []
    for i in range(cfg.layers):
        for lin in LLAVA_LORA_LINEARS:
            out, inn = dims[lin]
            s += [(f"language_model.model.layers.{i}.{lin}.lora_A", (cfg.lora_r, inn), a, 0.0),
                  (f"language_model.model.layers.{i}.{lin}.lora_B", (out, cfg.lora_r), a, 0.0)]
    return s
