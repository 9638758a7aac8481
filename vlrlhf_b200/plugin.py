"""Drop-in for the reference's plugin/operator API on the DPO hot path.

The reference resolves a model family through `ModelCoreMapper` (src/vlrlhf/models/utils.py:18-31) and drives the
step through three `VLDPOTrainer` override points (src/vlrlhf/base/trainer.py):

    get_batch_logps(logits, labels, average_log_prob, label_pad_token_id, is_encoder_decoder, mask_shared_tokens)  :148-188
    concatenated_forward(self, model, batch) -> (chosen_logps, rejected_logps, chosen_logits, rejected_logits)      :190-242
    dpo_loss(self, pc, pr, rc, rr) -> (losses, chosen_rewards, rejected_rewards)                                    :244-301

This module provides the same three callables (same names, argument meaning, return values and exception
types) backed by libvlb200, plus `B200LlavaForRL` (an nn.Module whose HF-named parameters are views of the
engine's flat arenas) and `install()` which registers a `core_mapper` with the reference when the `vlrlhf`
package is importable.  Nothing here falls back to PyTorch math: tensors must be CUDA tensors.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn

from . import host, ops
from .config import ModelConfig, TrainConfig
from .engine import LlavaDPOEngine


# --------------------------------------------------------------------------------------------
# get_batch_logps  (autograd-aware: backward is the fused dlogits kernel)
# --------------------------------------------------------------------------------------------
class _BatchLogps(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits2d, target, n_seq, weight, average):
        logps, _, lse = ops.logps_fwd(logits2d, target, n_seq, weight=weight, average_log_prob=average)
        ctx.save_for_backward(logits2d, target, lse, weight if weight is not None else torch.empty(0))
        ctx.n_seq, ctx.average, ctx.has_w = n_seq, average, weight is not None
        return logps

    @staticmethod
    def backward(ctx, g):
        logits2d, target, lse, w = ctx.saved_tensors
        d = ops.logps_bwd(logits2d, target, ctx.n_seq, lse, g.float().contiguous(), weight=w if ctx.has_w else None,
                          average_log_prob=ctx.average)
        return d.to(logits2d.dtype), None, None, None, None


def get_batch_logps(logits: torch.Tensor, labels: torch.Tensor, average_log_prob: bool = False,
                    label_pad_token_id: int = -100, is_encoder_decoder: bool = False,
                    mask_shared_tokens: bool = False) -> torch.Tensor:
    """VLDPOTrainer.get_batch_logps (base/trainer.py:148-188) on materialised logits [2B,S,V] (bf16|f32).
    fp32 math regardless of the logits dtype; returns fp32 [2B]."""
    if logits.shape[:-1] != labels.shape:
        raise ValueError("Logits (batch and sequence length dim) and labels must have the same shape.")
    if is_encoder_decoder:
        raise ValueError("encoder-decoder models are not supported by the B200 path")
    n_seq, S, V = logits.shape
    # shift (trainer.py:161-162): row (b,t) predicts labels[b,t+1]; the last row of each sequence predicts nothing
    target = torch.full((n_seq, S), -100, dtype=torch.int64, device=labels.device)
    target[:, :-1] = labels[:, 1:]
    target[target == label_pad_token_id] = -100
    weight = None
    if mask_shared_tokens:
        assert n_seq % 2 == 0
        shift = labels[:, 1:].clone()
        shift[shift == label_pad_token_id] = 0  # trainer.py:166
        rows = shift.tolist()  # the reference's device->host sync (trainer.py:177-180)
        w = torch.zeros(n_seq, S, dtype=torch.uint8)
        for i in range(n_seq // 2):
            c_mod, r_mod = host.get_diff_ids(rows[i], rows[n_seq // 2 + i], min_match_size=3)
            w[i, c_mod] = 1
            w[n_seq // 2 + i, r_mod] = 1
        weight = w.to(logits.device).reshape(-1)
    lg = logits.contiguous().view(n_seq * S, V)
    return _BatchLogps.apply(lg, target.reshape(-1).contiguous(), n_seq, weight, bool(average_log_prob))


# --------------------------------------------------------------------------------------------
# dpo_loss
# --------------------------------------------------------------------------------------------
class _DPOLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pol, ref, beta, ls, loss_type, reference_free):
        losses, cr, rr, stats, grad = ops.dpo_loss(pol, ref, beta, ls, loss_type, reference_free, 1.0, want_grad=True)
        ctx.save_for_backward(grad)
        ctx.n_losses, ctx.kto = losses.numel(), loss_type == "kto_pair"
        ctx.mark_non_differentiable(cr, rr)
        return losses, cr, rr

    @staticmethod
    def backward(ctx, g_losses, _g1, _g2):
        (grad_mean,) = ctx.saved_tensors  # d mean(losses) / d policy_logps
        n = grad_mean.numel() // 2
        if ctx.kto:
            # kto_pair couples every pair through the batch-mean KL terms: exact for the (universal) losses.mean() use
            if not torch.allclose(g_losses, g_losses[0].expand_as(g_losses)):
                raise NotImplementedError("kto_pair backward supports a uniform upstream gradient (losses.mean())")
            return grad_mean * (g_losses[0] * ctx.n_losses), None, None, None, None, None
        g = torch.cat([g_losses, g_losses]) * ctx.n_losses
        return grad_mean * g, None, None, None, None, None


def dpo_loss(self, policy_chosen_logps: torch.Tensor, policy_rejected_logps: torch.Tensor,
             reference_chosen_logps: torch.Tensor, reference_rejected_logps: torch.Tensor
             ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """VLDPOTrainer.dpo_loss (base/trainer.py:244-301).  `self` provides beta, label_smoothing, loss_type,
    reference_free exactly like the trainer instance the reference passes."""
    if self.loss_type not in ops.LOSS_TYPES:
        raise ValueError(f"Unknown loss type: {self.loss_type}. Should be one of ['sigmoid', 'hinge', 'ipo', 'kto_pair']")
    pol = torch.cat([policy_chosen_logps, policy_rejected_logps]).float()
    ref = torch.cat([reference_chosen_logps, reference_rejected_logps]).float().to(pol.device)
    return _DPOLoss.apply(pol, ref, float(self.beta), float(self.label_smoothing), self.loss_type, bool(self.reference_free))


# --------------------------------------------------------------------------------------------
# model wrapper + concatenated_forward
# --------------------------------------------------------------------------------------------
class B200LlavaForRL(nn.Module):
    """Holds the engine; exposes HF-named `nn.Parameter`s that are VIEWS of the engine's flat bf16 arena (torch
    keeps a handle, the engine keeps the storage), so state_dict()/save_pretrained-style tooling, parameter
    freezing and optimizers see the usual names.  Mirrors the model-side contract of docs/CustomizedModel.md:
    default_lora_target, get_vision_tower(), freeze_vision_tower(), prepare_default_generation_kwargs()."""

    def __init__(self, cfg: ModelConfig, train: Optional[TrainConfig] = None, device: str = "cuda",
                 with_optimizer: bool = True):
        super().__init__()
        self.engine = LlavaDPOEngine(cfg, train, device=device, with_optimizer=with_optimizer)
        self.cfg = cfg
        grads = self.engine.hf_state("grad")
        self._hf = {}
        for name, t in self.engine.hf_state("policy").items():
            trainable = not name.startswith("vision_tower.")
            p = nn.Parameter(t, requires_grad=trainable)
            if trainable:
                p.grad = grads[name]  # gradient storage = the engine's flat all-reduce buffer
            self._hf[name] = p
            self.register_parameter(name.replace(".", "__"), p)

    def hf_named_parameters(self):
        return self._hf.items()

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, *args, config=None, torch_dtype=None,
                        train: Optional[TrainConfig] = None, device: str = "cuda", with_optimizer: bool = True, **kwargs):
        """`MyAutoModel.from_pretrained(path, config=config, torch_dtype=…)` of utils/auto_load.py:522-535: a local HF
        checkpoint directory (config.json + safetensors) streamed into the flat arenas; the reference copy starts equal
        to the policy (TRL deep-copies the model when no ref_model is given).  HF-only kwargs (device_map,
        quantization_config, use_flash_attention_2, …) are accepted and ignored; bf16 is the only compute dtype."""
        import json
        import os
        from . import checkpoint
        if torch_dtype not in (None, torch.bfloat16, "bfloat16", "auto"):
            raise ValueError(f"torch_dtype {torch_dtype}: the B200 path computes in bf16 only")
        path = pretrained_model_name_or_path
        if not os.path.isdir(path):
            raise FileNotFoundError(f"{path}: a local checkpoint directory is required (there is no hub access)")
        with open(os.path.join(path, "config.json")) as f:
            cfg_dict = json.load(f)
        model = cls(checkpoint.config_from_hf(config if config is not None else cfg_dict), train, device=device,
                    with_optimizer=with_optimizer)
        model.hf_config_dict = cfg_dict
        model.config = config
        checkpoint.load_hf_checkpoint(model.engine, path)
        return model

    def save_pretrained(self, save_directory: str, max_shard_size: int = 5 << 30, **kwargs):
        """HF-layout export of the trained policy (safetensors shards + index + config.json, transformers-4.41 names)
        so the reference's eval harness / merge scripts reload it (dpo.py:89-95,147-149; utils/common.py:21-55)."""
        from . import checkpoint
        return checkpoint.save_hf_checkpoint(self.engine, save_directory, getattr(self, "hf_config_dict", None),
                                             max_shard_bytes=int(max_shard_size))

    @property
    def default_lora_target(self) -> List[str]:  # Llava/__init__.py:273-286, LlavaNext/__init__.py:347-360
        return ["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"]

    def get_vision_tower(self):
        return {k: v for k, v in self._hf.items() if k.startswith("vision_tower.")}

    def freeze_vision_tower(self):  # always frozen on this path (dpo.py:55 default)
        for k, v in self._hf.items():
            if k.startswith("vision_tower."):
                v.requires_grad_(False)

    def prepare_default_generation_kwargs(self, generation_config):  # Llava/__init__.py:294-298
        generation_config.max_new_tokens = 1024
        generation_config.do_sample = False
        return dict(generation_config=generation_config)

    def forward(self, *a, **k):
        raise RuntimeError("B200LlavaForRL is driven through concatenated_forward / engine.train_step; a full-logits "
                           "forward (generation, eval) is outside the hot path this package replaces")


class _EngineLogps(torch.autograd.Function):
    """Policy log-probs as a differentiable function of the engine's parameters: backward runs the hand-written
    backward pass and leaves the gradients in the parameters' .grad views."""

    @staticmethod
    def forward(ctx, anchor, engine, inputs, seq_lens, imgs_per_seq=1):
        logps, m, feats = engine.forward_logps(*inputs, which="policy", save=True, seq_lens=seq_lens,
                                               **({"imgs_per_seq": imgs_per_seq} if imgs_per_seq != 1 else {}))
        ctx.engine = engine
        return logps

    @staticmethod
    def backward(ctx, g):
        ctx.engine._backward(g.float().contiguous())
        return torch.zeros(1, device=g.device), None, None, None, None


def concatenated_forward(self, model: nn.Module, batch: Dict[str, Union[List, torch.LongTensor]]
                         ) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]:
    """VLDPOTrainer.concatenated_forward (base/trainer.py:190-242) on the engine.  Returns
    (chosen_logps, rejected_logps, None, None): the [B,S,V] logits the reference returns only feed a logging
    `.mean()` and are never materialised here.  `model` is a B200LlavaForRL or the string 'ref' wrapper."""
    eng: LlavaDPOEngine = model.engine
    which = getattr(model, "_which", "policy")
    cb = host.concatenated_inputs(batch, getattr(self, "is_encoder_decoder", False),
                                  getattr(self, "label_pad_token_id", -100), getattr(self, "padding_value", 0) or 0)
    ids, am, lb = host.right_pad_valid_tokens(cb["concatenated_input_ids"], cb["concatenated_attention_mask"],
                                              cb["concatenated_labels"], getattr(self, "padding_value", 0) or 0,
                                              getattr(self, "label_pad_token_id", -100), self.loss_type)
    if "img_input_dict" in batch:
        px = batch["img_input_dict"]["pixel_values"]
        sizes = batch["img_input_dict"].get("image_sizes")  # LLaVA-Next (LlavaNext/__init__.py:348-380 collators)
    else:  # Qwen-VL: the images are named inside the token stream (modeling_qwen.py:524-537)
        px, sizes = model_pixels(model, batch["chosen_input_ids"]), None
    wt = None
    if self.loss_type == "ddpo":
        wt = eng.ddpo_weights(ids, am, lb, sizes)
    k = eng.images_per_sequence(batch) if "img_input_dict" in batch else 1   # several <image> placeholders per sequence
    kw = {"imgs_per_seq": k} if k != 1 else {}
    inputs = eng.prepare_inputs(ids, am, lb, px, wt, sizes, **kw)
    # TrainConfig.pack_sequences: the merged lengths come from the host batch, so the step stays free of device read-backs
    seq_lens = eng.host_seq_lens(ids, am, sizes, **kw) if eng.tc.pack_sequences else None
    n = batch["chosen_labels"].shape[0]
    if which == "policy" and torch.is_grad_enabled():
        anchor = torch.zeros(1, device=eng.device, requires_grad=True)
        logps = _EngineLogps.apply(anchor, eng, inputs, seq_lens, k)
    else:
        with torch.no_grad():
            logps, _, _ = eng.forward_logps(*inputs, which=which, save=False, seq_lens=seq_lens, **kw)
    return logps[:n], logps[n:], None, None


def model_pixels(model, input_ids: torch.Tensor) -> torch.Tensor:
    owner = getattr(model, "_owner", model)
    if not hasattr(owner, "pixel_values_for"):
        raise ValueError("the batch carries no img_input_dict and the model cannot resolve images from its token stream")
    return owner.pixel_values_for(input_ids)


class RefView(nn.Module):
    """`ref_model` handle for TRL's concatenated_forward(self.ref_model, batch) call: the same engine, reference weights
    (full fine-tuning: the frozen copy; LoRA: the shared base with the adapters off).  A parameter-less nn.Module so the
    trainer's `.eval()` / dropout / accelerate bookkeeping accepts it without copying the engine."""

    def __init__(self, model):
        super().__init__()
        object.__setattr__(self, "engine", model.engine)
        object.__setattr__(self, "_owner", model)     # not registered as a sub-module: no parameters are exposed twice
        self._which = "ref"

    def forward(self, *a, **k):
        raise RuntimeError("RefView is consumed by concatenated_forward; it has no logits-producing forward")


def install():
    """Register the B200 path with the reference (needs the `vlrlhf` package importable):
    `vlrlhf.models.Llava.core_mapper` keeps its processor/collators and gets this model + trainer, so
    `src/vlrlhf/dpo.py --loss_type sigmoid|ddpo|kto_pair ...` runs unmodified (MODEL_NICKNAME_MAP resolution,
    utils/auto_load.py:41-61)."""
    import importlib
    llava = importlib.import_module("vlrlhf.models.Llava")
    from vlrlhf.base.trainer import VLDPOTrainer
    from vlrlhf.models.utils import ModelCoreMapper

    class LlavaB200DPOTrainer(VLDPOTrainer):
        get_batch_logps = staticmethod(get_batch_logps)
        concatenated_forward = concatenated_forward
        dpo_loss = dpo_loss

        def training_step(self, model, inputs):  # trainer.py:303-308 without the per-step empty_cache()/gc
            return super(VLDPOTrainer, self).training_step(model, inputs)

    def swap(mod):
        ref = mod.core_mapper
        mod.core_mapper = ModelCoreMapper(
            model=B200LlavaForRL, processor=ref.processor, dpo_collator=ref.dpo_collator, dpo_trainer=LlavaB200DPOTrainer,
            reward_model=ref.reward_model, value_model=ref.value_model, reward_collator=ref.reward_collator,
            reward_trainer=ref.reward_trainer, sft_collator=ref.sft_collator, sft_trainer=ref.sft_trainer,
            ppo_collator=ref.ppo_collator, ppo_trainer=ref.ppo_trainer)
        return mod.core_mapper

    mapper = swap(llava)
    try:  # LLaVA-Next shares the trainer; its processor / collators (image_sizes) stay the reference's
        swap(importlib.import_module("vlrlhf.models.LlavaNext"))
    except Exception:  # the reference's LlavaNext module needs a transformers that still has the 4.41 symbols
        pass
    return mapper
