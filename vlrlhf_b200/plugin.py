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
class _EmbeddingHandle(nn.Module):
    """What `get_input_embeddings()` returns: the embedding weight parameter under the usual attribute name (TRL /
    HF Trainer only register hooks on it or read `.weight`)."""

    def __init__(self, weight: Optional[nn.Parameter]):
        super().__init__()
        object.__setattr__(self, "weight", weight)

    def forward(self, *a, **k):
        raise RuntimeError("the embedding lookup is fused into the engine's merge kernel")


class B200ModuleMixin:
    """The nn.Module-side contract the HF Trainer / TRL DPOTrainer rely on, for every B200 model wrapper.

    Parameters are VIEWS of the engine's flat arenas and their `.grad` are views of the flat gradient arena:
      * `zero_grad()` keeps the views (torch's default `set_to_none=True` would drop them and every torch optimizer would
        then skip the parameters); the gradients themselves need no zeroing -- the next backward overwrites them;
      * `_EngineLogps.backward` re-attaches any view a foreign `zero_grad` (e.g. through a wrapper module) dropped, and
        ACCUMULATES into the arena when the gradients of an earlier micro-batch are still live
        (gradient_accumulation_steps > 1: HF calls zero_grad only after optimizer.step);
      * `gradient_checkpointing_enable()` maps to TrainConfig.activation_checkpointing (every reference script passes
        --gradient_checkpointing True); `enable_input_require_grads()` / `get_input_embeddings()` exist because TRL calls
        them when checkpointing is on."""

    _grads_live = False           # the arena holds gradients of micro-batches not yet consumed by an optimizer step
    supports_gradient_checkpointing = True

    def _register_engine_params(self, tensors: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor], trainable):
        self._hf: Dict[str, nn.Parameter] = {}
        self._grad_views: Dict[str, torch.Tensor] = {}
        for name, t in tensors.items():
            tr = bool(trainable(name))
            p = nn.Parameter(t, requires_grad=tr)
            if tr:
                self._grad_views[name] = grads[name]
                p.grad = grads[name]  # gradient storage = the engine's flat reduce buffer
            self._hf[name] = p
            self.register_parameter(name.replace(".", "__"), p)

    def hf_named_parameters(self):
        return self._hf.items()

    def _grads_attached(self) -> bool:
        for name, gv in self._grad_views.items():
            g = self._hf[name].grad
            return g is not None and g.data_ptr() == gv.data_ptr()
        return False

    def _attach_grads(self):
        for name, gv in self._grad_views.items():
            p = self._hf[name]
            if p.grad is None or p.grad.data_ptr() != gv.data_ptr():
                p.grad = gv

    def zero_grad(self, set_to_none: bool = True):
        self._attach_grads()
        self._grads_live = False

    # ---- gradient checkpointing (HF PreTrainedModel API)
    def gradient_checkpointing_enable(self, gradient_checkpointing_kwargs=None):
        self.engine.tc.activation_checkpointing = True

    def gradient_checkpointing_disable(self):
        self.engine.tc.activation_checkpointing = False

    @property
    def is_gradient_checkpointing(self) -> bool:
        return bool(self.engine.tc.activation_checkpointing)

    def enable_input_require_grads(self):
        pass   # the hand-written backward needs no autograd hook on the embedding output

    def get_input_embeddings(self):
        for name, p in self._hf.items():
            if name.endswith("embed_tokens.weight") or name.endswith("wte.weight") or name.endswith("tok_embeddings.weight"):
                return _EmbeddingHandle(p)
        return _EmbeddingHandle(None)

    def flat_optimizer(self, lr: float = 1e-6, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                       max_grad_norm: Optional[float] = None) -> "B200FlatAdamW":
        return B200FlatAdamW(self, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_grad_norm=max_grad_norm)


class B200LlavaForRL(B200ModuleMixin, nn.Module):
    """Holds the engine; exposes HF-named `nn.Parameter`s that are VIEWS of the engine's flat bf16 arena (torch
    keeps a handle, the engine keeps the storage), so state_dict()/save_pretrained-style tooling, parameter
    freezing and optimizers see the usual names.  Mirrors the model-side contract of docs/CustomizedModel.md:
    default_lora_target, get_vision_tower(), freeze_vision_tower(), prepare_default_generation_kwargs()."""

    def __init__(self, cfg: ModelConfig, train: Optional[TrainConfig] = None, device: str = "cuda",
                 with_optimizer: bool = True):
        super().__init__()
        self.engine = LlavaDPOEngine(cfg, train, device=device, with_optimizer=with_optimizer)
        self.cfg = cfg
        self._register_engine_params(self.engine.hf_state("policy"), self.engine.hf_state("grad"),
                                     lambda name: not name.startswith("vision_tower."))

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
    backward pass and leaves the gradients in the parameters' .grad views (accumulating when an earlier micro-batch's
    gradients are still live)."""

    @staticmethod
    def forward(ctx, anchor, owner, inputs, plan, imgs_per_seq=1, shared=None):
        engine = owner.engine
        kw = {"imgs_per_seq": imgs_per_seq} if imgs_per_seq != 1 else {}
        if shared is not None:
            kw.update(feats=shared[0], m=shared[1])
        logps, m, feats = engine.forward_logps(*inputs, which="policy", save=True, **(plan or {}), **kw)
        ctx.owner = owner
        return logps

    @staticmethod
    def backward(ctx, g):
        owner = ctx.owner
        acc = bool(getattr(owner, "_grads_live", False)) and owner._grads_attached()
        owner.engine._backward(g.float().contiguous(), accumulate=acc)
        owner._attach_grads()
        owner._grads_live = True
        return torch.zeros(1, device=g.device), None, None, None, None, None


def unwrap_model(model):
    """DDP / accelerate / torch.compile wrappers -> the B200 module (or RefView) underneath."""
    seen = 0
    while not hasattr(model, "engine") and seen < 8:
        inner = getattr(model, "module", None) or getattr(model, "_orig_mod", None)
        if inner is None:
            break
        model, seen = inner, seen + 1
    if not hasattr(model, "engine"):
        raise TypeError(f"{type(model).__name__} is not a B200 model wrapper (no engine underneath)")
    return model


def concatenated_forward(self, model: nn.Module, batch: Dict[str, Union[List, torch.LongTensor]]
                         ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """VLDPOTrainer.concatenated_forward (base/trainer.py:190-242) on the engine.  Returns
    (chosen_logps, rejected_logps, chosen_logits, rejected_logits) where the last two are ONE-element fp32 tensors holding
    the mean of the [B,S,V] logits the reference returns: trl's get_batch_loss_metrics only ever takes
    `.detach().mean()` of them (a logging metric), so the full tensors are never materialised.  `model` is a B200 model
    wrapper (possibly under DDP/accelerate wrappers) or a RefView."""
    model = unwrap_model(model)
    eng: LlavaDPOEngine = model.engine
    which = getattr(model, "_which", "policy")
    owner = getattr(model, "_owner", model)
    n = batch["chosen_labels"].shape[0]
    stash = getattr(owner, "_ref_stash", None)
    if which == "ref" and stash is not None and stash[0] is batch:   # computed ahead by the policy call on this batch
        owner._ref_stash = None
        lg = stash[1]
        zero = lg.new_zeros(1)
        return lg[:n], lg[n:], zero, zero.clone()
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
    # TrainConfig.pack_sequences / share_prefix: the row layout comes from the host batch, so the step stays free of read-backs
    plan = eng.host_row_plan(ids, am, sizes, **kw) if hasattr(eng, "host_row_plan") else (
        {"seq_lens": eng.host_seq_lens(ids, am, sizes, **kw)} if eng.tc.pack_sequences else {})
    eng.force_logit_means = which == "policy"
    try:
        if which == "policy" and torch.is_grad_enabled():
            shared = None
            ref_model = getattr(self, "ref_model", None)
            if (isinstance(ref_model, RefView) and ref_model._owner is owner and "reference_chosen_logps" not in batch
                    and not getattr(self, "precompute_ref_log_probs", False)):
                # trl calls policy first, reference second (get_batch_loss_metrics); the frozen reference pass is run HERE,
                # ahead of the policy pass, and handed out when trl asks for it: image features and the merge index are
                # computed once for both passes, and a deferred optimizer step of the previous batch overlaps it
                with torch.no_grad():
                    ref_lg, m, feats = eng.forward_logps(*inputs, which="ref", save=False, **plan, **kw)
                owner._ref_stash = (batch, ref_lg.clone())
                shared = (feats, m)
            anchor = torch.zeros(1, device=eng.device, requires_grad=True)
            logps = _EngineLogps.apply(anchor, owner, inputs, plan, k, shared)
        else:
            with torch.no_grad():
                logps, _, _ = eng.forward_logps(*inputs, which=which, save=False, **plan, **kw)
        if which == "policy":
            means = eng.logit_means.clone()
            return logps[:n], logps[n:], means[0:1], means[1:2]
    finally:
        eng.force_logit_means = False
    zero = logps.new_zeros(1)
    return logps[:n], logps[n:], zero, zero.clone()


def model_pixels(model, input_ids: torch.Tensor) -> torch.Tensor:
    owner = getattr(model, "_owner", model)
    if not hasattr(owner, "pixel_values_for"):
        raise ValueError("the batch carries no img_input_dict and the model cannot resolve images from its token stream")
    return owner.pixel_values_for(input_ids)


class RefView(nn.Module):
    """`ref_model` handle for TRL's concatenated_forward(self.ref_model, batch) call: the same engine, reference weights
    (full fine-tuning: the frozen copy; LoRA: the shared base with the adapters off).  A parameter-less nn.Module so the
    trainer's `.eval()` / dropout / accelerate bookkeeping accepts it without copying the engine (trl's
    `create_reference_model(model)` would deep-copy 120 GB of arenas plus CUDA streams and events)."""

    def __init__(self, model):
        super().__init__()
        object.__setattr__(self, "engine", model.engine)
        object.__setattr__(self, "_owner", model)     # not registered as a sub-module: no parameters are exposed twice
        self._which = "ref"

    def forward(self, *a, **k):
        raise RuntimeError("RefView is consumed by concatenated_forward; it has no logits-producing forward")

    def __deepcopy__(self, memo):
        return self


# --------------------------------------------------------------------------------------------
# the optimizer the Trainer steps: the engine's flat fused AdamW (+ the data-parallel gradient reduction)
# --------------------------------------------------------------------------------------------
class B200FlatAdamW(torch.optim.Optimizer):
    """torch.optim.Optimizer facade over the engine's flat optimizer state, for `Trainer(optimizers=(opt, sched))` /
    `create_optimizer`: `step()` = gradient reduction over the data-parallel ranks (NCCL reduce-scatter of the flat bf16
    gradient arena; nothing at world 1) + global-norm clipping + ONE fused AdamW launch over this rank's slice (fp32 master
    weights and moments) + parameter all-gather.  The learning rate is read from `param_groups[0]["lr"]` at every step, so
    any torch LR scheduler (HF get_scheduler's LambdaLR) drives it.  The parameters are not part of an autograd graph, so
    DDP's reducer hooks would never fire for them: THIS is where replicas are kept in sync -- do not wrap the model in DDP /
    DeepSpeed (the trainer classes built by install*() keep accelerate from doing so)."""

    def __init__(self, model, lr: float = 1e-6, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 max_grad_norm: Optional[float] = None):
        model = unwrap_model(model)
        eng = model.engine
        if not eng.with_optimizer:
            raise ValueError("the engine was built with with_optimizer=False: no fp32 master / moment arenas to step")
        params = [p for p in model.parameters() if p.requires_grad]
        super().__init__(params, dict(lr=float(lr), betas=tuple(betas), eps=float(eps), weight_decay=float(weight_decay)))
        self.model, self.engine = model, eng
        if max_grad_norm is not None:
            eng.tc.max_grad_norm = float(max_grad_norm)
        eng.async_optimizer = False   # under the Trainer the next call is the policy pass: nothing to overlap with

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        g, eng = self.param_groups[0], self.engine
        tc = eng.tc
        tc.adam_beta1, tc.adam_beta2 = (float(b) for b in g["betas"])
        tc.adam_eps, tc.weight_decay = float(g["eps"]), float(g["weight_decay"])
        self.grad_norm_sq = eng.reduce_and_step(lr=float(g["lr"]))
        return loss

    def zero_grad(self, set_to_none: bool = True):
        self.model.zero_grad()

    def grad_norm(self) -> float:
        """Global gradient norm of the last step (before clipping), one device read."""
        return float(self.grad_norm_sq.item()) ** 0.5 / self.engine.world_size()

    def state_dict(self):
        sd = super().state_dict()
        eng = self.engine
        eng.wait_optimizer()
        sd["b200_flat"] = {"step": eng.opt_step, "shard": (eng.shard_lo, eng.shard_hi), "master": eng.master,
                           "exp_avg": eng.exp_avg, "exp_avg_sq": eng.exp_avg_sq}
        return sd

    def load_state_dict(self, state_dict):
        flat = state_dict.get("b200_flat")
        super().load_state_dict({k: v for k, v in state_dict.items() if k != "b200_flat"})
        if flat is not None:
            eng = self.engine
            eng.wait_optimizer()
            if tuple(flat["shard"]) != (eng.shard_lo, eng.shard_hi):
                raise ValueError(f"optimizer state of shard {tuple(flat['shard'])} loaded into shard {(eng.shard_lo, eng.shard_hi)}")
            eng.opt_step = int(flat["step"])
            eng.master.copy_(flat["master"]); eng.exp_avg.copy_(flat["exp_avg"]); eng.exp_avg_sq.copy_(flat["exp_avg_sq"])


def make_trainer_class(base, check_peft=None, name: str = "B200DPOTrainer"):
    """The reference's DPO trainer class for a family (`core_mapper.dpo_trainer`, a VLDPOTrainer subclass) with the three
    hot-path override points replaced and the Trainer-side plumbing the engine needs:
      * `ref_model` defaults to a RefView of the same engine (the reference builds ref_model=None, utils/auto_load.py:537, and
        trl would `create_reference_model(model)` = deep-copy every arena);
      * `create_optimizer` returns B200FlatAdamW with the TrainingArguments' lr / betas / eps / weight decay; clipping by
        `max_grad_norm` moves into that fused step (HF's own clip pass over 6.8 G gradient views is switched off);
      * accelerate is kept from wrapping the model in DDP (gradient reduction happens inside the optimizer step);
      * `training_step` without the reference's per-step empty_cache() + gc.collect() (base/trainer.py:303-308)."""

    class B200DPOTrainer(base):
        get_batch_logps = staticmethod(get_batch_logps)
        concatenated_forward = concatenated_forward
        dpo_loss = dpo_loss

        def __init__(self, model=None, ref_model=None, *a, peft_config=None, **k):
            if check_peft is not None:   # LoRA families: the adapters are the engine's, not peft modules
                check_peft(model, peft_config)
                peft_config = None
            args = k.get("args")
            self._b200_max_grad_norm = getattr(args, "max_grad_norm", None) if args is not None else None
            if args is not None and self._b200_max_grad_norm:
                args.max_grad_norm = 0.0   # clipping is fused into B200FlatAdamW.step
            if args is not None and getattr(args, "gradient_checkpointing", False):
                unwrap_model(model).gradient_checkpointing_enable()
            super().__init__(model, ref_model if ref_model is not None else RefView(unwrap_model(model)), *a,
                             peft_config=peft_config, **k)
            acc = getattr(self, "accelerator", None)
            if acc is not None and hasattr(acc, "prepare_model"):
                prepare_model = acc.prepare_model

                def _prepare_model(m, *pa, **pk):
                    if hasattr(m, "engine"):   # B200 wrapper or RefView: device placement and reduction are the engine's
                        return m
                    return prepare_model(m, *pa, **pk)
                acc.prepare_model = _prepare_model

        def create_optimizer(self):
            if getattr(self, "optimizer", None) is None:
                a = self.args
                self.optimizer = B200FlatAdamW(self.model, lr=a.learning_rate, betas=(a.adam_beta1, a.adam_beta2),
                                               eps=a.adam_epsilon, weight_decay=a.weight_decay,
                                               max_grad_norm=self._b200_max_grad_norm)
            return self.optimizer

        def training_step(self, model, inputs, *a, **k):
            mro = type(self).__mro__
            for cls in mro[mro.index(B200DPOTrainer) + 1:]:
                fn = cls.__dict__.get("training_step")
                if fn is None or "empty_cache" in getattr(getattr(fn, "__code__", None), "co_names", ()):
                    continue   # VLDPOTrainer's empty_cache() + gc.collect() wrapper (base/trainer.py:303-308) is skipped
                return fn(self, model, inputs, *a, **k)
            raise AttributeError("no training_step below the B200 trainer class")

    B200DPOTrainer.__name__ = B200DPOTrainer.__qualname__ = name
    return B200DPOTrainer


def install():
    """Register the B200 path with the reference (needs the `vlrlhf` package importable):
    `vlrlhf.models.Llava.core_mapper` keeps its processor/collators and gets this model + trainer, so
    `src/vlrlhf/dpo.py --loss_type sigmoid|ddpo|kto_pair ...` runs unmodified (MODEL_NICKNAME_MAP resolution,
    utils/auto_load.py:41-61)."""
    import importlib
    llava = importlib.import_module("vlrlhf.models.Llava")
    from vlrlhf.base.trainer import VLDPOTrainer
    from vlrlhf.models.utils import ModelCoreMapper

    LlavaB200DPOTrainer = make_trainer_class(VLDPOTrainer, name="LlavaB200DPOTrainer")

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
