"""TEST-ONLY restatement of the Trainer-side control flow that drives the plugin in the reference: trl 0.8.1
`DPOTrainer.get_batch_loss_metrics` / `compute_loss`, and transformers 4.41 `Trainer.training_step` +
`_inner_training_loop` (gradient accumulation, clipping, optimizer / scheduler step, `model.zero_grad()`).

trl / accelerate / peft are not installable here (no wheels), so the plugin's trainer classes can never be built over the
real `VLDPOTrainer`; this stub base class has the same constructor surface (the arguments src/vlrlhf/dpo.py passes,
dpo.py:120-141) and the same call sequence, so `plugin.make_trainer_class(StubDPOTrainer)` exercises exactly the code the
real class would run: the three override points, the RefView default, `create_optimizer`, clipping hand-over, zero_grad.
The restated pieces cite the upstream lines they follow; the oracle's own restatement of the metrics
(oracle/restate.py get_batch_loss_metrics) is what the values are checked against in the tests.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List

import torch


def training_args(**kw):
    """The TrainingArguments fields the loop reads (HF defaults; scripts/dpo_llava.sh overrides some)."""
    d = dict(learning_rate=1e-6, adam_beta1=0.9, adam_beta2=0.999, adam_epsilon=1e-8, weight_decay=0.0, max_grad_norm=1.0,
             gradient_accumulation_steps=1, lr_scheduler_type="constant", warmup_steps=0, warmup_ratio=0.0, max_steps=0,
             gradient_checkpointing=False)
    d.update(kw)
    return SimpleNamespace(**d)


class StubDPOTrainer:
    """Constructor surface of VLDPOTrainer (base/trainer.py:34-107) + the trl 0.8.1 / HF 4.41 methods on the hot path."""

    def __init__(self, model=None, ref_model=None, beta: float = 0.1, label_smoothing: float = 0.0,
                 loss_type: str = "sigmoid", args=None, data_collator=None, label_pad_token_id: int = -100,
                 padding_value: int = 0, processor=None, peft_config=None, precompute_ref_log_probs: bool = False,
                 reference_free: bool = False, **unused):
        if peft_config is not None:
            raise RuntimeError("a peft_config reached trl: it would wrap the engine-backed model in peft modules")
        self.model, self.ref_model = model, ref_model
        self.beta, self.label_smoothing, self.loss_type = beta, label_smoothing, loss_type
        self.args = args if args is not None else training_args()
        self.label_pad_token_id, self.padding_value = label_pad_token_id, padding_value
        self.is_encoder_decoder = False
        self.precompute_ref_log_probs = precompute_ref_log_probs
        self.reference_free = reference_free
        self.optimizer, self.lr_scheduler = None, None
        self.logged: List[Dict[str, float]] = []

    # -- trl 0.8.1 dpo_trainer.py get_batch_loss_metrics (the oracle restates the same lines: oracle/restate.py)
    def get_batch_loss_metrics(self, model, batch, train_eval: str = "train"):
        metrics = {}
        pc, pr, pcl, prl = self.concatenated_forward(model, batch)
        if "reference_chosen_logps" in batch and "reference_rejected_logps" in batch:
            rc, rr = batch["reference_chosen_logps"], batch["reference_rejected_logps"]
        else:
            with torch.no_grad():
                if self.ref_model is None:
                    raise RuntimeError("ref_model is None: trl would run null_ref_context() on a peft model")
                rc, rr, _, _ = self.concatenated_forward(self.ref_model, batch)
        losses, chosen_rewards, rejected_rewards = self.dpo_loss(pc, pr, rc, rr)
        reward_accuracies = (chosen_rewards > rejected_rewards).float()
        prefix = "eval_" if train_eval == "eval" else ""
        metrics[f"{prefix}rewards/chosen"] = chosen_rewards.mean().cpu()
        metrics[f"{prefix}rewards/rejected"] = rejected_rewards.mean().cpu()
        metrics[f"{prefix}rewards/accuracies"] = reward_accuracies.mean().cpu()
        metrics[f"{prefix}rewards/margins"] = (chosen_rewards - rejected_rewards).mean().cpu()
        metrics[f"{prefix}logps/rejected"] = pr.detach().mean().cpu()
        metrics[f"{prefix}logps/chosen"] = pc.detach().mean().cpu()
        metrics[f"{prefix}logits/rejected"] = prl.detach().mean().cpu()
        metrics[f"{prefix}logits/chosen"] = pcl.detach().mean().cpu()
        self._last = dict(pc=pc, pr=pr, rc=rc, rr=rr, losses=losses)
        return losses.mean(), metrics

    def compute_loss(self, model, inputs):
        loss, metrics = self.get_batch_loss_metrics(model, inputs, train_eval="train")
        self.logged.append({k: float(v) for k, v in metrics.items()})
        return loss

    # -- transformers 4.41 Trainer.training_step (trainer.py: loss / gradient_accumulation_steps happens inside
    #    accelerator.backward via the accumulate context; the arithmetic is the same)
    def training_step(self, model, inputs):
        if hasattr(model, "train"):
            model.train()
        loss = self.compute_loss(model, inputs)
        (loss / self.args.gradient_accumulation_steps).backward()
        return loss.detach()

    def create_optimizer(self):
        if self.optimizer is None:
            a = self.args
            self.optimizer = torch.optim.AdamW([p for p in self.model.parameters() if p.requires_grad], lr=a.learning_rate,
                                               betas=(a.adam_beta1, a.adam_beta2), eps=a.adam_epsilon,
                                               weight_decay=a.weight_decay)
        return self.optimizer

    def create_scheduler(self, num_training_steps: int):
        import math
        from transformers import get_scheduler
        a = self.args
        warm = a.warmup_steps if a.warmup_steps > 0 else math.ceil(num_training_steps * a.warmup_ratio)
        kind = "constant_with_warmup" if a.lr_scheduler_type == "constant" and warm else a.lr_scheduler_type
        self.lr_scheduler = get_scheduler(kind, self.optimizer, num_warmup_steps=warm, num_training_steps=num_training_steps)
        return self.lr_scheduler

    # -- transformers 4.41 Trainer._inner_training_loop, the part between two optimizer steps
    def train_loop(self, batches: List[Dict]):
        a = self.args
        ga = a.gradient_accumulation_steps
        self.create_optimizer()
        self.create_scheduler(a.max_steps if a.max_steps > 0 else len(batches) // ga)
        self.model.zero_grad()
        losses = []
        for i, batch in enumerate(batches):
            losses.append(self.training_step(self.model, batch))
            if (i + 1) % ga == 0:
                if a.max_grad_norm is not None and a.max_grad_norm > 0:
                    torch.nn.utils.clip_grad_norm_([p for p in self.model.parameters() if p.requires_grad], a.max_grad_norm)
                self.optimizer.step()
                self.lr_scheduler.step()
                self.model.zero_grad()
        return losses
