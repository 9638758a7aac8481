"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- CPU restatement of the LoRA variant of the LLaVA-1.5 / LLaVA-Next
hot path: what every reference launch script trains (scripts/dpo_llava.sh:24-30, scripts/dpo_llavanext.sh:24-30,
scripts/kto_llava.sh, scripts/ddpo_llava.sh: `--use_lora True --lora_r 128 --lora_alpha 256 --lora_target_modules auto`).

Follows:
  * utils/auto_load.py:559-578      LoraConfig(r, lora_alpha, target_modules = model.default_lora_target, bias none)
  * models/Llava/__init__.py:273-286 / models/LlavaNext/__init__.py:347-360  default_lora_target = the language model's
                                    nn.Linear short names minus lm_head (q/k/v/o_proj, gate/up/down_proj)
  * trl 0.8.1 DPOTrainer (not on disk): with a peft model and ref_model None the reference pass runs the SAME model under
                                    `null_ref_context()` = `disable_adapter()`  ->  reference = base weights
  * peft lora.Linear.forward (not on disk): base(x) + lora_B(lora_A(x)) * alpha / r      (restate.lora_linear)
  * everything else: oracle/restate.py (LlavaForRL / LlavaNextForRL forward, get_batch_logps, dpo_loss).
Adapters sit on the decoder linears only (peft's suffix match would also wrap the CLIP tower's q/k/v_proj; see
vlrlhf_b200/config.py).  Pinned against the reference's LlavaForRL / LlavaNextForRL run here with the adapters applied by
hand (tests/golden/g11_*.npz, make_fixtures.py --lora).
"""
from __future__ import annotations

import dataclasses
import math
from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import restate as R


@dataclass
class LoraCfg(R.LlavaCfg):
    lora_r: int = 128
    lora_alpha: float = 256.0

    @property
    def lora_scale(self) -> float:
        return self.lora_alpha / self.lora_r


def with_lora(cfg: R.LlavaCfg, r: int = 128, alpha: float = 256.0) -> LoraCfg:
    return LoraCfg(**{f.name: getattr(cfg, f.name) for f in dataclasses.fields(R.LlavaCfg)}, lora_r=r, lora_alpha=alpha)


LLAVA15_7B_LORA = with_lora(R.LLAVA15_7B)
LLAVANEXT_MISTRAL_7B_LORA = with_lora(R.LLAVANEXT_MISTRAL_7B)
TINY_LORA = with_lora(R.TINY, 16, 32.0)
SMALL_LORA = with_lora(R.SMALL, 16, 32.0)
TINY_NEXT_LORA = with_lora(R.TINY_NEXT, 16, 32.0)
SMALL_NEXT_LORA = with_lora(R.SMALL_NEXT, 16, 32.0)

LINEARS = ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj", "mlp.gate_proj", "mlp.up_proj",
           "mlp.down_proj")


def _dims(cfg: LoraCfg) -> Dict[str, Tuple[int, int]]:
    d, hd, kvd = cfg.hidden, cfg.heads * cfg.head_dim, cfg.kv_heads * cfg.head_dim
    return {"self_attn.q_proj": (hd, d), "self_attn.k_proj": (kvd, d), "self_attn.v_proj": (kvd, d),
            "self_attn.o_proj": (d, hd), "mlp.gate_proj": (cfg.ff, d), "mlp.up_proj": (cfg.ff, d), "mlp.down_proj": (d, cfg.ff)}


def lora_specs(cfg: LoraCfg) -> List[Tuple[str, Tuple[int, ...], float, float]]:
    a = 0.02 * math.sqrt(3.0)
    dims = _dims(cfg)
    s = []
    for i in range(cfg.layers):
        for lin in LINEARS:
            out, inn = dims[lin]
            s += [(f"language_model.model.layers.{i}.{lin}.lora_A", (cfg.lora_r, inn), a, 0.0),
                  (f"language_model.model.layers.{i}.{lin}.lora_B", (out, cfg.lora_r), a, 0.0)]
    return s


def make_weights(cfg: LoraCfg, seed: int):
    """(base weights = restate.make_weights, adapters); bit-identical to LlavaLoRADPOEngine.init_synthetic(seed)."""
    lora = {n: R.bf16_round(R.hash_uniform(int(np.prod(sh)), R.tensor_seed(n, seed), sc, sf)).reshape(sh)
            for n, sh, sc, sf in lora_specs(cfg)}
    return R.make_weights(cfg, seed), lora


def concatenated_forward(cfg: LoraCfg, w, lora, batch, loss_type: str = "sigmoid"):
    """base/trainer.py:190-242 with the adapters on (`lora` dict) or disabled (None)."""
    cb = R.concatenated_inputs(batch, -100, 0)
    n = batch["chosen_labels"].shape[0]
    logits, labels, imap = R.model_forward(cfg, w, cb["concatenated_input_ids"], cb["concatenated_attention_mask"],
                                           cb["concatenated_labels"], lora=lora, lora_scale=cfg.lora_scale,
                                           **cb["concatenated_img_input_dict"])
    logps = R.get_batch_logps(logits, labels, mask_shared_tokens=(loss_type == "ddpo"))
    return logps[:n], logps[n:], logits[:n], logits[n:], imap, labels


def get_batch_loss_metrics(cfg: LoraCfg, w, lora, batch, beta: float = 0.1, loss_type: str = "sigmoid"):
    """trl 0.8.1 get_batch_loss_metrics for a peft policy: reference pass = the same base with the adapters disabled."""
    pc, pr, pcl, prl, _, _ = concatenated_forward(cfg, w, lora, batch, loss_type)
    with torch.no_grad():
        rc, rr, _, _, _, _ = concatenated_forward(cfg, w, None, batch, loss_type)
    losses, cr, rj = R.dpo_loss(pc, pr, rc, rr, beta, 0.0, loss_type, False)
    metrics = {"rewards/chosen": cr.mean(), "rewards/rejected": rj.mean(), "rewards/accuracies": (cr > rj).float().mean(),
               "rewards/margins": (cr - rj).mean(), "logps/rejected": pr.detach().mean(), "logps/chosen": pc.detach().mean(),
               "logits/rejected": prl.detach().mean(), "logits/chosen": pcl.detach().mean()}
    return losses.mean(), metrics, dict(policy_chosen_logps=pc, policy_rejected_logps=pr, reference_chosen_logps=rc,
                                        reference_rejected_logps=rr, losses=losses, chosen_rewards=cr, rejected_rewards=rj)
