"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- CPU restatement of the InternLM-XComposer2-VL variant of the hot
path (SURVEY.md §8 a12 / BASELINE.json configs[4]: KTO-pair / DPO with LoRA r=64 on the LM, frozen tower + projector).

Follows the reference's vendored model code:
  * models/InternLMXC2/__init__.py        InternLMXC2ForRL.forward :106-233, _merge_input_ids_with_image_features :28-104
                                          (the LLaVA-1.5 merge; its image map is the PLoRA row mask `im_mask`)
  * models/InternLMXC2/build_mlp.py       CLIPVisionTower :37-137 (CLIP-L/14 at 490 px: 35x35 position table, select_layer
                                          -1, patch features), build_vision_projector :14-28 (Linear-GELU-Linear),
                                          PLoRA.forward :194-203: res[im_mask] += Plora_B(Plora_A(x[im_mask])) * alpha/r
  * models/InternLMXC2/modeling_internlm2.py  InternLM2Attention.forward :299-385 (fused wqkv laid out per KV group as
                                          [q heads of the group | k | v], rotary on arange(S) -- position_ids only size the
                                          batch, :186-203), InternLM2MLP :206-224 (w2(silu(w1 x) * w3 x)),
                                          InternLM2DecoderLayer :509-570, InternLM2RMSNorm :77-90
  * peft LoRA (not on disk) on default_lora_target (:251-252): attention.wqkv, attention.wo, feed_forward.w1/w2/w3.
Pinned against the reference's InternLMXC2ForRL run here (tests/golden/g10_xc2_*.npz, make_fixtures.py --xc2); the
reference hard-codes the CLIP-L tower and 4096-wide projector in its constructor, so the fixture run swaps in
same-structure small builders (build_vision_tower / build_vision_projector) -- every forward line is the reference's.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import restate as R


@dataclass
class XC2Cfg(R.LlavaCfg):
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


# internlm-xcomposer2-vl-7b: CLIP-L/14 at 490 px (1225 patches, last layer), InternLM2-7B (GQA 32/8, ff 14336, vocab 92544)
XC2_VL_7B = XC2Cfg(image_size=490, vision_feature_layer=-1, hidden=4096, layers=32, heads=32, kv_heads=8, ff=14336,
                   vocab=92544, rms_eps=1e-5, rope_theta=1e6, image_token_index=92543, pad_token_id=2, family="xc2")
TINY_XC2 = XC2Cfg(image_size=70, patch_size=14, v_hidden=128, v_layers=2, v_heads=2, v_ff=256, vision_feature_layer=-1,
                  hidden=256, layers=2, heads=4, kv_heads=2, ff=512, vocab=512, rms_eps=1e-5, rope_theta=1e6,
                  image_token_index=500, pad_token_id=2, family="xc2", plora_r=32, plora_alpha=32.0, lora_r=16,
                  lora_alpha=16.0)
SMALL_XC2 = XC2Cfg(image_size=112, patch_size=14, v_hidden=256, v_layers=2, v_heads=4, v_ff=512, vision_feature_layer=-1,
                   hidden=512, layers=2, heads=4, kv_heads=2, ff=1024, vocab=2048, rms_eps=1e-5, rope_theta=1e6,
                   image_token_index=2000, pad_token_id=2, family="xc2", plora_r=64, plora_alpha=64.0, lora_r=16,
                   lora_alpha=16.0)

LINEARS = ("attention.wqkv", "attention.wo", "feed_forward.w1", "feed_forward.w3", "feed_forward.w2")


def _dims(cfg: XC2Cfg) -> Dict[str, Tuple[int, int]]:
    d, dh = cfg.hidden, cfg.head_dim
    return {"attention.wqkv": ((cfg.heads + 2 * cfg.kv_heads) * dh, d), "attention.wo": (d, cfg.heads * dh),
            "feed_forward.w1": (cfg.ff, d), "feed_forward.w3": (cfg.ff, d), "feed_forward.w2": (d, cfg.ff)}


def weight_specs(cfg: XC2Cfg) -> List[Tuple[str, Tuple[int, ...], float, float]]:
    a = 0.02 * math.sqrt(3.0)
    s: List[Tuple[str, Tuple[int, ...], float, float]] = []
    # CLIP tower under the reference's prefix (vit.vision_tower.vision_model.*), position table already 35x35 + CLS
    for name, shape, scale, shift in R.weight_specs(cfg):
        if name.startswith("vision_tower.vision_model."):
            s.append(("vit." + name, shape, scale, shift))
    s += [("vision_proj.0.weight", (cfg.hidden, cfg.v_hidden), a, 0.0), ("vision_proj.0.bias", (cfg.hidden,), 0.02, 0.0),
          ("vision_proj.2.weight", (cfg.hidden, cfg.hidden), a, 0.0), ("vision_proj.2.bias", (cfg.hidden,), 0.02, 0.0),
          ("model.tok_embeddings.weight", (cfg.vocab, cfg.hidden), a, 0.0)]
    dims = _dims(cfg)
    for i in range(cfg.layers):
        p = f"model.layers.{i}."
        s += [(p + "attention_norm.weight", (cfg.hidden,), 0.1, 1.0), (p + "ffn_norm.weight", (cfg.hidden,), 0.1, 1.0)]
        for lin in LINEARS:
            out, inn = dims[lin]
            s += [(p + lin + ".weight", (out, inn), a, 0.0), (p + lin + ".Plora_A.weight", (cfg.plora_r, inn), a, 0.0),
                  (p + lin + ".Plora_B.weight", (out, cfg.plora_r), a, 0.0)]
    s += [("model.norm.weight", (cfg.hidden,), 0.1, 1.0), ("output.weight", (cfg.vocab, cfg.hidden), 3.0 * a, 0.0)]
    return s


def lora_specs(cfg: XC2Cfg) -> List[Tuple[str, Tuple[int, ...], float, float]]:
    a = 0.02 * math.sqrt(3.0)
    dims = _dims(cfg)
    s = []
    for i in range(cfg.layers):
        for lin in LINEARS:
            out, inn = dims[lin]
            s += [(f"model.layers.{i}.{lin}.lora_A", (cfg.lora_r, inn), a, 0.0), (f"model.layers.{i}.{lin}.lora_B", (out, cfg.lora_r), a, 0.0)]
    return s


def _make(specs, seed):
    return {n: R.bf16_round(R.hash_uniform(int(np.prod(sh)), R.tensor_seed(n, seed), sc, sf)).reshape(sh)
            for n, sh, sc, sf in specs}


def make_weights(cfg: XC2Cfg, seed: int):
    return _make(weight_specs(cfg), seed), _make(lora_specs(cfg), seed)


def _plinear(cfg, x, w, lora, name, im_mask):
    """PLoRA.forward (+ the peft adapter that wraps it)."""
    y = F.linear(x, w[name + ".weight"])
    if im_mask is not None and bool(im_mask.any()):
        part = F.linear(F.linear(x[im_mask], w[name + ".Plora_A.weight"]), w[name + ".Plora_B.weight"]) * cfg.plora_scale
        y = y.clone()
        y[im_mask] = y[im_mask] + part
    if lora is not None:
        y = y + cfg.lora_scale * F.linear(F.linear(x, lora[name + ".lora_A"]), lora[name + ".lora_B"])
    return y


def decoder(cfg: XC2Cfg, w, lora, emb, attention_mask, im_mask):
    B, S, _ = emb.shape
    H, KV, dh = cfg.heads, cfg.kv_heads, cfg.head_dim
    n_rep = H // KV
    inv_freq = 1.0 / (cfg.rope_theta ** (torch.arange(0, dh, 2).float() / dh))
    freqs = torch.outer(torch.arange(S).float(), inv_freq)  # positions are arange(S) whatever position_ids says (:186-203)
    e = torch.cat((freqs, freqs), dim=-1)
    cos, sin = e.cos()[None, None], e.sin()[None, None]
    bias = torch.full((S, S), float("-inf")).triu(1)[None, None] + \
        torch.zeros(B, 1, 1, S).masked_fill(attention_mask[:, None, None, :] == 0, float("-inf"))
    x = emb
    for i in range(cfg.layers):
        p = f"model.layers.{i}."
        h = R.rms_norm(x, w[p + "attention_norm.weight"], cfg.rms_eps)
        qkv = _plinear(cfg, h, w, lora, p + "attention.wqkv", im_mask).view(B, S, KV, n_rep + 2, dh)
        q = qkv[..., :n_rep, :].reshape(B, S, H, dh).transpose(1, 2)
        k = qkv[..., -2, :].transpose(1, 2)
        v = qkv[..., -1, :].transpose(1, 2)
        q = q * cos + R.rotate_half(q) * sin
        k = k * cos + R.rotate_half(k) * sin
        k, v = k.repeat_interleave(n_rep, dim=1), v.repeat_interleave(n_rep, dim=1)
        att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh) + bias, dim=-1)
        o = (att @ v).transpose(1, 2).reshape(B, S, H * dh)
        x = x + _plinear(cfg, o, w, lora, p + "attention.wo", im_mask)
        h = R.rms_norm(x, w[p + "ffn_norm.weight"], cfg.rms_eps)
        g = F.silu(_plinear(cfg, h, w, lora, p + "feed_forward.w1", im_mask)) * _plinear(cfg, h, w, lora, p + "feed_forward.w3", im_mask)
        x = x + _plinear(cfg, g, w, lora, p + "feed_forward.w2", im_mask)
    x = R.rms_norm(x, w["model.norm.weight"], cfg.rms_eps)
    return F.linear(x, w["output.weight"]).float()


def forward(cfg: XC2Cfg, w, lora, input_ids, attention_mask, labels, pixel_values):
    """InternLMXC2ForRL.forward -> (logits, merged labels, im_mask)."""
    fake = torch.where(input_ids == cfg.image_token_index, torch.tensor(cfg.pad_token_id), input_ids)  # :130-132
    inputs_embeds = F.embedding(fake, w["model.tok_embeddings.weight"])
    vw = {k[len("vit."):]: v for k, v in w.items() if k.startswith("vit.")}
    feats = R.clip_vision_features(cfg, vw, pixel_values)[:, 1:]  # select_layer -1, "patch"
    img = F.linear(F.gelu(F.linear(feats, w["vision_proj.0.weight"], w["vision_proj.0.bias"])), w["vision_proj.2.weight"],
                   w["vision_proj.2.bias"])
    emb, mask, new_labels, _, im_mask = R.merge_input_ids_with_image_features(cfg, img, inputs_embeds, input_ids, attention_mask,
                                                                           labels)
    return decoder(cfg, w, lora, emb, mask, im_mask), new_labels, im_mask


def concatenated_forward(cfg: XC2Cfg, w, lora, batch, loss_type: str = "sigmoid"):
    cb = R.concatenated_inputs(batch, -100, 0)
    n = batch["chosen_labels"].shape[0]
    logits, labels, im_mask = forward(cfg, w, lora, cb["concatenated_input_ids"], cb["concatenated_attention_mask"],
                                      cb["concatenated_labels"], cb["concatenated_img_input_dict"]["pixel_values"])
    logps = R.get_batch_logps(logits, labels, mask_shared_tokens=(loss_type == "ddpo"))
    return logps[:n], logps[n:], logits[:n], logits[n:], im_mask, labels


def get_batch_loss_metrics(cfg: XC2Cfg, w, lora, batch, beta: float = 0.1, loss_type: str = "sigmoid"):
    pc, pr, pcl, prl, _, _ = concatenated_forward(cfg, w, lora, batch, loss_type)
    with torch.no_grad():
        rc, rr, _, _, _, _ = concatenated_forward(cfg, w, None, batch, loss_type)
    losses, cr, rj = R.dpo_loss(pc, pr, rc, rr, beta, 0.0, loss_type, False)
    metrics = {"rewards/chosen": cr.mean(), "rewards/rejected": rj.mean(), "rewards/accuracies": (cr > rj).float().mean(),
               "rewards/margins": (cr - rj).mean(), "logps/rejected": pr.detach().mean(), "logps/chosen": pc.detach().mean(),
               "logits/rejected": prl.detach().mean(), "logits/chosen": pcl.detach().mean()}
    return losses.mean(), metrics, dict(policy_chosen_logps=pc, policy_rejected_logps=pr, reference_chosen_logps=rc,
                                        reference_rejected_logps=rr, losses=losses, chosen_rewards=cr, rejected_rewards=rj)
