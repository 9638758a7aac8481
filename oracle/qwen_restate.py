"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- CPU restatement of the Qwen-VL variant of the hot path
(SURVEY.md §8 a12 / BASELINE.json configs[2]: Qwen-VL-Chat DPO, LoRA r=64 on the LM, frozen vision tower).

Follows the reference's vendored model code, each function citing it:
  * models/QwenVL/visual.py            VisionTransformer.forward :393-415, VisualAttention :186-241 (per-head interleaved
                                        q|k|v in_proj), VisualAttentionBlock :286-299, Resampler.forward :140-152,
                                        get_abs_pos :24-45 (bicubic interpolation of the position tables)
  * models/QwenVL/modeling_qwen.py     QWenModel.forward :509-699 (image features overwrite the 256 placeholder positions
                                        between <img> and </img>, positions = arange(S)), QWenAttention :282-306 + _attn
                                        :143-176, QWenMLP :314-326 (a1 * silu(a2)), QWenBlock :344-386, RMSNorm :1084-1099,
                                        RotaryEmbedding / apply_rotary_pos_emb :1031-1081, QWenLMHeadModel :800-855
  * peft LoraLayer (NOT on disk; published algorithm): y = W x + (alpha / r) * B (A x) on the target modules
    scripts/dpo_qwenvl.sh names (c_attn, attn.c_proj, w1, w2); the reference pass runs with the adapters disabled
    (trl null_ref_context), i.e. on the base weights.
  * base/trainer.py:190-301 through oracle.restate (concatenated_inputs, get_batch_logps, dpo_loss): Qwen's output has no
    `labels`, so the concatenated labels are used as they are (:225-229).
Pinned against the reference's own QWenLMHeadModel run here on a small config (tests/golden/g9_qwen_*.npz,
oracle/make_fixtures.py --qwen; the disk-reading `visual.encode` is replaced by the pixel tensors of the batch).
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
class QwenCfg:
    # visual (open_clip ViT-bigG/14 + resampler)
    image_size: int = 448
    patch_size: int = 14
    v_width: int = 1664
    v_layers: int = 48
    v_heads: int = 16
    v_mlp: int = 8192  # int(width * mlp_ratio 4.9231)
    n_queries: int = 256
    v_eps: float = 1e-6
    pos_table: int = 256  # visual.positional_embedding rows (16 x 16 grid), interpolated to the patch grid
    # language model
    hidden: int = 4096
    layers: int = 32
    heads: int = 32
    ff: int = 11008  # intermediate_size // 2
    vocab: int = 151936
    rms_eps: float = 1e-6
    rope_theta: float = 10000.0
    image_start_id: int = 151857  # <img>; +1 = </img>; +2 = <imgpad>
    pad_token_id: int = 151643
    ignore_index: int = -100
    # LoRA (scripts/dpo_qwenvl.sh: r 64, alpha 16, targets c_attn, attn.c_proj, w1, w2)
    lora_r: int = 64
    lora_alpha: float = 16.0

    @property
    def n_patches(self) -> int:
        return (self.image_size // self.patch_size) ** 2

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads

    @property
    def v_head_dim(self) -> int:
        return self.v_width // self.v_heads

    @property
    def r_heads(self) -> int:  # Resampler(num_heads=output_dim // 128)
        return self.hidden // 128

    @property
    def lora_scale(self) -> float:
        return self.lora_alpha / self.lora_r


QWEN_VL_CHAT = QwenCfg()
# head dims as in the real model (ViT 104 -> exercises the padded-head layout of the CUDA path; LM / resampler 128)
TINY_QWEN = QwenCfg(image_size=112, patch_size=14, v_width=208, v_layers=2, v_heads=2, v_mlp=416, n_queries=16,
                    hidden=256, layers=2, heads=2, ff=256, vocab=512, image_start_id=500, pad_token_id=499, lora_r=16,
                    lora_alpha=8.0)
SMALL_QWEN = QwenCfg(image_size=224, patch_size=14, v_width=416, v_layers=2, v_heads=4, v_mlp=1024, n_queries=64,
                     hidden=512, layers=2, heads=4, ff=1024, vocab=2048, image_start_id=2000, pad_token_id=1999, lora_r=16,
                     lora_alpha=8.0)

LORA_TARGETS = ("attn.c_attn", "attn.c_proj", "mlp.w1", "mlp.w2")


def weight_specs(cfg: QwenCfg) -> List[Tuple[str, Tuple[int, ...], float, float]]:
    """(reference state-dict name, shape, uniform half-width, shift) of the base model."""
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


def lora_specs(cfg: QwenCfg) -> List[Tuple[str, Tuple[int, ...], float, float]]:
    """LoRA tensors `<module>.lora_A` [r, in] / `.lora_B` [out, r].  peft initialises B = 0; the synthetic recipe uses a
    non-zero B so that the adapter path carries signal in the parity tests."""
    a = 0.02 * math.sqrt(3.0)
    d, r = cfg.hidden, cfg.lora_r
    outs = {"attn.c_attn": 3 * d, "attn.c_proj": d, "mlp.w1": cfg.ff, "mlp.w2": cfg.ff}
    s = []
    for i in range(cfg.layers):
        for t in LORA_TARGETS:
            s += [(f"transformer.h.{i}.{t}.lora_A", (r, d), a, 0.0), (f"transformer.h.{i}.{t}.lora_B", (outs[t], r), a, 0.0)]
    return s


def _make(specs, seed):
    return {n: R.bf16_round(R.hash_uniform(int(np.prod(sh)), R.tensor_seed(n, seed), sc, sf)).reshape(sh)
            for n, sh, sc, sf in specs}


def make_weights(cfg: QwenCfg, seed: int) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """-> (base weights, LoRA weights); bit-identical to what QwenVLDPOEngine.init_synthetic(seed) builds."""
    return _make(weight_specs(cfg), seed), _make(lora_specs(cfg), seed)


def sincos_2d(embed_dim: int, grid: int) -> torch.Tensor:
    """visual.py get_2d_sincos_pos_embed (:49-96), float32 [grid*grid, embed_dim]."""
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


def get_abs_pos(abs_pos: torch.Tensor, tgt: int) -> torch.Tensor:
    """visual.py:24-45."""
    src, t = int(math.sqrt(abs_pos.size(0))), int(math.sqrt(tgt))
    if src == t:
        return abs_pos
    return F.interpolate(abs_pos.float().reshape(1, src, src, -1).permute(0, 3, 1, 2), size=(t, t), mode="bicubic",
                         align_corners=False).permute(0, 2, 3, 1).flatten(0, 2).to(abs_pos.dtype)


def visual_forward(cfg: QwenCfg, w: Dict[str, torch.Tensor], pixels: torch.Tensor) -> torch.Tensor:
    """[B, 3, H, W] -> [B, n_queries, hidden] (visual.py:393-415)."""
    v = "transformer.visual."
    B = pixels.shape[0]
    x = F.conv2d(pixels, w[v + "conv1.weight"], stride=cfg.patch_size).flatten(2).permute(0, 2, 1)  # [B, P, width]
    x = x + get_abs_pos(w[v + "positional_embedding"], x.size(1))
    x = F.layer_norm(x, (cfg.v_width,), w[v + "ln_pre.weight"], w[v + "ln_pre.bias"], cfg.v_eps)
    H, hn = cfg.v_heads, cfg.v_head_dim
    P = x.shape[1]
    for i in range(cfg.v_layers):
        p = f"{v}transformer.resblocks.{i}."
        h = F.layer_norm(x, (cfg.v_width,), w[p + "ln_1.weight"], w[p + "ln_1.bias"], cfg.v_eps)
        mixed = F.linear(h, w[p + "attn.in_proj.weight"], w[p + "attn.in_proj.bias"]).view(B, P, H, 3 * hn)
        q, k, val = mixed.split(hn, dim=-1)  # per head: [q | k | v]
        att = torch.softmax((q.permute(0, 2, 1, 3) / math.sqrt(hn)) @ k.permute(0, 2, 3, 1), dim=-1)
        ctx = (att @ val.permute(0, 2, 1, 3)).permute(0, 2, 1, 3).reshape(B, P, cfg.v_width)
        x = x + F.linear(ctx, w[p + "attn.out_proj.weight"], w[p + "attn.out_proj.bias"])
        h = F.layer_norm(x, (cfg.v_width,), w[p + "ln_2.weight"], w[p + "ln_2.bias"], cfg.v_eps)
        h = F.gelu(F.linear(h, w[p + "mlp.c_fc.weight"], w[p + "mlp.c_fc.bias"]))
        x = x + F.linear(h, w[p + "mlp.c_proj.weight"], w[p + "mlp.c_proj.bias"])
    # Resampler (:140-152): one cross-attention from n_queries learned queries to the patch features
    p = v + "attn_pool."
    d = cfg.hidden
    qpos = sincos_2d(d, int(math.sqrt(cfg.n_queries)))
    kpos = get_abs_pos(qpos, P)
    kv = F.layer_norm(F.linear(x, w[p + "kv_proj.weight"]), (d,), w[p + "ln_kv.weight"], w[p + "ln_kv.bias"], cfg.v_eps)
    qin = F.layer_norm(w[p + "query"], (d,), w[p + "ln_q.weight"], w[p + "ln_q.bias"], cfg.v_eps) + qpos  # [Q, d]
    Wi, bi = w[p + "attn.in_proj_weight"], w[p + "attn.in_proj_bias"]
    Hh, dh = cfg.r_heads, d // cfg.r_heads
    q = F.linear(qin, Wi[:d], bi[:d]).view(1, cfg.n_queries, Hh, dh).expand(B, -1, -1, -1).permute(0, 2, 1, 3)
    k = F.linear(kv + kpos, Wi[d:2 * d], bi[d:2 * d]).view(B, P, Hh, dh).permute(0, 2, 1, 3)
    val = F.linear(kv, Wi[2 * d:], bi[2 * d:]).view(B, P, Hh, dh).permute(0, 2, 1, 3)
    att = torch.softmax((q / math.sqrt(dh)) @ k.transpose(-1, -2), dim=-1)
    o = (att @ val).permute(0, 2, 1, 3).reshape(B, cfg.n_queries, d)
    o = F.linear(o, w[p + "attn.out_proj.weight"], w[p + "attn.out_proj.bias"])
    o = F.layer_norm(o, (d,), w[v + "ln_post.weight"], w[v + "ln_post.bias"], cfg.v_eps)
    return o @ w[v + "proj"]


def rms_norm(x, weight, eps):
    return (x.float() * torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + eps)) * weight


def _lin(x, w, name, lora, scale, bias=None):
    y = F.linear(x, w[name + ".weight"], bias)
    if lora is not None and (name + ".lora_A") in lora:
        y = y + scale * F.linear(F.linear(x, lora[name + ".lora_A"]), lora[name + ".lora_B"])
    return y


def image_spans(cfg: QwenCfg, input_ids: torch.Tensor) -> torch.Tensor:
    """modeling_qwen.py:524-528 -> rows (sequence, a, b): <img> at a, </img> at b."""
    bos = torch.where(input_ids == cfg.image_start_id)
    eos = torch.where(input_ids == cfg.image_start_id + 1)
    assert (bos[0] == eos[0]).all()
    return torch.stack((bos[0], bos[1], eos[1]), dim=1)


def lm_forward(cfg: QwenCfg, w, lora, input_ids, attention_mask, images: Optional[torch.Tensor]):
    """QWenLMHeadModel.forward (training branch) -> (logits fp32 [B, S, V], image_position_map [B, S])."""
    B, S = input_ids.shape
    x = F.embedding(input_ids, w["transformer.wte.weight"]).clone()
    img_map = torch.zeros(B, S, dtype=torch.bool)
    if images is not None:
        for idx, (i, a, b) in enumerate(image_spans(cfg, input_ids).tolist()):
            x[i, a + 1:b] = images[idx]  # :618-621
            img_map[i, a + 1:b] = True
    H, dh = cfg.heads, cfg.head_dim
    inv_freq = 1.0 / (cfg.rope_theta ** (torch.arange(0, dh, 2).float() / dh))
    freqs = torch.outer(torch.arange(S).float(), inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    cos, sin = emb.cos()[None, None], emb.sin()[None, None]
    bias = torch.full((S, S), float("-inf")).triu(1)[None, None] + \
        torch.zeros(B, 1, 1, S).masked_fill(attention_mask[:, None, None, :] == 0, float("-inf"))
    sc = cfg.lora_scale
    for i in range(cfg.layers):
        p = f"transformer.h.{i}."
        h = rms_norm(x, w[p + "ln_1.weight"], cfg.rms_eps)
        qkv = _lin(h, w, p + "attn.c_attn", lora, sc, w[p + "attn.c_attn.bias"])
        q, k, v = (t.view(B, S, H, dh).transpose(1, 2) for t in qkv.split(cfg.hidden, dim=2))
        q = q * cos + R.rotate_half(q) * sin
        k = k * cos + R.rotate_half(k) * sin
        att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh) + bias, dim=-1)
        o = (att @ v).transpose(1, 2).reshape(B, S, cfg.hidden)
        x = x + _lin(o, w, p + "attn.c_proj", lora, sc)
        h = rms_norm(x, w[p + "ln_2.weight"], cfg.rms_eps)
        a1 = _lin(h, w, p + "mlp.w1", lora, sc)
        a2 = _lin(h, w, p + "mlp.w2", lora, sc)
        x = x + F.linear(a1 * F.silu(a2), w[p + "mlp.c_proj.weight"])
    x = rms_norm(x, w["transformer.ln_f.weight"], cfg.rms_eps)
    return F.linear(x, w["lm_head.weight"]).float(), img_map


def make_batch(cfg: QwenCfg, n_pairs: int, text_len: int, prompt_len: int, seed: int, ddpo_like: bool = False) -> Dict:
    """Collated DPO batch in the Qwen-VL format: the prompt holds <img> + n_queries placeholder tokens (<imgpad>; in the
    reference the first of them spell the image path) + </img>; images ride in img_input_dict.pixel_values."""
    g = np.random.RandomState(seed)
    lo, hi = 3, min(cfg.image_start_id, cfg.pad_token_id, cfg.vocab) - 1
    B, L = n_pairs, text_len
    img_block = 2 + cfg.n_queries
    assert prompt_len >= 1 + img_block and L > prompt_len + 4
    prompt = g.randint(lo, hi, size=(B, prompt_len))
    prompt[:, 0] = 1
    prompt[:, 1] = cfg.image_start_id
    prompt[:, 2:2 + cfg.n_queries] = cfg.image_start_id + 2
    prompt[:, 2 + cfg.n_queries] = cfg.image_start_id + 1
    chosen_len = np.full(B, L)
    rejected_len = g.randint(prompt_len + (L - prompt_len) * 3 // 4, L + 1, size=B)
    swap = g.rand(B) < 0.5
    chosen_len, rejected_len = np.where(swap, rejected_len, chosen_len), np.where(swap, chosen_len, rejected_len)
    base_resp = g.randint(lo, hi, size=(B, L))
    out: Dict = {}
    for key, lens in (("chosen", chosen_len), ("rejected", rejected_len)):
        ids = np.full((B, L), cfg.pad_token_id, dtype=np.int64)
        mask = np.zeros((B, L), dtype=np.int64)
        labels = np.full((B, L), -100, dtype=np.int64)
        for b in range(B):
            n = int(lens[b])
            resp = base_resp[b].copy() if ddpo_like else g.randint(lo, hi, size=L)
            if ddpo_like and key == "rejected":
                for _ in range(3):
                    s0 = g.randint(prompt_len, max(prompt_len + 1, n - 8))
                    resp[s0:s0 + g.randint(1, 6)] = g.randint(lo, hi)
            ids[b, :prompt_len] = prompt[b]
            ids[b, prompt_len:n] = resp[prompt_len:n]
            mask[b, :n] = 1
            labels[b, prompt_len:n] = ids[b, prompt_len:n]
        out[f"{key}_input_ids"] = torch.from_numpy(ids)
        out[f"{key}_attention_mask"] = torch.from_numpy(mask)
        out[f"{key}_labels"] = torch.from_numpy(labels)
    n_pix = B * 3 * cfg.image_size * cfg.image_size
    out["img_input_dict"] = {"pixel_values": R.bf16_round(R.hash_uniform(n_pix, R.tensor_seed("pixel_values", seed), 1.7320508)
                                                           ).reshape(B, 3, cfg.image_size, cfg.image_size)}
    return out


def concatenated_forward(cfg: QwenCfg, w, lora, batch, loss_type: str = "sigmoid"):
    """base/trainer.py:190-242 for a model whose output carries no `labels`."""
    cb = R.concatenated_inputs(batch, -100, 0)
    n = batch["chosen_labels"].shape[0]
    images = visual_forward(cfg, w, cb["concatenated_img_input_dict"]["pixel_values"])
    logits, img_map = lm_forward(cfg, w, lora, cb["concatenated_input_ids"], cb["concatenated_attention_mask"], images)
    logps = R.get_batch_logps(logits, cb["concatenated_labels"], mask_shared_tokens=(loss_type == "ddpo"))
    return logps[:n], logps[n:], logits[:n], logits[n:], img_map


def get_batch_loss_metrics(cfg: QwenCfg, w, lora, batch, beta: float = 0.1, loss_type: str = "sigmoid"):
    """policy = base + LoRA (grad), reference = base with the adapters disabled (no grad)."""
    pc, pr, pcl, prl, _ = concatenated_forward(cfg, w, lora, batch, loss_type)
    with torch.no_grad():
        rc, rr, _, _, _ = concatenated_forward(cfg, w, None, batch, loss_type)
    losses, cr, rj = R.dpo_loss(pc, pr, rc, rr, beta, 0.0, loss_type, False)
    metrics = {"rewards/chosen": cr.mean(), "rewards/rejected": rj.mean(), "rewards/accuracies": (cr > rj).float().mean(),
               "rewards/margins": (cr - rj).mean(), "logps/rejected": pr.detach().mean(), "logps/chosen": pc.detach().mean(),
               "logits/rejected": prl.detach().mean(), "logits/chosen": pcl.detach().mean()}
    return losses.mean(), metrics, dict(policy_chosen_logps=pc, policy_rejected_logps=pr, reference_chosen_logps=rc,
                                        reference_rejected_logps=rr, losses=losses, chosen_rewards=cr, rejected_rewards=rj)
