"""Mint golden vectors by EXECUTING THE REFERENCE'S OWN FUNCTIONS (build container only).

    python -m oracle.make_fixtures            # G1-G5 (seconds)
    python -m oracle.make_fixtures --config1  # BASELINE.json configs[0]: 7B shapes, fp32 CPU (minutes, ~30 GB)

Outputs: tests/golden/*.npz.  Everything here comes from /root/reference/src code
(VLDPOTrainer.get_batch_logps / dpo_loss, diff_lib.get_diff_ids, LlavaForRL.forward) except
the trl-0.8.1 glue (concatenated_inputs / get_batch_loss_metrics), which is not on disk.
Inputs are regenerated from seeds by `oracle.restate` (hash-based, device independent).
"""
from __future__ import annotations

import argparse
import os
import sys
import time
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim, restate as R  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def hf_config(cfg: R.LlavaCfg):
    from transformers import CLIPVisionConfig, LlamaConfig, LlavaConfig
    v = CLIPVisionConfig(hidden_size=cfg.v_hidden, intermediate_size=cfg.v_ff, num_hidden_layers=cfg.v_layers,
                         num_attention_heads=cfg.v_heads, image_size=cfg.image_size, patch_size=cfg.patch_size,
                         hidden_act="quick_gelu", layer_norm_eps=cfg.v_eps, projection_dim=cfg.v_hidden)
    t = LlamaConfig(vocab_size=cfg.vocab, hidden_size=cfg.hidden, intermediate_size=cfg.ff,
                    num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads, num_key_value_heads=cfg.kv_heads,
                    rms_norm_eps=cfg.rms_eps, rope_theta=cfg.rope_theta, max_position_embeddings=4096,
                    attention_bias=False, mlp_bias=False, tie_word_embeddings=False)
    c = LlavaConfig(vision_config=v, text_config=t, image_token_index=cfg.image_token_index,
                    projector_hidden_act="gelu", vision_feature_select_strategy="default",
                    vision_feature_layer=cfg.vision_feature_layer, tie_word_embeddings=False)
    c.pad_token_id = cfg.pad_token_id
    c.ignore_index = cfg.ignore_index
    c._attn_implementation = "eager"
    return c


def hf_config_next(cfg: R.LlavaCfg):
    """llava-v1.6-mistral-7b-hf style config: CLIP tower + MistralConfig decoder (no sliding window) + anyres pinpoints."""
    from transformers import CLIPVisionConfig, LlavaNextConfig, MistralConfig
    v = CLIPVisionConfig(hidden_size=cfg.v_hidden, intermediate_size=cfg.v_ff, num_hidden_layers=cfg.v_layers,
                         num_attention_heads=cfg.v_heads, image_size=cfg.image_size, patch_size=cfg.patch_size,
                         hidden_act="quick_gelu", layer_norm_eps=cfg.v_eps, projection_dim=cfg.v_hidden)
    t = MistralConfig(vocab_size=cfg.vocab, hidden_size=cfg.hidden, intermediate_size=cfg.ff,
                      num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads, num_key_value_heads=cfg.kv_heads,
                      rms_norm_eps=cfg.rms_eps, rope_theta=cfg.rope_theta, max_position_embeddings=4096,
                      sliding_window=None, tie_word_embeddings=False, head_dim=cfg.head_dim)
    c = LlavaNextConfig(vision_config=v, text_config=t, image_token_index=cfg.image_token_index,
                        projector_hidden_act="gelu", vision_feature_select_strategy="default",
                        vision_feature_layer=cfg.vision_feature_layer, tie_word_embeddings=False,
                        image_grid_pinpoints=[list(p) for p in cfg.image_grid_pinpoints])
    c.pad_token_id = cfg.pad_token_id
    c.ignore_index = cfg.ignore_index
    c._attn_implementation = "eager"
    return c


def hf_name(name: str) -> str:
    """transformers-4.41 parameter name -> transformers-5.5 LlavaForConditionalGeneration name."""
    if name == "language_model.lm_head.weight":
        return "lm_head.weight"
    if name.startswith("language_model.model."):
        return "model.language_model." + name[len("language_model.model."):]
    return "model." + name


def streamed_weights(cfg: R.LlavaCfg, seed: int, which: str):
    """Yield (name, tensor) one tensor at a time (7B shapes: never hold two full copies in RAM).
    Same values as R.make_policy_and_ref(cfg, seed)[0 if which == 'policy' else 1]."""
    for name, shape, scale, shift in R.weight_specs(cfg):
        n = int(np.prod(shape))
        t = R.bf16_round(R.hash_uniform(n, R.tensor_seed(name, seed), scale, shift))
        if which == "ref" and not name.startswith("vision_tower."):
            o = R.bf16_round(R.hash_uniform(n, R.tensor_seed(name, seed + 1), scale, shift))
            t = R.bf16_round(t + 0.05 * (o - (1.0 if name.endswith("norm.weight") else 0.0)))
        yield name, t.reshape(shape)


def build_reference_model(cfg: R.LlavaCfg, weights):
    """weights: dict name->tensor, or an iterator of (name, tensor)."""
    _, _, LlavaShim = ref_shim.reference_symbols()
    if cfg.family == "llava_next":
        from oracle._llavanext_shim import LlavaNextShim as LlavaShim  # noqa: F811
        hc = hf_config_next(cfg)
    else:
        hc = hf_config(cfg)
    with torch.device("meta"):
        m = LlavaShim(hc)
    m = m.to_empty(device="cpu")
    sd = m.state_dict()
    missing = []
    names = []
    with torch.no_grad():
        for n, t in (weights.items() if isinstance(weights, dict) else weights):
            names.append(n)
            k = hf_name(n)
            if k not in sd:
                missing.append(k)
                continue
            sd[k].copy_(t)
            del t
    assert not missing, missing[:5]
    loaded = {hf_name(n) for n in names}
    not_set = [k for k in sd if k not in loaded and "post_layernorm" not in k]
    assert not not_set, not_set[:5]
    for k in sd:  # post_layernorm is unused by the path (hidden_states[-2]); make it finite
        if "post_layernorm" in k:
            sd[k].fill_(1.0 if k.endswith("weight") else 0.0)
    # non-persistent buffers (position_ids, rotary inv_freq) are lost by to_empty(): rebuild them
    emb = m.model.vision_tower.vision_model.embeddings
    emb.position_ids = torch.arange(emb.num_positions).expand((1, -1))
    rot = m.model.language_model.rotary_emb
    inv, scale = rot.compute_default_rope_parameters(rot.config, "cpu") if hasattr(rot, "compute_default_rope_parameters") \
        else rot.rope_init_fn(rot.config, "cpu")
    rot.inv_freq = inv
    rot.original_inv_freq = inv
    rot.attention_scaling = scale
    m.eval()
    return m


def reference_concatenated_forward(model, cfg, batch, loss_type="sigmoid"):
    """The reference's VLDPOTrainer.concatenated_forward body (base/trainer.py:204-242) driven with
    the restated trl concatenated_inputs; model(...) and get_batch_logps are the reference's own."""
    VLDPOTrainer, _, _ = ref_shim.reference_symbols()
    cb = R.concatenated_inputs(batch, -100, 0)
    with torch.no_grad():
        out = model(input_ids=cb["concatenated_input_ids"], attention_mask=cb["concatenated_attention_mask"],
                    labels=cb["concatenated_labels"], use_cache=False, **cb["concatenated_img_input_dict"])
    logps = VLDPOTrainer.get_batch_logps(out.logits, out.labels, average_log_prob=False, is_encoder_decoder=False,
                                         label_pad_token_id=-100, mask_shared_tokens=(loss_type == "ddpo"))
    return logps, out


def ref_dpo_loss(pc, pr, rc, rr, beta, ls, loss_type, reference_free=False):
    VLDPOTrainer, _, _ = ref_shim.reference_symbols()
    self = SimpleNamespace(beta=beta, label_smoothing=ls, loss_type=loss_type, reference_free=reference_free,
                           accelerator=SimpleNamespace(device="cpu"))
    return VLDPOTrainer.dpo_loss(self, pc, pr, rc, rr)


def g1_logps():
    VLDPOTrainer, _, _ = ref_shim.reference_symbols()
    out = {}
    g = torch.Generator().manual_seed(1)
    for tag, (B2, S, V) in {"a": (4, 37, 320), "b": (2, 200, 2048), "c": (2, 12, 32064)}.items():
        logits = torch.randn(B2, S, V, generator=g) * 3.0
        labels = torch.randint(0, V, (B2, S), generator=g)
        for b in range(B2):
            p = int(torch.randint(1, S // 2, (1,), generator=g))
            e = int(torch.randint(S // 2 + 1, S + 1, (1,), generator=g))
            labels[b, :p] = -100
            labels[b, e:] = -100
        out[f"{tag}_logits"] = logits.numpy()
        out[f"{tag}_labels"] = labels.numpy()
        out[f"{tag}_sum"] = VLDPOTrainer.get_batch_logps(logits, labels).numpy()
        out[f"{tag}_avg"] = VLDPOTrainer.get_batch_logps(logits, labels, average_log_prob=True).numpy()
        bl = logits.to(torch.bfloat16)
        out[f"{tag}_sum_bf16in_refdtype"] = VLDPOTrainer.get_batch_logps(bl, labels).float().numpy()
        out[f"{tag}_sum_bf16in_fp32math"] = VLDPOTrainer.get_batch_logps(bl.float(), labels).numpy()
    np.savez_compressed(os.path.join(GOLDEN, "g1_logps.npz"), **out)


def g2_loss():
    out = {}
    g = torch.Generator().manual_seed(2)
    pc, pr, rc, rr = [(-torch.rand(5, generator=g) * 400 - 20) for _ in range(4)]
    rc = pc + torch.randn(5, generator=g) * 4
    rr = pr + torch.randn(5, generator=g) * 4
    out.update(pc=pc.numpy(), pr=pr.numpy(), rc=rc.numpy(), rr=rr.numpy())
    for lt in ("sigmoid", "ddpo", "hinge", "ipo", "kto_pair"):
        for ls in (0.0, 0.1):
            for rf in (False, True):
                l, c, r = ref_dpo_loss(pc, pr, rc, rr, 0.1, ls, lt, rf)
                k = f"{lt}_ls{ls}_rf{int(rf)}"
                out[k + "_losses"], out[k + "_cr"], out[k + "_rr"] = l.numpy(), c.numpy(), r.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "g2_loss.npz"), **out)


def g3_ddpo():
    _, get_diff_ids, _ = ref_shim.reference_symbols()
    rs = np.random.RandomState(3)
    cases = {}
    # (a) single substitution inside shared text
    a = rs.randint(3, 30000, size=1598).tolist()
    b = list(a)
    b[740:745] = rs.randint(3, 30000, size=3).tolist()
    cases["subst"] = (a, b)
    # (b) repetitive response triggering autojunk (len >= 200, a token > 1% + 1 times)
    a = ([5, 6, 7, 8] * 250)[:1000]
    a[100:103] = [11, 12, 13]
    b = list(a)
    b[900:905] = [21, 22]
    cases["autojunk"] = (a, b)
    # (c) identical
    a = rs.randint(3, 30000, size=300).tolist()
    cases["identical"] = (a, list(a))
    # (d) pure insertion (not counted: both-sides-non-empty rule)
    b = list(a)
    b[150:150] = [7, 7, 7, 7]
    cases["insertion"] = (a, b)
    # (e) masked-label style: zeros prefix + zeros suffix (labels==-100 rewritten to 0), several edits
    a = [0] * 600 + rs.randint(3, 30000, size=700).tolist() + [0] * 298
    b = list(a)
    for s in (650, 800, 1000, 1200):
        b[s:s + 4] = rs.randint(3, 30000, size=rs.randint(1, 8)).tolist()
    b = (b + [0] * 1598)[:1598]
    cases["masked"] = (a, b)
    # (f) short blocks below min_match_size
    a = rs.randint(3, 50, size=120).tolist()
    b = rs.randint(3, 50, size=110).tolist()
    cases["noisy_small_vocab"] = (a, b)
    # (g) random small-alphabet long sequences (autojunk + many blocks)
    a = rs.randint(3, 40, size=900).tolist()
    b = list(a)
    for s in range(50, 850, 90):
        b[s:s + 5] = rs.randint(3, 40, size=4).tolist()
    cases["small_alpha_long"] = (a, b)
    out = {}
    for k, (a, b) in cases.items():
        ia, ib = get_diff_ids(a, b, min_match_size=3)
        out[k + "_a"], out[k + "_b"] = np.array(a, dtype=np.int64), np.array(b, dtype=np.int64)
        out[k + "_ia"], out[k + "_ib"] = np.array(ia, dtype=np.int64), np.array(ib, dtype=np.int64)
    np.savez_compressed(os.path.join(GOLDEN, "g3_ddpo.npz"), **out)


def g45_llava(tag: str, cfg: R.LlavaCfg, n_pairs: int, text_len: int, prompt_len: int, seed: int, ddpo: bool,
              image_sizes=None):
    batch = R.make_batch(cfg, n_pairs, text_len, prompt_len, seed, ddpo_like=ddpo, image_sizes=image_sizes)
    out = {"seed": seed, "n_pairs": n_pairs, "text_len": text_len, "prompt_len": prompt_len}
    if image_sizes is not None:
        out["image_sizes"] = np.asarray(image_sizes, dtype=np.int64)
    res = {}
    for who in ("policy", "ref"):
        t0 = time.time()
        m = build_reference_model(cfg, streamed_weights(cfg, seed, who))
        print(f"[{tag}] {who} weights ready {time.time() - t0:.1f}s", flush=True)
        logps, o = reference_concatenated_forward(m, cfg, batch, "sigmoid")
        res[who] = logps
        out[f"{who}_logps"] = logps.numpy()
        if ddpo:  # same forward, the reference's get_batch_logps with mask_shared_tokens=True (trainer.py:169-184)
            VLDPOTrainer, _, _ = ref_shim.reference_symbols()
            dl = VLDPOTrainer.get_batch_logps(o.logits, o.labels, average_log_prob=False, is_encoder_decoder=False,
                                              label_pad_token_id=-100, mask_shared_tokens=True)
            out[f"{who}_logps_ddpo"] = dl.numpy()
            res[who + "_ddpo"] = dl
        if who == "policy":
            out["labels"] = o.labels.numpy()
            out["image_position_map"] = o.image_position_map.numpy()
            if o.logits.numel() < 1_000_000:
                out["policy_logits"] = o.logits.numpy()
            out["policy_logits_mean_chosen"] = o.logits[:n_pairs].mean().numpy()
            out["policy_logits_mean_rejected"] = o.logits[n_pairs:].mean().numpy()
        print(f"[{tag}] {who} forward {time.time() - t0:.1f}s", flush=True)
        del m, o
    n = n_pairs
    for lt in ("sigmoid", "ipo", "hinge", "kto_pair") + (("ddpo",) if ddpo else ()):
        sfx = "_ddpo" if lt == "ddpo" else ""
        pl, rl = res["policy" + sfx], res["ref" + sfx]
        l, c, r = ref_dpo_loss(pl[:n], pl[n:], rl[:n], rl[n:], 0.1, 0.0, lt)
        out[f"{lt}_losses"], out[f"{lt}_cr"], out[f"{lt}_rr"] = l.numpy(), c.numpy(), r.numpy()
    np.savez_compressed(os.path.join(GOLDEN, f"{tag}.npz"), **out)
    print(tag, {k: v for k, v in out.items() if k.endswith("logps") or k == "sigmoid_losses"}, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config1", action="store_true")
    ap.add_argument("--config1-bf16", dest="config1_bf16", action="store_true")
    ap.add_argument("--next", action="store_true", help="only the LLaVA-Next fixtures (g6_*)")
    ap.add_argument("--config4", action="store_true", help="LLaVA-Next-Mistral-7B shapes, DDPO (BASELINE.json configs[3])")
    ap.add_argument("--config2", action="store_true", help="LLaVA-1.5-7B at the FULL headline length: 1 pair, text 1024 -> 1599 merged")
    ap.add_argument("--config3", action="store_true", help="Qwen-VL-Chat 7B shapes + LoRA r 64 (BASELINE.json configs[2])")
    ap.add_argument("--config5", action="store_true", help="InternLM-XComposer2-VL-7B shapes + PLoRA/LoRA, KTO (configs[4])")
    ap.add_argument("--preprocess", action="store_true", help="only the CLIP image-preprocessing fixture (g7)")
    ap.add_argument("--qwen", action="store_true", help="only the Qwen-VL + LoRA fixtures (g9_*)")
    ap.add_argument("--xc2", action="store_true", help="only the InternLM-XComposer2 + PLoRA/LoRA fixtures (g10_*)")
    ap.add_argument("--lora", action="store_true", help="only the LLaVA / LLaVA-Next + LoRA fixtures (g11_*)")
    args = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    if args.config1_bf16:
        config1_reference_bf16()
        return
    if args.config1:
        # BASELINE.json configs[0]: LLaVA-1.5-7B shapes, 2 pairs, text 128 (+575 -> 703), fp32 CPU
        g45_llava("g5_config1_7b", R.LLAVA15_7B, 2, 128, 32, 0, ddpo=False)
        return
    if args.next:
        g6_next()
        return
    if args.config4:
        # BASELINE.json configs[3] at parity size: Mistral-7B decoder + CLIP-L/336 anyres, 1 pair, text 96, one wide
        # image (1x2 grid -> 3 crops, unpadded), DDPO token weights, fp32 CPU
        g45_llava("g8_config4_next7b", R.LLAVANEXT_MISTRAL_7B, 1, 96, 24, 0, ddpo=True, image_sizes=[(400, 640)])
        return
    if args.config2:
        # BASELINE.json configs[1] at its full sequence length (the headline bench shape): ONE pair, text 1024 (+575 image
        # rows -> 1599 merged), prompt 128, fp32 CPU -- oracle values for the shape the bench runs, not only properties
        g45_llava("g14_config2_full_7b", R.LLAVA15_7B, 1, 1024, 128, 0, ddpo=False)
        return
    if args.config3:
        g12_qwen7b()
        return
    if args.config5:
        g13_xc2_7b()
        return
    if args.preprocess:
        g7_clip_preprocess()
        return
    if args.qwen:
        g9_qwen()
        return
    if args.xc2:
        g10_xc2()
        return
    if args.lora:
        g11_lora()
        return
    g1_logps()
    g2_loss()
    g3_ddpo()
    g45_llava("g4_tiny", R.TINY, 2, 24, 8, 0, ddpo=True)
    g45_llava("g4_small", R.SMALL, 2, 96, 24, 0, ddpo=True)
    g6_next()
    g7_clip_preprocess()


from oracle.image_restate import G7_SIZES, synthetic_image  # noqa: E402


def g7_clip_preprocess():
    """What LlavaDPODataCollatorWithPadding (models/Llava/__init__.py:435-443) gets from
    `processor.image_processor(images=imgs, return_tensors="pt")`: transformers' PIL-backend CLIP processor (the
    4.41 slow processor's code path: Pillow bicubic resize, numpy crop / rescale / normalize) at the llava-1.5
    settings (shortest_edge 336, crop 336, OpenAI CLIP mean/std), run here on seeded synthetic RGB images.
    Stored: a per-image digest of the float32 output plus the full tensor for two images (keeps the file small)."""
    import hashlib
    from PIL import Image
    from transformers.models.clip.image_processing_pil_clip import CLIPImageProcessorPil
    proc = CLIPImageProcessorPil(size={"shortest_edge": 336}, crop_size={"height": 336, "width": 336})
    out = {"sizes": np.array(G7_SIZES, dtype=np.int64)}
    for i, (h, w) in enumerate(G7_SIZES):
        img = synthetic_image(h, w, i)
        pv = proc(images=[Image.fromarray(img)], return_tensors="np")["pixel_values"][0]
        assert pv.dtype == np.float32 and pv.shape == (3, 336, 336)
        out[f"sha256_{i}"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(pv).tobytes()).digest(), dtype=np.uint8)
        out[f"sum_{i}"] = np.float64(pv.astype(np.float64).sum())
        if i in (0, 4):
            out[f"pixel_values_{i}"] = pv
            nw, nh = (int(336 * w / h), 336) if h <= w else (336, int(336 * h / w))
            out[f"resized_u8_{i}"] = np.array(Image.fromarray(img).resize((nw, nh), Image.BICUBIC))
    np.savez_compressed(os.path.join(GOLDEN, "g7_clip_preprocess.npz"), **out)
    print("g7_clip_preprocess", {k: v.shape for k, v in out.items() if k.startswith("pixel")})


class _LoraLinear(torch.nn.Module):
    """peft.tuners.lora.Linear.forward restated (peft is not installed): result = base(x) + lora_B(lora_A(x)) * scaling
    (dropout is disabled by DPOTrainer's disable_dropout=True)."""

    def __init__(self, base: torch.nn.Linear, A: torch.Tensor, B: torch.Tensor, scaling: float):
        super().__init__()
        self.base, self.scaling = base, scaling
        self.A, self.B = torch.nn.Parameter(A.clone()), torch.nn.Parameter(B.clone())

    def forward(self, x):
        return self.base(x) + torch.nn.functional.linear(torch.nn.functional.linear(x, self.A), self.B) * self.scaling


def build_reference_qwen(qcfg, base_w):
    """The reference's vendored QWenLMHeadModel (models/QwenVL/modeling_qwen.py) on a small config, fp32."""
    ref_shim.install()
    from vlrlhf.models.QwenVL.configuration_qwen import QWenConfig
    from vlrlhf.models.QwenVL import QwenVLForRL
    hc = QWenConfig(vocab_size=qcfg.vocab, hidden_size=qcfg.hidden, num_hidden_layers=qcfg.layers,
                    num_attention_heads=qcfg.heads, kv_channels=qcfg.head_dim, intermediate_size=2 * qcfg.ff,
                    seq_length=2048, max_position_embeddings=2048, bf16=False, fp16=False, fp32=True, use_flash_attn=False,
                    use_dynamic_ntk=True, use_logn_attn=True, rotary_emb_base=qcfg.rope_theta,
                    layer_norm_epsilon=qcfg.rms_eps, no_bias=True,
                    visual=dict(heads=qcfg.v_heads, image_size=qcfg.image_size, image_start_id=qcfg.image_start_id,
                                layers=qcfg.v_layers, mlp_ratio=qcfg.v_mlp / qcfg.v_width, output_dim=qcfg.hidden,
                                patch_size=qcfg.patch_size, width=qcfg.v_width, n_queries=qcfg.n_queries))
    m = QwenVLForRL(hc)
    sd = m.state_dict()
    seen = set()
    with torch.no_grad():   # base_w: dict, or an iterator of (name, tensor) (7B shapes: one tensor in flight)
        for k, v in (base_w.items() if isinstance(base_w, dict) else base_w):
            assert k in sd and sd[k].shape == v.shape, k
            sd[k].copy_(v)
            seen.add(k)
            del v
    assert set(sd) == seen, (sorted(set(sd) ^ seen)[:8])
        # Resampler.pos_embed is a non-trainable sincos table built in __init__ (visual.py:112-114); keep it
    # transformers 5.x dropped ModuleUtilsMixin.get_head_mask; 4.41 returns [None] * n for head_mask=None
    m.transformer.get_head_mask = lambda head_mask, n, *a, **k: [None] * n
    m.eval()
    return m


def g9_qwen():
    """Qwen-VL + LoRA (BASELINE.json configs[2] at parity size): reference QwenVLForRL.forward with the adapters on
    (policy) and off (reference), VLDPOTrainer.get_batch_logps / dpo_loss on top."""
    from oracle import qwen_restate as Q
    _qwen_cases((("g9_qwen_tiny", Q.TINY_QWEN, 2, 48, 24), ("g9_qwen_small", Q.SMALL_QWEN, 2, 128, 72)))


def g12_qwen7b():
    """BASELINE.json configs[2] at 7B SHAPES (Qwen-VL-Chat: ViT-bigG 48 x 1664 with 104-wide heads, 256-query resampler,
    32 x 4096 LM, V = 151936, LoRA r 64 alpha 16 on c_attn / attn.c_proj / w1 / w2): the reference's vendored QwenVLForRL in
    fp32 on the CPU, ONE pair, text 352 (prompt 272 incl. the 258-token image span), adapters off = reference pass."""
    from oracle import qwen_restate as Q
    _qwen_cases((("g12_config3_qwen7b", Q.QWEN_VL_CHAT, 1, 352, 272),), stream=True)


def _stream_specs(specs, seed):
    for n, sh, sc, sf in specs:
        yield n, R.bf16_round(R.hash_uniform(int(np.prod(sh)), R.tensor_seed(n, seed), sc, sf)).reshape(sh)


def _qwen_cases(cases, stream=False):
    from oracle import qwen_restate as Q
    VLDPOTrainer, _, _ = ref_shim.reference_symbols()
    for tag, qcfg, n_pairs, text_len, prompt_len in cases:
        seed = 0
        t0 = time.time()
        pos = {"transformer.visual.attn_pool.pos_embed": Q.sincos_2d(qcfg.hidden, int(qcfg.n_queries ** 0.5))}
        if stream:   # 9.6 G fp32 parameters: never hold a second copy
            import itertools
            lora_w = Q._make(Q.lora_specs(qcfg), seed)
            m = build_reference_qwen(qcfg, itertools.chain(_stream_specs(Q.weight_specs(qcfg), seed), pos.items()))
            print(f"[{tag}] weights ready {time.time() - t0:.1f}s", flush=True)
        else:
            base_w, lora_w = Q.make_weights(qcfg, seed)
            m = build_reference_qwen(qcfg, {k: v for k, v in base_w.items()} | pos)
        batch = Q.make_batch(qcfg, n_pairs, text_len, prompt_len, seed, ddpo_like=True)
        cb = R.concatenated_inputs(batch, -100, 0)
        pixels = cb["concatenated_img_input_dict"]["pixel_values"]
        m.transformer.visual.encode = lambda paths, _m=m, _p=pixels: _m.transformer.visual(_p)  # no disk reads here
        out = {"seed": seed, "n_pairs": n_pairs, "text_len": text_len, "prompt_len": prompt_len}
        res = {}
        for who in ("ref", "policy"):
            if who == "policy":  # adapters on
                for i, blk in enumerate(m.transformer.h):
                    for t in Q.LORA_TARGETS:
                        parent = blk.attn if t.startswith("attn.") else blk.mlp
                        name = t.split(".")[1]
                        setattr(parent, name, _LoraLinear(getattr(parent, name), lora_w[f"transformer.h.{i}.{t}.lora_A"],
                                                          lora_w[f"transformer.h.{i}.{t}.lora_B"], qcfg.lora_scale))
            with torch.no_grad():
                o = m(input_ids=cb["concatenated_input_ids"], attention_mask=cb["concatenated_attention_mask"],
                      use_cache=False, return_dict=True)
            logits = o.logits.float()
            for lt in ("sigmoid", "ddpo"):
                lp = VLDPOTrainer.get_batch_logps(logits, cb["concatenated_labels"], average_log_prob=False,
                                                  is_encoder_decoder=False, label_pad_token_id=-100,
                                                  mask_shared_tokens=(lt == "ddpo"))
                res[who + ("_ddpo" if lt == "ddpo" else "")] = lp
                out[f"{who}_logps" + ("_ddpo" if lt == "ddpo" else "")] = lp.numpy()
            if who == "policy":
                out["image_position_map"] = o.image_position_map.numpy()
                out["policy_logits_mean_chosen"] = logits[:n_pairs].mean().numpy()
                out["policy_logits_mean_rejected"] = logits[n_pairs:].mean().numpy()
                if logits.numel() < 2_000_000:
                    out["policy_logits"] = logits.numpy()
            print(f"[{tag}] {who} forward done {time.time() - t0:.1f}s", flush=True)
            del o, logits
        n = n_pairs
        for lt in ("sigmoid", "ipo", "hinge", "kto_pair", "ddpo"):
            sfx = "_ddpo" if lt == "ddpo" else ""
            pl, rl = res["policy" + sfx], res["ref" + sfx]
            l, c, r = ref_dpo_loss(pl[:n], pl[n:], rl[:n], rl[n:], 0.1, 0.0, lt)
            out[f"{lt}_losses"], out[f"{lt}_cr"], out[f"{lt}_rr"] = l.numpy(), c.numpy(), r.numpy()
        np.savez_compressed(os.path.join(GOLDEN, f"{tag}.npz"), **out)
        print(tag, {k: v for k, v in out.items() if k.endswith("logps") or k == "sigmoid_losses"}, flush=True)
        del m


class _LoraOverPLoRA(torch.nn.Module):
    """peft lora.Linear over the reference's PLoRA module: result = base_layer(x, im_mask) + lora_B(lora_A(x)) * scaling."""

    def __init__(self, base, A, B, scaling):
        super().__init__()
        self.base, self.scaling = base, scaling
        self.A, self.B = torch.nn.Parameter(A.clone()), torch.nn.Parameter(B.clone())

    def forward(self, x, im_mask=None):
        return self.base(x, im_mask) + torch.nn.functional.linear(torch.nn.functional.linear(x, self.A), self.B) * self.scaling


def build_reference_xc2(xcfg, base_w):
    """The reference's InternLMXC2ForRL (models/InternLMXC2) on a small config.  Its constructor hard-codes the CLIP-L/336
    tower (downloaded) and the 1024->4096 projector; same-structure small builders are patched in (build_mlp.py:6-28,37-91),
    everything that runs in forward is the reference's code."""
    ref_shim.install()
    import torch.nn as nn
    from transformers import CLIPVisionConfig, CLIPVisionModel
    from vlrlhf.models.InternLMXC2.configuration_internlm_xcomposer2 import InternLMXcomposer2Config
    import vlrlhf.models.InternLMXC2.modeling_internlm_xcomposer2 as M
    import vlrlhf.models.InternLMXC2.modeling_internlm2 as M2
    import vlrlhf.models.InternLMXC2.build_mlp as BM
    # the reference hard-codes PLoRA(r=256, alpha=256) in the attention / MLP constructors: scale the rank with the config
    orig_plora_init = BM.PLoRA.__init__

    def plora_init(self, *a, lora_r=8, lora_alpha=16, **k):
        orig_plora_init(self, *a, lora_r=xcfg.plora_r, lora_alpha=xcfg.plora_alpha, **{**k, "lora_dropout": 0.0})
    BM.PLoRA.__init__ = plora_init

    class Tower(nn.Module):
        def __init__(self):
            super().__init__()
            vc = CLIPVisionConfig(hidden_size=xcfg.v_hidden, intermediate_size=xcfg.v_ff, num_hidden_layers=xcfg.v_layers,
                                  num_attention_heads=xcfg.v_heads, image_size=xcfg.image_size, patch_size=xcfg.patch_size,
                                  hidden_act="quick_gelu", layer_norm_eps=xcfg.v_eps, projection_dim=xcfg.v_hidden)
            vc._attn_implementation = "eager"
            self.vision_tower = CLIPVisionModel(vc)
            self.select_layer, self.select_feature, self.is_loaded = -1, "patch", True
        feature_select = BM.CLIPVisionTower.feature_select
        forward = BM.CLIPVisionTower.forward
        dtype = property(lambda self: self.vision_tower.dtype)
        device = property(lambda self: self.vision_tower.device)

    M.build_vision_tower = lambda: Tower()
    M.build_vision_projector = lambda: nn.Sequential(nn.Linear(xcfg.v_hidden, xcfg.hidden), nn.GELU(),
                                                     nn.Linear(xcfg.hidden, xcfg.hidden))
    c = InternLMXcomposer2Config(vocab_size=xcfg.vocab, hidden_size=xcfg.hidden, intermediate_size=xcfg.ff,
                                 num_hidden_layers=xcfg.layers, num_attention_heads=xcfg.heads,
                                 num_key_value_heads=xcfg.kv_heads, rms_norm_eps=xcfg.rms_eps, rope_theta=xcfg.rope_theta,
                                 max_position_embeddings=4096, bias=False, pad_token_id=xcfg.pad_token_id)
    c._attn_implementation = "eager"
    c.rope_scaling = None       # transformers 5.x rewrites rope_scaling into {"rope_type": ...}; 4.x default is None
    c.max_length = 4096
    c.img_size = xcfg.image_size
    c.image_token_index, c.ignore_index = xcfg.image_token_index, xcfg.ignore_index
    from vlrlhf.models.InternLMXC2 import InternLMXC2ForRL
    m = InternLMXC2ForRL(c)
    BM.PLoRA.__init__ = orig_plora_init
    sd = m.state_dict()
    seen = set()
    with torch.no_grad():   # base_w: dict, or an iterator of (name, tensor) (7B shapes: one tensor in flight)
        for k, v in (base_w.items() if isinstance(base_w, dict) else base_w):
            assert k in sd, k
            assert sd[k].shape == v.shape, (k, sd[k].shape, v.shape)
            sd[k].copy_(v)
            seen.add(k)
            del v
        for k in sd:
            if k not in seen:
                assert "post_layernorm" in k or "position_ids" in k, k
                if "post_layernorm" in k:
                    sd[k].fill_(1.0 if k.endswith("weight") else 0.0)
    emb = m.vit.vision_tower.vision_model.embeddings
    emb.position_ids = torch.arange(emb.num_positions).expand((1, -1))
    m.eval()
    return m


def g10_xc2():
    """InternLM-XComposer2-VL + PLoRA + LoRA (BASELINE.json configs[4] at parity size): policy = adapters on, reference =
    adapters off (the frozen PLoRA image-token adapters stay on in both), DPO / DDPO / KTO-pair losses on top."""
    from oracle import xc2_restate as X
    _xc2_cases((("g10_xc2_tiny", X.TINY_XC2, 2, 24, 8), ("g10_xc2_small", X.SMALL_XC2, 2, 96, 24)))


def g13_xc2_7b():
    """BASELINE.json configs[4] at 7B SHAPES (internlm-xcomposer2-vl-7b: CLIP-L/14 at 490 px -> 1225 image rows, InternLM2-7B
    GQA 32/8 with the per-group interleaved wqkv, PLoRA r 256 on the image rows of all five linears, trainable LoRA r 64):
    the reference's InternLMXC2ForRL in fp32 on the CPU, ONE pair, text 96 (S = 1320), KTO-pair / DPO / DDPO losses."""
    from oracle import xc2_restate as X
    _xc2_cases((("g13_config5_xc2_7b", X.XC2_VL_7B, 1, 96, 24),), stream=True)


def _xc2_cases(cases, stream=False):
    from oracle import xc2_restate as X
    VLDPOTrainer, _, _ = ref_shim.reference_symbols()
    for tag, xcfg, n_pairs, text_len, prompt_len in cases:
        seed = 0
        t0 = time.time()
        if stream:
            lora_w = X._make(X.lora_specs(xcfg), seed)
            m = build_reference_xc2(xcfg, _stream_specs(X.weight_specs(xcfg), seed))
            print(f"[{tag}] weights ready {time.time() - t0:.1f}s", flush=True)
        else:
            base_w, lora_w = X.make_weights(xcfg, seed)
            m = build_reference_xc2(xcfg, base_w)
        batch = R.make_batch(xcfg, n_pairs, text_len, prompt_len, seed, ddpo_like=True)
        cb = R.concatenated_inputs(batch, -100, 0)
        out = {"seed": seed, "n_pairs": n_pairs, "text_len": text_len, "prompt_len": prompt_len}
        res = {}
        for who in ("ref", "policy"):
            if who == "policy":
                for i, layer in enumerate(m.model.layers):
                    for lin in X.LINEARS:
                        parent = layer.attention if lin.startswith("attention.") else layer.feed_forward
                        name = lin.split(".")[1]
                        setattr(parent, name, _LoraOverPLoRA(getattr(parent, name), lora_w[f"model.layers.{i}.{lin}.lora_A"],
                                                             lora_w[f"model.layers.{i}.{lin}.lora_B"], xcfg.lora_scale))
            with torch.no_grad():
                o = m(input_ids=cb["concatenated_input_ids"], attention_mask=cb["concatenated_attention_mask"],
                      labels=cb["concatenated_labels"], use_cache=False, return_dict=True,
                      **cb["concatenated_img_input_dict"])
            logits = o.logits.float()
            for lt in ("sigmoid", "ddpo"):
                lp = VLDPOTrainer.get_batch_logps(logits, o.labels, average_log_prob=False, is_encoder_decoder=False,
                                                  label_pad_token_id=-100, mask_shared_tokens=(lt == "ddpo"))
                res[who + ("_ddpo" if lt == "ddpo" else "")] = lp
                out[f"{who}_logps" + ("_ddpo" if lt == "ddpo" else "")] = lp.numpy()
            if who == "policy":
                out["labels"] = o.labels.numpy()
                out["image_position_map"] = o.image_position_map.numpy()
                out["policy_logits_mean_chosen"] = logits[:n_pairs].mean().numpy()
                out["policy_logits_mean_rejected"] = logits[n_pairs:].mean().numpy()
                if logits.numel() < 2_000_000:
                    out["policy_logits"] = logits.numpy()
        n = n_pairs
        for lt in ("sigmoid", "ipo", "hinge", "kto_pair", "ddpo"):
            sfx = "_ddpo" if lt == "ddpo" else ""
            pl, rl = res["policy" + sfx], res["ref" + sfx]
            l, c, r = ref_dpo_loss(pl[:n], pl[n:], rl[:n], rl[n:], 0.1, 0.0, lt)
            out[f"{lt}_losses"], out[f"{lt}_cr"], out[f"{lt}_rr"] = l.numpy(), c.numpy(), r.numpy()
        np.savez_compressed(os.path.join(GOLDEN, f"{tag}.npz"), **out)
        print(tag, {k: v for k, v in out.items() if k.endswith("logps") or k == "kto_pair_losses"}, flush=True)


def g11_lora():
    """LLaVA-1.5 / LLaVA-Next with peft-style LoRA on the decoder linears (what scripts/dpo_llava.sh, dpo_llavanext.sh,
    kto_*.sh, ddpo_*.sh train): the reference's LlavaForRL / LlavaNextForRL with the adapters applied by hand (peft is not
    installed; `_LoraLinear` = peft lora.Linear.forward) = policy, adapters off = reference (TRL's null_ref_context)."""
    from oracle import lora_restate as LR
    VLDPOTrainer, _, _ = ref_shim.reference_symbols()
    cases = (("g11_lora_tiny", LR.TINY_LORA, 2, 24, 8, None), ("g11_lora_small", LR.SMALL_LORA, 2, 96, 24, None),
             ("g11_next_lora_tiny", LR.TINY_NEXT_LORA, 3, 24, 8, [(28, 28), (20, 50), (60, 25)]),
             ("g11_next_lora_small", LR.SMALL_NEXT_LORA, 2, 96, 24, [(112, 112), (90, 300)]))
    for tag, cfg, n_pairs, text_len, prompt_len, image_sizes in cases:
        seed = 0
        base_w, lora_w = LR.make_weights(cfg, seed)
        m = build_reference_model(cfg, base_w)
        batch = R.make_batch(cfg, n_pairs, text_len, prompt_len, seed, ddpo_like=True, image_sizes=image_sizes)
        cb = R.concatenated_inputs(batch, -100, 0)
        out = {"seed": seed, "n_pairs": n_pairs, "text_len": text_len, "prompt_len": prompt_len}
        if image_sizes is not None:
            out["image_sizes"] = np.asarray(image_sizes, dtype=np.int64)
        res = {}
        for who in ("ref", "policy"):
            if who == "policy":  # adapters on
                for i, layer in enumerate(m.model.language_model.layers):
                    for lin in LR.LINEARS:
                        parent_name, name = lin.split(".")
                        parent = getattr(layer, parent_name)
                        key = f"language_model.model.layers.{i}.{lin}"
                        setattr(parent, name, _LoraLinear(getattr(parent, name), lora_w[key + ".lora_A"],
                                                          lora_w[key + ".lora_B"], cfg.lora_scale))
            with torch.no_grad():
                o = m(input_ids=cb["concatenated_input_ids"], attention_mask=cb["concatenated_attention_mask"],
                      labels=cb["concatenated_labels"], use_cache=False, **cb["concatenated_img_input_dict"])
            logits = o.logits.float()
            for lt in ("sigmoid", "ddpo"):
                lp = VLDPOTrainer.get_batch_logps(logits, o.labels, average_log_prob=False, is_encoder_decoder=False,
                                                  label_pad_token_id=-100, mask_shared_tokens=(lt == "ddpo"))
                res[who + ("_ddpo" if lt == "ddpo" else "")] = lp
                out[f"{who}_logps" + ("_ddpo" if lt == "ddpo" else "")] = lp.numpy()
            if who == "policy":
                out["labels"] = o.labels.numpy()
                out["image_position_map"] = o.image_position_map.numpy()
                out["policy_logits_mean_chosen"] = logits[:n_pairs].mean().numpy()
                out["policy_logits_mean_rejected"] = logits[n_pairs:].mean().numpy()
        n = n_pairs
        for lt in ("sigmoid", "ipo", "hinge", "kto_pair", "ddpo"):
            sfx = "_ddpo" if lt == "ddpo" else ""
            pl, rl = res["policy" + sfx], res["ref" + sfx]
            l, c, r = ref_dpo_loss(pl[:n], pl[n:], rl[:n], rl[n:], 0.1, 0.0, lt)
            out[f"{lt}_losses"], out[f"{lt}_cr"], out[f"{lt}_rr"] = l.numpy(), c.numpy(), r.numpy()
        np.savez_compressed(os.path.join(GOLDEN, f"{tag}.npz"), **out)
        print(tag, {k: v for k, v in out.items() if k.endswith("logps") or k == "sigmoid_losses"}, flush=True)


def g6_next():
    """LLaVA-Next (models/LlavaNext LlavaNextForRL through oracle/_llavanext_shim.py): anyres crops of mixed aspect
    ratios (square -> 2x2 grid, wide -> 1x2 / 1x3 with unpadding, tall -> 2x1), image_newline, GQA decoder, DDPO."""
    g45_llava("g6_next_tiny", R.TINY_NEXT, 3, 24, 8, 0, ddpo=True, image_sizes=[(28, 28), (20, 50), (60, 25)])
    g45_llava("g6_next_small", R.SMALL_NEXT, 2, 96, 24, 0, ddpo=True, image_sizes=[(112, 112), (90, 300)])




def config1_reference_bf16():
    """The reference's OWN bf16 path (model weights + activations bf16, as `torch_dtype=bfloat16` in
    utils/auto_load.py:510,535) on the config-1 inputs: its deviation from the fp32 run is the noise floor of the
    path itself and is stored next to the fp32 golden values."""
    cfg, seed = R.LLAVA15_7B, 0
    path = os.path.join(GOLDEN, "g5_config1_7b.npz")
    d = dict(np.load(path))
    batch = R.make_batch(cfg, int(d["n_pairs"]), int(d["text_len"]), int(d["prompt_len"]), seed)
    batch["img_input_dict"]["pixel_values"] = batch["img_input_dict"]["pixel_values"].to(torch.bfloat16)
    t0 = time.time()
    m = build_reference_model(cfg, streamed_weights(cfg, seed, "policy"))
    m = m.to(torch.bfloat16)
    print(f"[config1-bf16] model ready {time.time() - t0:.1f}s", flush=True)
    logps, o = reference_concatenated_forward(m, cfg, batch, "sigmoid")
    d["policy_logps_refdtype_bf16"] = logps.float().numpy()
    print("[config1-bf16] reference bf16 logps", d["policy_logps_refdtype_bf16"], "fp32", d["policy_logps"],
          f"{time.time() - t0:.1f}s", flush=True)
    np.savez_compressed(path, **d)



if __name__ == "__main__":
    main()
