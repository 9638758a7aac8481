"""bench.py --impl library: the honest GPU baseline BASELINE.md §2.2 promised -- the STOCK library stack the reference
delegates every FLOP to, on the same B200, same shapes, same step:

    transformers' LlavaForConditionalGeneration (the class vlrlhf's LlavaForRL subclasses, models/Llava/__init__.py:23)
    in bf16 with attn_implementation="sdpa" and gradient checkpointing (every scripts/*.sh passes --gradient_checkpointing
    True), policy + frozen reference copy, VLDPOTrainer.get_batch_logps / dpo_loss as plain torch ops
    (base/trainer.py:148-188, 244-301), autograd backward, clip_grad_norm_(1.0), torch.optim.AdamW(fused=True).

Nothing of this repo's kernels, engine or plugin runs here (and nothing under oracle/): it is cuBLAS + SDPA + ATen.  The
reference's own control flow is kept where it costs time: chosen and rejected are concatenated on the batch axis, the image
tensor is duplicated [v, v] (base/trainer.py:135-145) so the vision tower runs on 2B images per pass, the full [2B, S, V]
logits are materialised, and the reference pass runs under no_grad.  transformers 5.x expects the <image> placeholder
already expanded to 576 tokens (its processor does that), so the synthetic ids carry 576 image tokens: S = 1599 as in the
reference's in-model merge.
"""
from __future__ import annotations

import json
import os
import time


def build_model(shapes, device, dtype):
    import torch
    from transformers import CLIPVisionConfig, LlamaConfig, LlavaConfig, LlavaForConditionalGeneration
    v = CLIPVisionConfig(hidden_size=shapes["v_hidden"], intermediate_size=shapes["v_ff"], num_hidden_layers=shapes["v_layers"],
                         num_attention_heads=shapes["v_heads"], image_size=shapes["image_size"], patch_size=shapes["patch_size"],
                         hidden_act="quick_gelu", projection_dim=shapes["v_hidden"])
    t = LlamaConfig(vocab_size=shapes["vocab"], hidden_size=shapes["hidden"], intermediate_size=shapes["ff"],
                    num_hidden_layers=shapes["layers"], num_attention_heads=shapes["heads"],
                    num_key_value_heads=shapes["kv_heads"], max_position_embeddings=4096, tie_word_embeddings=False)
    c = LlavaConfig(vision_config=v, text_config=t, image_token_index=shapes["image_token_index"], projector_hidden_act="gelu",
                    vision_feature_select_strategy="default", vision_feature_layer=-2, tie_word_embeddings=False)
    c._attn_implementation = "sdpa"
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        with torch.device(device):
            m = LlavaForConditionalGeneration(c)
    finally:
        torch.set_default_dtype(prev)
    return m.to(dtype)


def get_batch_logps(logits, labels, label_pad_token_id=-100):
    """VLDPOTrainer.get_batch_logps (base/trainer.py:148-188), sigmoid-loss branch."""
    import torch
    labels = labels[:, 1:].clone()
    logits = logits[:, :-1, :]
    loss_mask = labels != label_pad_token_id
    labels[labels == label_pad_token_id] = 0
    per_token = torch.gather(logits.log_softmax(-1), dim=2, index=labels.unsqueeze(2)).squeeze(2)
    return (per_token * loss_mask).sum(-1)


def dpo_loss(pc, pr, rc, rr, beta=0.1):
    """VLDPOTrainer.dpo_loss, loss_type "sigmoid", label_smoothing 0 (base/trainer.py:244-301)."""
    import torch.nn.functional as F
    logits = (pc - pr) - (rc - rr)
    losses = -F.logsigmoid(beta * logits)
    return losses, beta * (pc - rc).detach(), beta * (pr - rr).detach()


LLAVA15_7B = dict(v_hidden=1024, v_ff=4096, v_layers=24, v_heads=16, image_size=336, patch_size=14, vocab=32064, hidden=4096,
                  ff=11008, layers=32, heads=32, kv_heads=32, image_token_index=32000)
SMALL = dict(v_hidden=256, v_ff=512, v_layers=3, v_heads=4, image_size=112, patch_size=14, vocab=2048, hidden=512, ff=1024,
             layers=2, heads=4, kv_heads=4, image_token_index=2000)


def make_batch(shapes, n_pairs, text_len, prompt_len, seed, device):
    """Concatenated batch (chosen first) with the <image> placeholder expanded to n_patches tokens; right-padded, ragged
    rejected lengths like the repo's synthetic batch; labels = -100 on prompt, image and padding."""
    import torch
    g = torch.Generator().manual_seed(seed)
    P = (shapes["image_size"] // shapes["patch_size"]) ** 2
    S = text_len - 1 + P
    n = 2 * n_pairs
    ids = torch.randint(3, shapes["image_token_index"] - 1, (n, S), generator=g)
    ids[:, 0] = 1
    ids[:, 1:1 + P] = shapes["image_token_index"]
    lens = torch.full((n,), S)
    lens[n_pairs:] = torch.randint(int(0.75 * text_len), text_len + 1, (n_pairs,), generator=g) - 1 + P
    mask = (torch.arange(S)[None] < lens[:, None]).long()
    labels = ids.clone()
    labels[:, :prompt_len - 1 + P] = -100
    labels[mask == 0] = -100
    ids[mask == 0] = 0
    px = torch.randn(n_pairs, 3, shapes["image_size"], shapes["image_size"], generator=g)
    return dict(input_ids=ids.to(device), attention_mask=mask.to(device), labels=labels.to(device), pixel_values=px), S


def run_library(args, small: bool = False, device: str = "cuda"):
    """-> the JSON line (dict) of the library arm; prints it when run as `bench.py --impl library`."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    dt = torch.bfloat16
    shapes = SMALL if small else LLAVA15_7B
    n_pairs, text_len, prompt_len = (4, 1024, 128) if not small else (2, 96, 24)
    if device == "cuda":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    t0 = time.time()
    policy = build_model(shapes, device, dt)
    ref = build_model(shapes, device, dt)
    ref.load_state_dict(policy.state_dict())
    ref.eval().requires_grad_(False)
    policy.train()
    policy.model.vision_tower.requires_grad_(False)          # --freeze_vision_tower True (dpo.py:55)
    policy.gradient_checkpointing_enable(gradient_checkpointing_kwargs={"use_reentrant": False})
    policy.config.use_cache = False
    params = [p for p in policy.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-6, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.0, fused=(device == "cuda"))
    batch, S = make_batch(shapes, n_pairs, text_len, prompt_len, 1000, device)
    px_host = batch["pixel_values"].pin_memory() if device == "cuda" else batch["pixel_values"]
    build_s = time.time() - t0

    def fwd(model, px):
        # concatenated_inputs duplicates the image tensor [v, v] (base/trainer.py:135-145)
        out = model(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"], pixel_values=torch.cat([px, px], 0),
                    use_cache=False)
        lp = get_batch_logps(out.logits, batch["labels"])
        return lp[:n_pairs], lp[n_pairs:], out.logits

    last = {}

    def step(e2e: bool):
        px = px_host.to(device, non_blocking=True).to(dt) if e2e else px_dev
        pc, pr, logits = fwd(policy, px)
        with torch.no_grad():
            rc, rr, _ = fwd(ref, px)
        losses, cr, rj = dpo_loss(pc.float(), pr.float(), rc.float(), rr.float())
        loss = losses.mean()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        policy.zero_grad(set_to_none=True)
        if e2e:
            last.update(loss=float(loss.detach()), margin=float((cr - rj).mean()))   # the D2H read of the step's result

    px_dev = px_host.to(device).to(dt)

    def timed(fn, steps):
        if device != "cuda":
            t = time.perf_counter()
            for _ in range(steps):
                fn()
            return (time.perf_counter() - t) * 1e3 / steps
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    for _ in range(max(1, min(args.warmup, 3))):
        step(False)
    ms_dev = timed(lambda: step(False), args.steps)
    ms_e2e = timed(lambda: step(True), args.steps)
    h2d = int(px_host.numel() * px_host.element_size())
    line = {"metric": "preference_pairs_per_sec", "value": n_pairs / (ms_dev / 1e3), "unit": "pairs/s", "impl": "library",
            "n_gpus": 1, "steps": args.steps, "warmup": max(1, min(args.warmup, 3)), "ms_per_step": ms_dev, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": ("LLaVA-1.5-7B DPO bf16 full-FT, 4 pairs/GPU, text 1024 (1599 merged), 1x336px image/pair"
                                    if not small else "small (dev config, NOT the benchmark)"),
                       "stack": "transformers LlavaForConditionalGeneration bf16, sdpa attention, gradient checkpointing "
                                "(non-reentrant), frozen CLIP tower, frozen reference copy under no_grad, full [2B,S,V] logits, "
                                "clip_grad_norm_ 1.0, torch.optim.AdamW(fused=True) on bf16 parameters (no fp32 master)",
                       "merged_len": S, "build_s": round(build_s, 1)},
            "e2e": {"value": n_pairs / (ms_e2e / 1e3), "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                    "ms_per_step": ms_e2e, "last_metrics": last}}
    try:
        import transformers
        line["config"]["transformers"] = transformers.__version__
        line["config"]["torch"] = torch.__version__
    except Exception:
        pass
    del policy, ref, opt, params
    if device == "cuda":
        torch.cuda.empty_cache()
    return line


if __name__ == "__main__":   # CPU smoke of the arm on the small shapes: python bench_library.py
    from types import SimpleNamespace
    print(json.dumps(run_library(SimpleNamespace(steps=1, warmup=1), small=True, device="cpu")))
