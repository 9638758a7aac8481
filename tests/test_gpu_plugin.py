"""The Trainer-side boundary ON THE CUDA KERNELS: plugin.make_trainer_class / B200FlatAdamW / B200ModuleMixin driven through
the restated trl 0.8.1 + HF 4.41 control flow (tests/trl_loop.py), i.e. what `src/vlrlhf/dpo.py` does with the model and
trainer classes `plugin.install*()` registers:

    concatenated_forward(policy) -> concatenated_forward(RefView) under no_grad -> dpo_loss -> (losses.mean()/k).backward()
    -> [k micro-batches] -> optimizer.step() (B200FlatAdamW) -> scheduler.step() -> model.zero_grad()

Checked against the oracle (metric values incl. logits/*; accumulated gradients vs autograd) and the engine's own fast path
(engine.train_step: same kernels, reduction and optimizer on a side stream) -- the weights after three optimizer steps must
be IDENTICAL, which also covers the `_EngineLogps.backward` vs deferred-optimizer ordering the CPU mock cannot see.
"""
import os

import numpy as np
import pytest
import torch

from oracle import lora_restate as LR
from oracle import restate as R
from tests import trl_loop

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def pkg():
    import vlrlhf_b200  # noqa: F401
    from vlrlhf_b200 import config, plugin
    return config, plugin


def _model(pkg, cfg_name, seed, **tc):
    config, plugin = pkg
    model = plugin.B200LlavaForRL(getattr(config, cfg_name), config.TrainConfig(**tc))
    model.engine.init_synthetic(seed)
    return model


def _batches(rcfg, n, seed0, npairs=2, tl=96, pl=24):
    return [R.make_batch(rcfg, npairs, tl, pl, seed0 + i, ddpo_like=True) for i in range(n)]


def test_trl_metrics_through_the_plugin_match_oracle(pkg):
    config, plugin = pkg
    d = np.load(os.path.join(G, "g4_small.npz"))
    seed = int(d["seed"])
    model = _model(pkg, "SMALL", seed)
    tr = plugin.make_trainer_class(trl_loop.StubDPOTrainer)(model, None, args=trl_loop.training_args())
    assert isinstance(tr.ref_model, plugin.RefView) and tr.ref_model.engine is model.engine
    batch = R.make_batch(R.SMALL, 2, 96, 24, seed, ddpo_like=True)
    loss, metrics = tr.get_batch_loss_metrics(model, batch)
    pol = torch.cat([tr._last["pc"], tr._last["pr"]]).detach().cpu().numpy()
    ref = torch.cat([tr._last["rc"], tr._last["rr"]]).cpu().numpy()
    np.testing.assert_allclose(pol, d["policy_logps"], rtol=1e-3)      # golden = the reference's LlavaForRL + get_batch_logps
    np.testing.assert_allclose(ref, d["ref_logps"], rtol=1e-3)
    wp, wr = R.make_policy_and_ref(R.SMALL, seed)
    with torch.no_grad():
        want_loss, want, _ = R.get_batch_loss_metrics(R.SMALL, wp, wr, batch)
    tol = 1e-3 * float(np.abs(d["policy_logps"]).max()) * 0.1 * 2    # beta x two log-probs at the 1e-3 bound
    assert abs(float(loss.detach()) - float(want_loss)) < tol
    for k, v in want.items():
        lim = 1e-3 * abs(float(v)) if k.startswith("logps/") else (2e-3 if k.startswith("logits/") else tol)
        assert abs(float(metrics[k]) - float(v)) <= max(lim, 1e-6), (k, float(metrics[k]), float(v))
    with torch.no_grad():   # evaluation pass: policy under no_grad still carries logits statistics
        _, ev = tr.get_batch_loss_metrics(model, batch, train_eval="eval")
    # (the no-grad pass fuses GELU / SwiGLU into the GEMM epilogues, one rounding fewer than the saved-for-backward pass)
    assert abs(float(ev["eval_logits/chosen"]) - float(want["logits/chosen"])) < 2e-3
    assert abs(float(ev["eval_logps/chosen"]) - float(want["logps/chosen"])) < 1e-3 * abs(float(want["logps/chosen"]))


def test_plugin_backward_accumulates_like_oracle_autograd(pkg):
    config, plugin = pkg
    seed = 3
    model = _model(pkg, "SMALL", seed)
    tr = plugin.make_trainer_class(trl_loop.StubDPOTrainer)(
        model, None, args=trl_loop.training_args(gradient_accumulation_steps=2, max_grad_norm=0.0))
    batches = _batches(R.SMALL, 2, 10)
    model.zero_grad()
    torch.nn.Module.zero_grad(model)    # a wrapper's zero_grad(set_to_none=True): the views are dropped ...
    for b in batches:
        tr.training_step(model, b)      # ... and re-attached by the first backward; the second one accumulates
    torch.cuda.synchronize()
    wp, wr = R.make_policy_and_ref(R.SMALL, seed)
    names = ["language_model.model.layers.1.mlp.down_proj.weight", "language_model.model.layers.0.self_attn.q_proj.weight",
             "language_model.model.layers.0.self_attn.v_proj.weight", "language_model.model.layers.1.mlp.gate_proj.weight",
             "language_model.model.norm.weight", "language_model.model.layers.0.input_layernorm.weight",
             "language_model.lm_head.weight", "multi_modal_projector.linear_1.weight", "multi_modal_projector.linear_2.bias",
             "language_model.model.embed_tokens.weight"]
    leaves = {n: wp[n].clone().requires_grad_(True) for n in names}
    w = {**wp, **leaves}
    total = 0
    for b in batches:
        loss, _, _ = R.get_batch_loss_metrics(R.SMALL, w, wr, b)
        total = total + loss / 2
    total.backward()
    params = dict(model.hf_named_parameters())
    for n in names:
        assert params[n].grad is not None
        a, e = params[n].grad.float().cpu().view(-1), leaves[n].grad.float().view(-1)
        rel = float((a - e).norm() / e.norm().clamp_min(1e-12))
        cos = float(torch.dot(a, e) / (a.norm() * e.norm()).clamp_min(1e-20))
        print(f"[accumulate x2] {n}: rel-l2 {rel:.3e} cos {cos:.6f}")
        assert rel < 4e-2 and cos > 0.998, (n, rel, cos)


@pytest.mark.parametrize("ckpt", [False, True])
def test_trainer_loop_equals_engine_fast_path(pkg, ckpt):
    config, plugin = pkg
    seed, ga, n_opt = 5, 2, 3
    kw = dict(learning_rate=2e-3, adam_beta1=0.9, adam_beta2=0.98, adam_eps=1e-6, weight_decay=0.01, max_grad_norm=1.0)
    batches = _batches(R.SMALL, ga * n_opt, 20)
    m1 = _model(pkg, "SMALL", seed)
    args = trl_loop.training_args(learning_rate=kw["learning_rate"], adam_beta1=0.9, adam_beta2=0.98, adam_epsilon=1e-6,
                                  weight_decay=0.01, max_grad_norm=1.0, gradient_accumulation_steps=ga,
                                  lr_scheduler_type="cosine", warmup_steps=1, max_steps=n_opt, gradient_checkpointing=ckpt)
    tr = plugin.make_trainer_class(trl_loop.StubDPOTrainer)(m1, None, args=args)
    assert m1.is_gradient_checkpointing == ckpt
    tr.train_loop(batches)
    assert isinstance(tr.optimizer, plugin.B200FlatAdamW) and m1.engine.opt_step == n_opt
    m2 = _model(pkg, "SMALL", seed, gradient_accumulation_steps=ga, lr_scheduler_type="cosine", warmup_steps=1,
                max_steps=n_opt, activation_checkpointing=ckpt, **kw)
    for i, b in enumerate(batches):
        got = m2.engine.train_step(b)
        for k in ("rewards/chosen", "rewards/margins", "logps/chosen", "logits/chosen", "logits/rejected"):
            assert abs(got[k] - tr.logged[i][k]) <= 1e-5 * max(1.0, abs(got[k])), (i, k, got[k], tr.logged[i][k])
    m2.engine.wait_optimizer()
    torch.cuda.synchronize()
    assert torch.equal(m1.engine.params, m2.engine.params)
    assert torch.equal(m1.engine.master, m2.engine.master)
    # the weights did move, and a torch-visible parameter shows the engine's update
    p = dict(m1.hf_named_parameters())["language_model.model.layers.0.mlp.up_proj.weight"]
    fresh = _model(pkg, "SMALL", seed)
    assert not torch.equal(p.detach(), dict(fresh.hf_named_parameters())["language_model.model.layers.0.mlp.up_proj.weight"].detach())


def test_lora_wrapper_through_the_trainer_loop(pkg):
    """The LoRA family wrapper (what every scripts/*.sh trains): LoraConfig check, RefView = adapters off, accumulation and
    the flat optimizer over the adapter-only arena; == engine_lora fast path."""
    config, plugin = pkg
    from types import SimpleNamespace
    from vlrlhf_b200 import plugin_lora
    d = np.load(os.path.join(G, "g11_lora_small.npz"))
    seed = int(d["seed"])
    cfg = config.SMALL_LORA
    batches = [R.make_batch(LR.SMALL_LORA, int(d["n_pairs"]), int(d["text_len"]), int(d["prompt_len"]), seed + i, ddpo_like=True)
               for i in range(4)]
    peft_cfg = SimpleNamespace(r=cfg.lora_r, lora_alpha=cfg.lora_alpha,
                               target_modules=["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"])
    Trainer = plugin_lora._trainer_class(trl_loop.StubDPOTrainer, plugin)
    kw = dict(learning_rate=1e-3, adam_beta1=0.9, adam_beta2=0.98, adam_eps=1e-6, weight_decay=0.05, max_grad_norm=1.0)
    m1 = plugin_lora.B200LlavaLoRAForRL(cfg, config.TrainConfig())
    m1.engine.init_synthetic(seed)
    with pytest.raises(ValueError):
        Trainer(m1, None, args=trl_loop.training_args(), peft_config=None)
    args = trl_loop.training_args(learning_rate=1e-3, adam_beta1=0.9, adam_beta2=0.98, adam_epsilon=1e-6, weight_decay=0.05,
                                  max_grad_norm=1.0, gradient_accumulation_steps=2, lr_scheduler_type="linear", max_steps=2)
    tr = Trainer(m1, None, args=args, peft_config=peft_cfg)
    tr.train_loop(batches)
    np.testing.assert_allclose([tr.logged[0]["logps/chosen"]], [float(np.mean(d["policy_logps"][:int(d["n_pairs"])]))], rtol=1e-3)
    m2 = plugin_lora.B200LlavaLoRAForRL(cfg, config.TrainConfig(gradient_accumulation_steps=2, lr_scheduler_type="linear",
                                                                max_steps=2, **kw))
    m2.engine.init_synthetic(seed)
    for i, b in enumerate(batches):
        got = m2.engine.train_step(b)
        assert abs(got["rewards/margins"] - tr.logged[i]["rewards/margins"]) <= 1e-5 * max(1.0, abs(got["rewards/margins"]))
    m2.engine.wait_optimizer()
    torch.cuda.synchronize()
    assert torch.equal(m1.engine.params, m2.engine.params)
    assert float(tr.logged[-1]["rewards/margins"]) != float(tr.logged[0]["rewards/margins"])
