"""Trainer-side boundary (plugin.make_trainer_class / B200FlatAdamW / B200ModuleMixin) driven through the restated trl 0.8.1 +
HF 4.41 control flow of tests/trl_loop.py -- CPU, over tests/mock_ops.py.  The same scenarios run on the CUDA kernels in
tests/test_gpu_plugin.py.

Checked against: the oracle's get_batch_loss_metrics (values of every trl metric incl. logits/*), the oracle's autograd
(accumulated gradients), and the engine's own fast path (engine.train_step with gradient accumulation + LR schedule)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from oracle import restate as R
from tests import trl_loop

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def cpu_plugin():
    from tests.conftest import mocked_ops
    with mocked_ops("vlrlhf_b200.engine", "vlrlhf_b200.plugin") as m:
        from vlrlhf_b200 import config
        yield m.modules["plugin"], config


def _model(plugin, config, seed, **tc):
    model = plugin.B200LlavaForRL(config.TINY, config.TrainConfig(**tc), device="cpu")
    model.engine.init_synthetic(seed)
    return model


def _batches(n, seed0=0):
    return [R.make_batch(R.TINY, 2, 24, 8, seed0 + i, ddpo_like=True) for i in range(n)]


def test_concatenated_forward_returns_logit_stats_trl_can_use(cpu_plugin):
    """trl 0.8.1 takes `.detach().mean().cpu()` of items 3-4 (they were None -> AttributeError at step 1)."""
    plugin, config = cpu_plugin
    d = np.load(os.path.join(G, "g4_tiny.npz"))
    seed = int(d["seed"])
    model = _model(plugin, config, seed)
    Trainer = plugin.make_trainer_class(trl_loop.StubDPOTrainer)
    tr = Trainer(model, None, args=trl_loop.training_args())
    assert isinstance(tr.ref_model, plugin.RefView) and tr.ref_model.engine is model.engine   # no deep copy of the arenas
    batch = R.make_batch(R.TINY, 2, 24, 8, seed, ddpo_like=True)
    loss, metrics = tr.get_batch_loss_metrics(model, batch)
    wp, wr = R.make_policy_and_ref(R.TINY, seed)
    with torch.no_grad():
        want_loss, want, _ = R.get_batch_loss_metrics(R.TINY, wp, wr, batch)
    assert abs(float(loss) - float(want_loss)) < 2e-3
    for k, v in want.items():
        assert abs(float(metrics[k]) - float(v)) < 2e-3 * max(1.0, abs(float(v))), (k, float(metrics[k]), float(v))
    # evaluation: the policy under no_grad still yields usable logits statistics
    with torch.no_grad():
        _, ev = tr.get_batch_loss_metrics(model, batch, train_eval="eval")
    assert abs(float(ev["eval_logits/chosen"]) - float(want["logits/chosen"])) < 2e-3
    assert abs(float(ev["eval_logps/chosen"]) - float(want["logps/chosen"])) < 2e-3 * abs(float(want["logps/chosen"]))


def test_zero_grad_set_to_none_does_not_lose_the_gradient_views(cpu_plugin):
    """HF calls model.zero_grad() (set_to_none=True by default): the .grad views must survive or be re-attached."""
    plugin, config = cpu_plugin
    model = _model(plugin, config, 0)
    name = "language_model.model.layers.0.self_attn.q_proj.weight"
    p = dict(model.hf_named_parameters())[name]
    view_ptr = model.engine.hf_state("grad")[name].data_ptr()
    model.zero_grad()                      # the override keeps the views
    assert p.grad is not None and p.grad.data_ptr() == view_ptr
    torch.nn.Module.zero_grad(model)       # what a wrapper module (DDP, accelerate) would do: views dropped
    assert p.grad is None
    tr = plugin.make_trainer_class(trl_loop.StubDPOTrainer)(model, None, args=trl_loop.training_args())
    tr.training_step(model, _batches(1)[0])
    assert p.grad is not None and p.grad.data_ptr() == view_ptr and float(p.grad.float().abs().sum()) > 0
    opt = torch.optim.SGD([q for q in model.parameters() if q.requires_grad], lr=1.0)
    before = p.detach().float().clone()
    opt.step()                             # a foreign torch optimizer sees the gradients too
    assert float((p.detach().float() - before).abs().sum()) > 0


def test_gradient_accumulation_matches_oracle_autograd(cpu_plugin):
    """Two micro-batches, loss / 2 each (HF Trainer.training_step): the gradient arena must hold the SUM, not the last one."""
    plugin, config = cpu_plugin
    seed = 3
    model = _model(plugin, config, seed)
    tr = plugin.make_trainer_class(trl_loop.StubDPOTrainer)(
        model, None, args=trl_loop.training_args(gradient_accumulation_steps=2, max_grad_norm=0.0))
    batches = _batches(2, 10)
    model.zero_grad()
    for b in batches:
        tr.training_step(model, b)
    wp, wr = R.make_policy_and_ref(R.TINY, seed)
    names = ["language_model.model.layers.1.mlp.down_proj.weight", "language_model.model.layers.0.self_attn.q_proj.weight",
             "language_model.model.norm.weight", "language_model.model.layers.0.input_layernorm.weight",
             "language_model.lm_head.weight", "multi_modal_projector.linear_1.bias", "language_model.model.embed_tokens.weight"]
    leaves = {n: wp[n].clone().requires_grad_(True) for n in names}
    w = {**wp, **leaves}
    total = 0
    for b in batches:
        loss, _, _ = R.get_batch_loss_metrics(R.TINY, w, wr, b)
        total = total + loss / 2
    total.backward()
    got = model.engine.hf_state("grad")
    for n in names:
        a, e = got[n].float().view(-1), leaves[n].grad.float().view(-1)
        rel = float((a - e).norm() / e.norm().clamp_min(1e-12))
        assert rel < 3e-2, (n, rel)          # bf16 gradient storage, two roundings


@pytest.mark.parametrize("sched", ["cosine", "linear"])
def test_trainer_loop_equals_engine_fast_path(cpu_plugin, sched):
    """plugin trainer loop (trl/HF control flow, B200FlatAdamW stepped by the HF scheduler) == engine.train_step with
    TrainConfig(gradient_accumulation_steps, lr_scheduler_type, warmup): same metrics every micro-step, same weights after."""
    plugin, config = cpu_plugin
    seed, ga, n_opt = 5, 2, 3
    kw = dict(learning_rate=2e-3, adam_beta1=0.9, adam_beta2=0.98, adam_eps=1e-6, weight_decay=0.01, max_grad_norm=1.0)
    batches = _batches(ga * n_opt, 20)
    m1 = _model(plugin, config, seed)
    args = trl_loop.training_args(learning_rate=kw["learning_rate"], adam_beta1=0.9, adam_beta2=0.98, adam_epsilon=1e-6,
                                  weight_decay=0.01, max_grad_norm=1.0, gradient_accumulation_steps=ga,
                                  lr_scheduler_type=sched, warmup_steps=1, max_steps=n_opt)
    tr = plugin.make_trainer_class(trl_loop.StubDPOTrainer)(m1, None, args=args)
    tr.train_loop(batches)
    assert isinstance(tr.optimizer, plugin.B200FlatAdamW) and m1.engine.tc.max_grad_norm == 1.0
    assert m1.engine.opt_step == n_opt
    m2 = _model(plugin, config, seed, gradient_accumulation_steps=ga, lr_scheduler_type=sched, warmup_steps=1,
                max_steps=n_opt, **kw)
    lrs = []
    for i, b in enumerate(batches):
        got = m2.engine.train_step(b)
        for k in ("rewards/chosen", "rewards/margins", "logps/chosen", "logits/chosen", "logits/rejected"):
            assert abs(got[k] - tr.logged[i][k]) < 1e-4 * max(1.0, abs(got[k])), (i, k, got[k], tr.logged[i][k])
        if (i + 1) % ga == 0:
            lrs.append(m2.engine.last_lr)
    m2.engine.wait_optimizer()
    assert lrs[0] == 0.0 and lrs[1] == pytest.approx(kw["learning_rate"])   # warm-up step, then the peak
    assert torch.equal(m1.engine.params, m2.engine.params)
    assert torch.equal(m1.engine.master, m2.engine.master)


def test_flat_optimizer_state_dict_round_trip(cpu_plugin):
    plugin, config = cpu_plugin
    m1 = _model(plugin, config, 1)
    opt = m1.flat_optimizer(lr=1e-3)
    tr = plugin.make_trainer_class(trl_loop.StubDPOTrainer)(m1, None, args=trl_loop.training_args())
    tr.training_step(m1, _batches(1)[0]); opt.step(); m1.zero_grad()
    sd = opt.state_dict()
    m2 = _model(plugin, config, 1)
    opt2 = m2.flat_optimizer(lr=5e-4)
    opt2.load_state_dict(sd)
    assert m2.engine.opt_step == 1 and torch.equal(m2.engine.exp_avg, m1.engine.exp_avg)
    assert opt2.param_groups[0]["lr"] == 1e-3


def test_bad_batches_raise_before_any_device_work(cpu_plugin):
    """a truncated prompt can lose its <image> token; ids beyond the vocabulary would index out of bounds."""
    plugin, config = cpu_plugin
    model = _model(plugin, config, 0)
    tr = plugin.make_trainer_class(trl_loop.StubDPOTrainer)(model, None, args=trl_loop.training_args())
    batch = _batches(1)[0]
    bad = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}
    bad["chosen_input_ids"][0, 1] = 5        # the <image> placeholder is gone from one sequence
    with pytest.raises(ValueError, match="number of image tokens"):
        tr.get_batch_loss_metrics(model, bad)
    bad = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}
    bad["rejected_input_ids"][1, 5] = config.TINY.vocab + 3
    with pytest.raises(IndexError):
        tr.get_batch_loss_metrics(model, bad)


def test_rope_tables_grow_with_the_sequence(cpu_plugin):
    plugin, config = cpu_plugin
    import dataclasses
    cfg = dataclasses.replace(config.TINY, max_positions=16)
    model = plugin.B200LlavaForRL(cfg, config.TrainConfig(), device="cpu")
    model.engine.init_synthetic(0)
    assert model.engine.rope_cos.shape[0] == 16
    d = np.load(os.path.join(G, "g4_tiny.npz"))
    model.engine.init_synthetic(int(d["seed"]))
    batch = R.make_batch(R.TINY, 2, 24, 8, int(d["seed"]), ddpo_like=True)
    tr = plugin.make_trainer_class(trl_loop.StubDPOTrainer)(model, None, args=trl_loop.training_args())
    with torch.no_grad():
        pc, pr, _, _ = tr.concatenated_forward(model, batch)
    assert model.engine.rope_cos.shape[0] >= model.engine._bufs["s.x0"].shape[0] // 4
    np.testing.assert_allclose(torch.cat([pc, pr]).numpy(), d["policy_logps"], rtol=1e-3)
