"""Shared-prefix rows (TrainConfig.share_prefix) on the CPU: the host-side row plan, the CPU mirror of
vlb200_share_prefix_rows, and the engine's orchestration over tests/mock_ops.py -- the step with ONE copy of every pair's common
prompt + image prefix must give the padded step's log-probs / losses and the oracle's gradients.  The CUDA kernels
(vlb200_attn_*_tc_ctx, vlb200_share_prefix_rows) are checked against the same mirrors in tests/test_gpu_share_prefix.py."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from oracle import restate as R

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def cpu_pkg():
    from tests.conftest import mocked_ops
    from tests import mock_ops
    with mocked_ops("vlrlhf_b200.engine", "vlrlhf_b200.engine_lora", "vlrlhf_b200.plugin") as m:
        from vlrlhf_b200 import config, host
        yield config, m.modules["engine"], host, mock_ops


def test_shared_prefix_rows_plan(cpu_pkg):
    config, engine, host, ops = cpu_pkg
    IMG, P = 90, 5
    #            0  1    2  3  4  5  6  7
    c = [[1, IMG, 7, 8, 9, 10, 11, 12], [1, IMG, 7, 8, 9, 3, 0, 0], [1, 7, IMG, 8, 9, 10, 0, 0], [1, IMG, 5, 6, 0, 0, 0, 0]]
    r = [[1, IMG, 7, 8, 4, 10, 11, 0], [1, IMG, 7, 8, 9, 3, 0, 0], [1, 7, 6, IMG, 9, 10, 11, 0], [1, IMG, 5, 6, 7, 8, 9, 0]]
    ids = torch.tensor(c + r)
    am = (ids != 0).long()
    rows = host.shared_prefix_rows(ids, am, IMG, P)
    # pair 0: common tokens 0..3 (4 tokens incl. the image) -> 4 + (P-1) rows
    # pair 1: identical sequences of 6 tokens: capped so that each keeps one row of its own -> 5 tokens -> 5 + 4 rows
    # pair 2: the sequences part BEFORE the image placeholder -> the image rows stay per sequence -> nothing shared
    # pair 3: chosen (4 tokens) is a prefix of rejected: capped at 3 tokens -> 3 + 4 rows
    assert rows == [4 + P - 1, 5 + P - 1, 0, 3 + P - 1]
    assert host.shared_prefix_rows(ids, am, IMG, P, min_suffix=2)[1] == 4 + P - 1
    with pytest.raises(ValueError):
        host.shared_prefix_rows(ids[:3], am[:3], IMG, P)


def _step(cpu_pkg, cfg_name, rcfg, batch, seed, mode, loss_type="sigmoid", eng_cls=None):
    config, engine, host, ops = cpu_pkg
    tc = config.TrainConfig(pack_sequences=(mode == "packed"), share_prefix=(mode == "shared"), loss_type=loss_type)
    eng = (eng_cls or engine.LlavaDPOEngine)(getattr(config, cfg_name), tc, device="cpu", with_optimizer=False)
    eng.init_synthetic(seed)
    out = eng.train_step(batch, train=True)
    return eng, out


@pytest.mark.parametrize("tag,cfg_name,rcfg,shape", [("g4_tiny", "TINY", R.TINY, (2, 24, 8)), ("g4_small", "SMALL", R.SMALL, (2, 96, 24))])
@pytest.mark.parametrize("loss_type", ["sigmoid", "ddpo"])
def test_shared_step_equals_padded_step_and_oracle(cpu_pkg, tag, cfg_name, rcfg, shape, loss_type):
    config, engine, host, ops = cpu_pkg
    d = np.load(os.path.join(G, tag + ".npz"))
    seed = int(d["seed"])
    batch = R.make_batch(rcfg, *shape, seed, ddpo_like=True)
    e0, o0 = _step(cpu_pkg, cfg_name, rcfg, batch, seed, "padded", loss_type)
    e1, o1 = _step(cpu_pkg, cfg_name, rcfg, batch, seed, "shared", loss_type)
    m = e1._saved["m"]
    assert m.shared and m.shared_rows > 0 and m.T == sum(int(x) for x in m.att_lens) < e0._saved["m"].T - m.shared_rows + 1
    # prompt (shape[2] tokens incl. the image placeholder) is common by construction; ddpo_like responses share more
    assert m.shared_rows >= 2 * (shape[2] + rcfg.n_patches - 1)
    # (logits/* are means over the attended positions once the padding rows are gone: the one documented difference)
    for k in ("loss", "rewards/chosen", "rewards/rejected", "rewards/margins", "logps/chosen", "logps/rejected"):
        assert abs(o0[k] - o1[k]) <= 2e-5 * max(1.0, abs(o0[k])), (k, o0[k], o1[k])
    if loss_type == "sigmoid":
        key = "policy_logps"
        got = np.array([o1["logps/chosen"], o1["logps/rejected"]])
        n = shape[0]
        np.testing.assert_allclose(got, [d[key][:n].mean(), d[key][n:].mean()], rtol=1e-3)   # the reference's own numbers
    # gradients against the oracle's autograd: the shared layout must be as close as the padded one
    wp, wr = R.make_policy_and_ref(rcfg, seed)
    names = ["language_model.model.layers.0.self_attn.q_proj.weight", "language_model.model.layers.0.self_attn.k_proj.weight",
             "language_model.model.layers.0.self_attn.v_proj.weight", "language_model.model.layers.1.mlp.down_proj.weight",
             "language_model.model.layers.0.input_layernorm.weight", "language_model.model.embed_tokens.weight",
             "language_model.lm_head.weight", "multi_modal_projector.linear_1.weight", "multi_modal_projector.linear_2.bias"]
    leaves = {n: wp[n].clone().requires_grad_(True) for n in names}
    loss, _, _ = R.get_batch_loss_metrics(rcfg, {**wp, **leaves}, wr, batch, loss_type=loss_type)
    loss.backward()
    g0, g1 = e0.hf_state("grad"), e1.hf_state("grad")
    for n in names:
        w = leaves[n].grad.float().view(-1)
        r0 = float((g0[n].float().view(-1) - w).norm() / w.norm().clamp_min(1e-12))
        r1 = float((g1[n].float().view(-1) - w).norm() / w.norm().clamp_min(1e-12))
        assert r1 < max(2.0 * r0, 2e-2), (n, r0, r1)


def test_shared_step_with_a_pair_that_shares_nothing(cpu_pkg):
    """One pair's sequences differ at token 0 (no common prefix): its image rows stay per sequence, the other pair shares."""
    config, engine, host, ops = cpu_pkg
    seed = 4
    batch = R.make_batch(R.TINY, 2, 24, 8, seed, ddpo_like=True)
    batch["rejected_input_ids"] = batch["rejected_input_ids"].clone()
    batch["rejected_input_ids"][1, 0] = 7
    e0, o0 = _step(cpu_pkg, "TINY", R.TINY, batch, seed, "padded")
    e1, o1 = _step(cpu_pkg, "TINY", R.TINY, batch, seed, "shared")
    m = e1._saved["m"]
    assert int(m.att_lens[2 + 1]) == 0 and int(m.att_ctx[1]) == -1 and int(m.att_lens[2 + 0]) > 0
    for k in ("loss", "logps/chosen", "logps/rejected"):
        assert abs(o0[k] - o1[k]) <= 2e-5 * max(1.0, abs(o0[k])), (k, o0[k], o1[k])
    g0, g1 = e0.grads.float(), e1.grads.float()
    assert float((g0 - g1).norm() / g0.norm()) < 1e-2


def test_shared_step_lora_engine_and_plugin(cpu_pkg):
    config, engine, host, ops = cpu_pkg
    from oracle import lora_restate as LR
    engine_lora = importlib.import_module("vlrlhf_b200.engine_lora")
    seed = 0
    batch = R.make_batch(LR.TINY_LORA, 2, 24, 8, seed, ddpo_like=True)
    e0, o0 = _step(cpu_pkg, "TINY_LORA", LR.TINY_LORA, batch, seed, "padded", eng_cls=engine_lora.LlavaLoRADPOEngine)
    e1, o1 = _step(cpu_pkg, "TINY_LORA", LR.TINY_LORA, batch, seed, "shared", eng_cls=engine_lora.LlavaLoRADPOEngine)
    for k in ("loss", "logps/chosen", "logps/rejected"):
        assert abs(o0[k] - o1[k]) <= 2e-5 * max(1.0, abs(o0[k])), (k, o0[k], o1[k])
    g0, g1 = e0.grads.float(), e1.grads.float()
    assert float((g0 - g1).norm() / g0.norm()) < 2e-2
    # through the Trainer-side boundary: concatenated_forward builds the row plan from the host batch
    plugin = importlib.import_module("vlrlhf_b200.plugin")
    from tests import trl_loop
    model = plugin.B200LlavaForRL(config.TINY, config.TrainConfig(share_prefix=True), device="cpu")
    model.engine.init_synthetic(seed)
    tr = plugin.make_trainer_class(trl_loop.StubDPOTrainer)(model, None, args=trl_loop.training_args())
    b2 = R.make_batch(R.TINY, 2, 24, 8, seed, ddpo_like=True)
    tr.training_step(model, b2)
    assert model.engine._saved["m"].shared
    wp, wr = R.make_policy_and_ref(R.TINY, seed)
    with torch.no_grad():
        _, want, _ = R.get_batch_loss_metrics(R.TINY, wp, wr, b2)
    assert abs(tr.logged[0]["logps/chosen"] - float(want["logps/chosen"])) < 1e-3 * abs(float(want["logps/chosen"]))
    assert abs(tr.logged[0]["rewards/margins"] - float(want["rewards/margins"])) < 2e-3
