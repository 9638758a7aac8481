"""InternLM-XComposer2-VL + PLoRA/LoRA variant (BASELINE.json configs[4]) -- CPU tests.

* oracle/xc2_restate.py against the fixtures minted from the reference's InternLMXC2ForRL (tests/golden/g10_xc2_*.npz);
* the engine's orchestration (vlrlhf_b200/engine_xc2.py over tests/mock_ops.py) against the fixtures and the oracle's
  autograd: log-probs, DDPO, KTO-pair, adapter gradients, activation checkpointing, optimizer.
"""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from oracle import restate as R
from oracle import xc2_restate as X

G = os.path.join(os.path.dirname(__file__), "golden")
CASES = {"g10_xc2_tiny": ("TINY_XC2", X.TINY_XC2), "g10_xc2_small": ("SMALL_XC2", X.SMALL_XC2)}


@pytest.mark.parametrize("tag", list(CASES))
def test_xc2_oracle_matches_reference_fixture(tag):
    xcfg = CASES[tag][1]
    d = np.load(os.path.join(G, tag + ".npz"))
    w, lora = X.make_weights(xcfg, int(d["seed"]))
    batch = R.make_batch(xcfg, int(d["n_pairs"]), int(d["text_len"]), int(d["prompt_len"]), int(d["seed"]), ddpo_like=True)
    with torch.no_grad():
        pc, pr, pcl, prl, imap, labels = X.concatenated_forward(xcfg, w, lora, batch)
        rc, rr, _, _, _, _ = X.concatenated_forward(xcfg, w, None, batch)
    assert np.array_equal(imap.numpy(), d["image_position_map"]) and np.array_equal(labels.numpy(), d["labels"])
    np.testing.assert_allclose(torch.cat([pc, pr]).numpy(), d["policy_logps"], rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(torch.cat([rc, rr]).numpy(), d["ref_logps"], rtol=1e-5, atol=1e-3)
    if "policy_logits" in d.files:
        np.testing.assert_allclose(torch.cat([pcl, prl]).numpy(), d["policy_logits"], rtol=2e-4, atol=2e-4)
    assert np.abs(d["policy_logps"] - d["ref_logps"]).max() > 0.05
    with torch.no_grad():
        for lt in ("sigmoid", "ddpo", "kto_pair", "ipo", "hinge"):
            _, _, aux = X.get_batch_loss_metrics(xcfg, w, lora, batch, loss_type=lt)
            np.testing.assert_allclose(aux["losses"].numpy(), d[f"{lt}_losses"], rtol=1e-3, atol=1e-4)


@pytest.fixture(scope="module")
def xpkg():
    import vlrlhf_b200
    from tests import mock_ops
    names = ("vlrlhf_b200.ops", "vlrlhf_b200.engine", "vlrlhf_b200.engine_qwen", "vlrlhf_b200.engine_xc2")
    saved = {k: sys.modules.get(k) for k in names}
    saved_attr = {k: getattr(vlrlhf_b200, k.split(".")[1], None) for k in names}
    sys.modules["vlrlhf_b200.ops"] = mock_ops
    vlrlhf_b200.ops = mock_ops
    for k in names[1:]:
        sys.modules.pop(k, None)
        if hasattr(vlrlhf_b200, k.split(".")[1]):
            delattr(vlrlhf_b200, k.split(".")[1])
    EX = importlib.import_module("vlrlhf_b200.engine_xc2")
    from vlrlhf_b200 import config, host
    yield config, EX, host, mock_ops
    for k, v in saved.items():
        attr = k.split(".")[1]
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
        if saved_attr[k] is None:
            if hasattr(vlrlhf_b200, attr):
                delattr(vlrlhf_b200, attr)
        else:
            setattr(vlrlhf_b200, attr, saved_attr[k])


def _setup(xpkg, tag, loss_type="sigmoid", with_optimizer=False, **tc):
    config, EX, host, ops = xpkg
    name, xcfg = CASES[tag]
    d = np.load(os.path.join(G, tag + ".npz"))
    eng = EX.XC2DPOEngine(getattr(config, name), config.TrainConfig(loss_type=loss_type, learning_rate=1e-3, **tc), device="cpu",
                          with_optimizer=with_optimizer)
    eng.init_synthetic(int(d["seed"]))
    batch = R.make_batch(xcfg, int(d["n_pairs"]), int(d["text_len"]), int(d["prompt_len"]), int(d["seed"]), ddpo_like=True)
    return eng, xcfg, d, batch


def test_xc2_weights_mirror_oracle_incl_qkv_relayout(xpkg):
    config, EX, host, ops = xpkg
    eng, xcfg, d, batch = _setup(xpkg, "g10_xc2_tiny")
    w, lora = X.make_weights(xcfg, 0)
    st = eng.hf_state("policy")
    for k, v in lora.items():
        assert torch.equal(st[k].float(), v), k
    for k, v in w.items():
        if not k.startswith("vit."):
            assert torch.equal(st[k].float().reshape(v.shape), v), k
    vis = eng._vision_names()
    for k, v in w.items():
        if k.startswith("vit."):
            assert torch.equal(vis[k].float().reshape(v.shape), v), k
    # the engine stores wqkv rows as [all q | all k | all v]
    H, KV, dh = xcfg.heads, xcfg.kv_heads, xcfg.head_dim
    ref = w["model.layers.0.attention.wqkv.weight"].view(KV, H // KV + 2, dh, -1)
    got = eng.base["L0.wqkv"].float()
    assert torch.equal(got[:H * dh], ref[:, :H // KV].reshape(H * dh, -1))
    assert torch.equal(got[H * dh:(H + KV) * dh], ref[:, -2].reshape(KV * dh, -1))
    assert torch.equal(got[(H + KV) * dh:], ref[:, -1].reshape(KV * dh, -1))


@pytest.mark.parametrize("tag", list(CASES))
def test_xc2_engine_forward_parity_cpu_mock(xpkg, tag):
    config, EX, host, ops = xpkg
    eng, xcfg, d, batch = _setup(xpkg, tag)
    cb = host.concatenated_inputs(batch)
    ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    px = cb["concatenated_img_input_dict"]["pixel_values"]
    out = eng.step(*eng.prepare_inputs(ids, am, lb, px), train=False)
    np.testing.assert_allclose(out.policy_logps.numpy(), d["policy_logps"], rtol=1e-3)
    np.testing.assert_allclose(out.ref_logps.numpy(), d["ref_logps"], rtol=1e-3)
    wt = eng.ddpo_weights(ids, am, lb)
    assert int(wt.sum()) > 0
    out = eng.step(*eng.prepare_inputs(ids, am, lb, px, wt), train=False)
    np.testing.assert_allclose(out.policy_logps.numpy(), d["policy_logps_ddpo"], rtol=1e-3, atol=1e-2)
    # KTO-pair (configs[4]) on the same log-probs
    eng.tc.loss_type = "kto_pair"
    out = eng.step(*eng.prepare_inputs(ids, am, lb, px), train=False)
    # losses are sigmoids of beta * (differences of ~1e2..1e3-sized log-probs, each good to 1e-3 relative)
    np.testing.assert_allclose(out.losses.numpy(), d["kto_pair_losses"], atol=0.1 * 1e-3 * np.abs(d["policy_logps"]).max() * 4)


def test_xc2_engine_adapter_gradients_match_oracle_autograd(xpkg):
    config, EX, host, ops = xpkg
    grads = {}
    for ckpt in (False, True):
        eng, xcfg, d, batch = _setup(xpkg, "g10_xc2_tiny", loss_type="kto_pair", activation_checkpointing=ckpt)
        metrics = eng.train_step(batch, train=True)
        grads[ckpt] = eng.grads.clone()
    assert torch.equal(grads[False], grads[True])
    got = {k: v.float() for k, v in eng.hf_state("grad").items()}
    w, lora = X.make_weights(xcfg, int(d["seed"]))
    leaves = {k: v.clone().requires_grad_(True) for k, v in lora.items()}
    loss, want_metrics, _ = X.get_batch_loss_metrics(xcfg, w, leaves, batch, loss_type="kto_pair")
    loss.backward()
    for k, leaf in leaves.items():
        rel = (got[k] - leaf.grad).norm().item() / max(leaf.grad.norm().item(), 1e-12)
        assert rel < 5e-2, f"{k}: rel {rel}"
    assert abs(metrics["loss"] - float(loss.detach())) < 2e-3
    for k in ("rewards/chosen", "rewards/rejected", "logps/chosen", "logps/rejected", "logits/chosen", "logits/rejected"):
        assert abs(metrics[k] - float(want_metrics[k])) <= 2e-3 * max(1.0, abs(float(want_metrics[k]))), k


def test_xc2_engine_optimizer_updates_only_adapters(xpkg):
    config, EX, host, ops = xpkg
    eng, xcfg, d, batch = _setup(xpkg, "g10_xc2_tiny", with_optimizer=True, weight_decay=0.1)
    base0, vis0 = eng.bparams.clone(), eng.vparams.clone()
    l0 = eng.train_step(batch, train=True)["loss"]
    for _ in range(4):
        l1 = eng.train_step(batch, train=True)["loss"]
    assert l1 < l0
    assert torch.equal(eng.bparams, base0) and torch.equal(eng.vparams, vis0)
    assert torch.equal(eng.params, eng.master.to(torch.bfloat16))


@pytest.mark.parametrize("loss_type,ckpt", [("kto_pair", False), ("ddpo", True)])
def test_xc2_packed_step_equals_padded_step(xpkg, loss_type, ckpt):
    """TrainConfig.pack_sequences on the XC2 engine (the partial-LoRA image rows become absolute packed rows, the arange
    rotary positions are packed along): same log-probs, loss, rewards and adapter gradients as the padded step."""
    res = []
    for pack in (False, True):
        eng, xcfg, d, batch = _setup(xpkg, "g10_xc2_tiny", loss_type=loss_type, pack_sequences=pack, activation_checkpointing=ckpt)
        assert int((batch["chosen_attention_mask"] == 0).sum() + (batch["rejected_attention_mask"] == 0).sum()) > 0
        metrics = eng.train_step(batch, train=True)
        m = eng._saved["m"]
        assert m.packed == pack and (not pack or m.T < m.n_seq * m.S)
        res.append((metrics, eng.grads.clone()))
    (m0, g0), (m1, g1) = res
    for k in m0:
        if not k.startswith("logits/"):
            assert m0[k] == m1[k], k
    assert torch.equal(g0, g1) and float(g0.float().abs().sum()) > 0


@pytest.mark.parametrize("loss_type,ckpt", [("kto_pair", False), ("ddpo", True)])
def test_xc2_shared_prefix_step_equals_padded_step(xpkg, loss_type, ckpt):
    """TrainConfig.share_prefix on the XC2 engine: one copy of every pair's prompt + 1225-row image prefix (the partial-LoRA
    image rows are listed once, the rejected copies become -1 and are skipped by the gathers / scatters): log-probs, loss and
    rewards of the padded step, adapter gradients equal up to the accumulation order."""
    res = []
    for share in (False, True):
        eng, xcfg, d, batch = _setup(xpkg, "g10_xc2_tiny", loss_type=loss_type, share_prefix=share, activation_checkpointing=ckpt)
        metrics = eng.train_step(batch, train=True)
        m = eng._saved["m"]
        assert m.shared == share
        if share:
            assert m.shared_rows >= int(d["n_pairs"]) * (int(d["prompt_len"]) + xcfg.n_patches - 1) and m.T < m.n_seq * m.S - m.shared_rows + 1
            assert int((eng._img_rows >= 0).sum()) == int(d["n_pairs"]) * xcfg.n_patches   # every image row once
        res.append((metrics, eng.grads.clone()))
    (m0, g0), (m1, g1) = res
    for k in ("loss", "rewards/chosen", "rewards/rejected", "rewards/margins", "logps/chosen", "logps/rejected"):
        assert abs(m0[k] - m1[k]) <= 5e-5 * max(1.0, abs(m0[k])), (k, m0[k], m1[k])
    g0, g1 = g0.float(), g1.float()
    assert float(g0.abs().sum()) > 0 and float((g0 - g1).norm() / g0.norm()) < 2e-2
