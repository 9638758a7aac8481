"""LLaVA-1.5 / LLaVA-Next + LoRA variant (what scripts/dpo_llava.sh, dpo_llavanext.sh, kto_*.sh, ddpo_*.sh train) -- CPU tests.

* oracle/lora_restate.py against the fixtures minted from the reference's LlavaForRL / LlavaNextForRL with hand-applied
  peft-style adapters (tests/golden/g11_*.npz, oracle/make_fixtures.py --lora);
* the engine's orchestration (vlrlhf_b200/engine_lora.py over tests/mock_ops.py) against the fixtures and the oracle's
  autograd: log-probs, DDPO, all loss types, adapter gradients, activation checkpointing, optimizer.
"""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from oracle import lora_restate as LR
from oracle import restate as R

G = os.path.join(os.path.dirname(__file__), "golden")
CASES = {"g11_lora_tiny": ("TINY_LORA", LR.TINY_LORA), "g11_lora_small": ("SMALL_LORA", LR.SMALL_LORA),
         "g11_next_lora_tiny": ("TINY_NEXT_LORA", LR.TINY_NEXT_LORA), "g11_next_lora_small": ("SMALL_NEXT_LORA", LR.SMALL_NEXT_LORA)}


def _batch(cfg, d):
    sizes = [tuple(int(x) for x in s) for s in d["image_sizes"]] if "image_sizes" in d.files else None
    return R.make_batch(cfg, int(d["n_pairs"]), int(d["text_len"]), int(d["prompt_len"]), int(d["seed"]), ddpo_like=True,
                        image_sizes=sizes)


@pytest.mark.parametrize("tag", list(CASES))
def test_lora_oracle_matches_reference_fixture(tag):
    cfg = CASES[tag][1]
    d = np.load(os.path.join(G, tag + ".npz"))
    w, lora = LR.make_weights(cfg, int(d["seed"]))
    batch = _batch(cfg, d)
    with torch.no_grad():
        pc, pr, pcl, prl, imap, labels = LR.concatenated_forward(cfg, w, lora, batch)
        rc, rr, _, _, _, _ = LR.concatenated_forward(cfg, w, None, batch)
    assert np.array_equal(imap.numpy(), d["image_position_map"]) and np.array_equal(labels.numpy(), d["labels"])
    np.testing.assert_allclose(torch.cat([pc, pr]).numpy(), d["policy_logps"], rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(torch.cat([rc, rr]).numpy(), d["ref_logps"], rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(float(pcl.mean()), float(d["policy_logits_mean_chosen"]), rtol=1e-3, atol=1e-5)
    assert np.abs(d["policy_logps"] - d["ref_logps"]).max() > 0.05   # the adapters carry signal
    with torch.no_grad():
        for lt in ("sigmoid", "ddpo", "kto_pair", "ipo", "hinge"):
            _, _, aux = LR.get_batch_loss_metrics(cfg, w, lora, batch, loss_type=lt)
            np.testing.assert_allclose(aux["losses"].numpy(), d[f"{lt}_losses"], rtol=1e-3, atol=1e-4)


@pytest.fixture(scope="module")
def lpkg():
    import vlrlhf_b200
    from tests import mock_ops
    names = ("vlrlhf_b200.ops", "vlrlhf_b200.engine", "vlrlhf_b200.engine_lora")
    saved = {k: sys.modules.get(k) for k in names}
    saved_attr = {k: getattr(vlrlhf_b200, k.split(".")[1], None) for k in names}
    sys.modules["vlrlhf_b200.ops"] = mock_ops
    vlrlhf_b200.ops = mock_ops
    for k in names[1:]:
        sys.modules.pop(k, None)
        if hasattr(vlrlhf_b200, k.split(".")[1]):
            delattr(vlrlhf_b200, k.split(".")[1])
    EL = importlib.import_module("vlrlhf_b200.engine_lora")
    from vlrlhf_b200 import config, host
    yield config, EL, host, mock_ops
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


def _setup(lpkg, tag, loss_type="sigmoid", with_optimizer=False, **tc):
    config, EL, host, ops = lpkg
    name, cfg = CASES[tag]
    d = np.load(os.path.join(G, tag + ".npz"))
    eng = EL.LlavaLoRADPOEngine(getattr(config, name), config.TrainConfig(loss_type=loss_type, learning_rate=1e-3, **tc),
                                device="cpu", with_optimizer=with_optimizer)
    eng.init_synthetic(int(d["seed"]))
    return eng, cfg, d, _batch(cfg, d)


def test_lora_weights_mirror_oracle(lpkg):
    eng, cfg, d, batch = _setup(lpkg, "g11_next_lora_tiny")
    w, lora = LR.make_weights(cfg, 0)
    st = eng.hf_state("policy")
    for k, v in lora.items():
        assert torch.equal(st[k].float(), v), k
    for k, v in w.items():
        if k in st:
            assert torch.equal(st[k].float().reshape(v.shape), v), k
    assert set(eng.hf_state("grad")) == set(lora)
    assert not any(k.endswith(("lora_A", "lora_B")) for k in eng.hf_state("ref"))
    # adapters only: the trainable arena is the adapters', the base is a separate frozen arena
    assert eng.params.numel() < eng.bparams.numel() and eng.ref_params.numel() == 0


@pytest.mark.parametrize("tag", list(CASES))
def test_lora_engine_forward_parity_cpu_mock(lpkg, tag):
    config, EL, host, ops = lpkg
    eng, cfg, d, batch = _setup(lpkg, tag)
    cb = host.concatenated_inputs(batch)
    ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    img = cb["concatenated_img_input_dict"]
    px, sizes = img["pixel_values"], img.get("image_sizes")
    out = eng.step(*eng.prepare_inputs(ids, am, lb, px, None, sizes), train=False)
    np.testing.assert_allclose(out.policy_logps.numpy(), d["policy_logps"], rtol=1e-3)
    np.testing.assert_allclose(out.ref_logps.numpy(), d["ref_logps"], rtol=1e-3)
    wt = eng.ddpo_weights(ids, am, lb, sizes)
    assert int(wt.sum()) > 0
    out = eng.step(*eng.prepare_inputs(ids, am, lb, px, wt, sizes), train=False)
    np.testing.assert_allclose(out.policy_logps.numpy(), d["policy_logps_ddpo"], rtol=1e-3, atol=1e-2)
    np.testing.assert_allclose(out.ref_logps.numpy(), d["ref_logps_ddpo"], rtol=1e-3, atol=1e-2)
    slack = 0.1 * 1e-3 * np.abs(d["policy_logps"]).max() * 4
    for lt in ("sigmoid", "kto_pair", "hinge"):
        eng.tc.loss_type = lt
        out = eng.step(*eng.prepare_inputs(ids, am, lb, px, None, sizes), train=False)
        np.testing.assert_allclose(out.losses.numpy(), d[f"{lt}_losses"], atol=slack)


@pytest.mark.parametrize("tag,loss_type", [("g11_lora_tiny", "sigmoid"), ("g11_next_lora_tiny", "ddpo")])
def test_lora_engine_adapter_gradients_match_oracle_autograd(lpkg, tag, loss_type):
    grads = {}
    for ckpt in (False, True):
        eng, cfg, d, batch = _setup(lpkg, tag, loss_type=loss_type, activation_checkpointing=ckpt)
        metrics = eng.train_step(batch, train=True)
        grads[ckpt] = eng.grads.clone()
    assert torch.equal(grads[False], grads[True])
    got = {k: v.float() for k, v in eng.hf_state("grad").items()}
    w, lora = LR.make_weights(cfg, int(d["seed"]))
    leaves = {k: v.clone().requires_grad_(True) for k, v in lora.items()}
    loss, want_metrics, _ = LR.get_batch_loss_metrics(cfg, w, leaves, batch, loss_type=loss_type)
    loss.backward()
    for k, leaf in leaves.items():
        rel = (got[k] - leaf.grad).norm().item() / max(leaf.grad.norm().item(), 1e-12)
        assert rel < 5e-2, f"{k}: rel {rel}"
    assert abs(metrics["loss"] - float(loss.detach())) < 2e-3
    for k in ("rewards/chosen", "rewards/rejected", "logps/chosen", "logps/rejected", "logits/chosen", "logits/rejected"):
        assert abs(metrics[k] - float(want_metrics[k])) <= 2e-3 * max(1.0, abs(float(want_metrics[k]))), k


def test_lora_engine_optimizer_updates_only_adapters(lpkg):
    eng, cfg, d, batch = _setup(lpkg, "g11_lora_tiny", with_optimizer=True, weight_decay=0.1)
    base0, vis0 = eng.bparams.clone(), eng.vparams.clone()
    l0 = eng.train_step(batch, train=True)["loss"]
    for _ in range(4):
        l1 = eng.train_step(batch, train=True)["loss"]
    assert l1 < l0
    assert torch.equal(eng.bparams, base0) and torch.equal(eng.vparams, vis0)
    assert torch.equal(eng.params, eng.master.to(torch.bfloat16))


def test_lora_peft_init_makes_policy_equal_reference(lpkg):
    """peft starts lora_B at zero: the first step's policy and reference log-probs coincide (loss = ln 2)."""
    config, EL, host, ops = lpkg
    eng, cfg, d, batch = _setup(lpkg, "g11_lora_tiny")
    eng.reset_adapters(0)
    m = eng.train_step(batch, train=False)
    assert abs(m["loss"] - float(np.log(2.0))) < 1e-6 and m["rewards/margins"] == 0.0


@pytest.mark.parametrize("tag,ckpt,loss_type", [("g11_lora_tiny", False, "sigmoid"), ("g11_next_lora_tiny", True, "ddpo")])
def test_lora_packed_step_equals_padded_step(lpkg, tag, ckpt, loss_type):
    """TrainConfig.pack_sequences on the LoRA engine: same log-probs, loss, rewards and adapter gradients as the padded step."""
    res = []
    for pack in (False, True):
        eng, cfg, d, batch = _setup(lpkg, tag, loss_type=loss_type, pack_sequences=pack, activation_checkpointing=ckpt)
        metrics = eng.train_step(batch, train=True)
        assert eng._saved["m"].packed == pack
        res.append((metrics, eng.grads.clone()))
    (m0, g0), (m1, g1) = res
    for k in m0:
        if not k.startswith("logits/"):
            assert m0[k] == m1[k], k
    assert torch.equal(g0, g1) and float(g0.float().abs().sum()) > 0
