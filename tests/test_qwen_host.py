"""Qwen-VL + LoRA variant (BASELINE.json configs[2]) -- CPU tests.

* oracle/qwen_restate.py against the fixtures minted from the reference's vendored QWenLMHeadModel + VisionTransformer
  (tests/golden/g9_qwen_*.npz: policy = base + adapters, reference = adapters off);
* the engine's orchestration (vlrlhf_b200/engine_qwen.py over tests/mock_ops.py) against the fixtures and the oracle's
  autograd: log-probs, DDPO, adapter gradients, activation checkpointing, optimizer, metrics.
"""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from oracle import qwen_restate as Q
from oracle import restate as R

G = os.path.join(os.path.dirname(__file__), "golden")
CASES = {"g9_qwen_tiny": ("TINY_QWEN", Q.TINY_QWEN), "g9_qwen_small": ("SMALL_QWEN", Q.SMALL_QWEN)}


@pytest.mark.parametrize("tag", list(CASES))
def test_qwen_oracle_matches_reference_fixture(tag):
    qcfg = CASES[tag][1]
    d = np.load(os.path.join(G, tag + ".npz"))
    w, lora = Q.make_weights(qcfg, int(d["seed"]))
    batch = Q.make_batch(qcfg, int(d["n_pairs"]), int(d["text_len"]), int(d["prompt_len"]), int(d["seed"]), ddpo_like=True)
    with torch.no_grad():
        pc, pr, pcl, prl, imap = Q.concatenated_forward(qcfg, w, lora, batch)
        rc, rr, _, _, _ = Q.concatenated_forward(qcfg, w, None, batch)
    assert np.array_equal(imap.numpy(), d["image_position_map"])
    np.testing.assert_allclose(torch.cat([pc, pr]).numpy(), d["policy_logps"], rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(torch.cat([rc, rr]).numpy(), d["ref_logps"], rtol=1e-5, atol=1e-3)
    if "policy_logits" in d.files:
        np.testing.assert_allclose(torch.cat([pcl, prl]).numpy(), d["policy_logits"], rtol=2e-4, atol=2e-4)
    assert np.abs(d["policy_logps"] - d["ref_logps"]).max() > 0.05  # the adapters carry signal
    with torch.no_grad():
        for lt in ("sigmoid", "ddpo", "kto_pair", "ipo", "hinge"):
            _, _, aux = Q.get_batch_loss_metrics(qcfg, w, lora, batch, loss_type=lt)
            np.testing.assert_allclose(aux["losses"].numpy(), d[f"{lt}_losses"], rtol=1e-3, atol=1e-4)
            if lt == "ddpo":
                pol = torch.cat([aux["policy_chosen_logps"], aux["policy_rejected_logps"]]).numpy()
                np.testing.assert_allclose(pol, d["policy_logps_ddpo"], rtol=1e-5, atol=1e-3)


@pytest.fixture(scope="module")
def qpkg():
    import vlrlhf_b200  # noqa: F401
    from tests import mock_ops
    names = ("vlrlhf_b200.ops", "vlrlhf_b200.engine", "vlrlhf_b200.engine_qwen")
    saved = {k: sys.modules.get(k) for k in names}
    saved_attr = {k: getattr(vlrlhf_b200, k.split(".")[1], None) for k in names}
    sys.modules["vlrlhf_b200.ops"] = mock_ops
    vlrlhf_b200.ops = mock_ops  # `from . import ops` resolves through the package attribute when it exists
    for k in names[1:]:
        sys.modules.pop(k, None)
        if hasattr(vlrlhf_b200, k.split(".")[1]):
            delattr(vlrlhf_b200, k.split(".")[1])
    importlib.import_module("vlrlhf_b200.engine")
    engine_qwen = importlib.import_module("vlrlhf_b200.engine_qwen")
    from vlrlhf_b200 import config, host
    yield config, engine_qwen, host, mock_ops
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


def _setup(qpkg, tag, loss_type="sigmoid", with_optimizer=False, **tc):
    config, EQ, host, ops = qpkg
    name, qcfg = CASES[tag]
    d = np.load(os.path.join(G, tag + ".npz"))
    eng = EQ.QwenVLDPOEngine(getattr(config, name), config.TrainConfig(loss_type=loss_type, learning_rate=1e-3, **tc),
                             device="cpu", with_optimizer=with_optimizer)
    eng.init_synthetic(int(d["seed"]))
    batch = Q.make_batch(qcfg, int(d["n_pairs"]), int(d["text_len"]), int(d["prompt_len"]), int(d["seed"]), ddpo_like=True)
    return eng, qcfg, d, batch


def test_qwen_config_and_weights_mirror_oracle(qpkg):
    config, EQ, host, ops = qpkg
    eng, qcfg, d, batch = _setup(qpkg, "g9_qwen_tiny")
    w, lora = Q.make_weights(qcfg, 0)
    st = eng.hf_state("policy")
    for k, v in lora.items():
        assert torch.equal(st[k].float(), v), k
    for k, v in w.items():
        if not k.startswith("transformer.visual."):
            assert torch.equal(st[k].float().reshape(v.shape), v), k
    assert set(eng.hf_state("ref")) == {k for k in w if not k.startswith("transformer.visual.")}
    assert set(eng.hf_state("grad")) == set(lora)
    # position tables: the host transforms equal the oracle's
    assert torch.equal(host.sincos_2d(qcfg.hidden, 4), Q.sincos_2d(qcfg.hidden, 4))
    t = torch.randn(256, 8)
    assert torch.equal(host.interpolate_pos_table(t, 64), Q.get_abs_pos(t, 64))


def test_qwen_merge_index_matches_reference_placement(qpkg):
    config, EQ, host, ops = qpkg
    qcfg = Q.TINY_QWEN
    batch = Q.make_batch(qcfg, 3, 60, 24, seed=3)
    cb = R.concatenated_inputs(batch)
    ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    m = ops.qwen_merge_index(ids, am, lb, qcfg.n_queries, 3, 1, qcfg.image_start_id)
    assert int(m.status) == 0 and m.S == 60
    src = m.src_map.view(6, 60).long()
    spans = Q.image_spans(qcfg, ids).tolist()
    for (i, a, b) in spans:
        assert b - a - 1 == qcfg.n_queries
        assert torch.equal(src[i, a + 1:b], -1 - ((i % 3) * qcfg.n_queries + torch.arange(qcfg.n_queries)))
        other = torch.ones(60, dtype=torch.bool); other[a + 1:b] = False
        assert torch.equal(src[i][other], ids[i][other])
    assert torch.equal(m.seqlens.long(), am.sum(-1)) and torch.equal(m.pos.view(6, 60)[0].long(), torch.arange(60))
    assert torch.equal(m.target.view(6, 59), torch.where(lb[:, 1:] == -100, torch.full_like(lb[:, 1:], -100), lb[:, 1:]))
    bad = ids.clone(); bad[0, 5] = qcfg.pad_token_id  # one placeholder missing inside the span -> malformed? no: still Q tokens
    bad[0, 2 + qcfg.n_queries] = 7                      # the </img> marker is gone
    assert int(ops.qwen_merge_index(bad, am, lb, qcfg.n_queries, 3, 1, qcfg.image_start_id).status) == 2


@pytest.mark.parametrize("tag", list(CASES))
def test_qwen_engine_forward_parity_cpu_mock(qpkg, tag):
    config, EQ, host, ops = qpkg
    eng, qcfg, d, batch = _setup(qpkg, tag)
    cb = host.concatenated_inputs(batch)
    ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    a = eng.prepare_inputs(ids, am, lb, cb["concatenated_img_input_dict"]["pixel_values"])
    # the resampler output against the oracle's visual tower
    w, lora = Q.make_weights(qcfg, int(d["seed"]))
    with torch.no_grad():
        want = Q.visual_forward(qcfg, w, batch["img_input_dict"]["pixel_values"])
    got = eng.vision_features(a[3]).float().view(want.shape)
    assert (got - want).abs().max() <= 3e-2 * want.abs().max()
    out = eng.step(*a, train=False)
    np.testing.assert_allclose(out.policy_logps.numpy(), d["policy_logps"], rtol=1e-3)
    np.testing.assert_allclose(out.ref_logps.numpy(), d["ref_logps"], rtol=1e-3)
    wt = eng.ddpo_weights(ids, am, lb)
    assert int(wt.sum()) > 0
    out = eng.step(*eng.prepare_inputs(ids, am, lb, cb["concatenated_img_input_dict"]["pixel_values"], wt), train=False)
    np.testing.assert_allclose(out.policy_logps.numpy(), d["policy_logps_ddpo"], rtol=1e-3, atol=1e-2)
    np.testing.assert_allclose(out.ref_logps.numpy(), d["ref_logps_ddpo"], rtol=1e-3, atol=1e-2)


def test_qwen_engine_adapter_gradients_match_oracle_autograd(qpkg):
    config, EQ, host, ops = qpkg
    grads = {}
    for ckpt in (False, True):
        eng, qcfg, d, batch = _setup(qpkg, "g9_qwen_tiny", activation_checkpointing=ckpt)
        metrics = eng.train_step(batch, train=True)
        grads[ckpt] = eng.grads.clone()
    assert torch.equal(grads[False], grads[True])  # recompute == keep
    got = {k: v.float() for k, v in eng.hf_state("grad").items()}
    w, lora = Q.make_weights(qcfg, int(d["seed"]))
    leaves = {k: v.clone().requires_grad_(True) for k, v in lora.items()}
    loss, want_metrics, _ = Q.get_batch_loss_metrics(qcfg, w, leaves, batch)
    loss.backward()
    for k, leaf in leaves.items():
        rel = (got[k] - leaf.grad).norm().item() / max(leaf.grad.norm().item(), 1e-12)
        assert rel < 5e-2, f"{k}: rel {rel}"
    assert abs(metrics["loss"] - float(loss.detach())) < 2e-3
    for k in ("rewards/chosen", "rewards/rejected", "logps/chosen", "logps/rejected", "logits/chosen", "logits/rejected"):
        assert abs(metrics[k] - float(want_metrics[k])) <= 2e-3 * max(1.0, abs(float(want_metrics[k]))), k


def test_qwen_engine_optimizer_updates_only_adapters(qpkg):
    config, EQ, host, ops = qpkg
    eng, qcfg, d, batch = _setup(qpkg, "g9_qwen_tiny", with_optimizer=True, weight_decay=0.05)
    base0 = eng.bparams.clone()
    vis0 = eng.vparams.clone()
    l0 = eng.train_step(batch, train=True)["loss"]
    for _ in range(4):
        l1 = eng.train_step(batch, train=True)["loss"]
    assert l1 < l0
    assert torch.equal(eng.bparams, base0) and torch.equal(eng.vparams, vis0)
    assert torch.equal(eng.params, eng.master.to(torch.bfloat16))
    assert eng.master.numel() == eng.layout.size  # optimizer state covers the adapters only


# ------------------------------------------------------------------------------------------
# plugin side: checkpoints, adapters, images named in the token stream
# ------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def qplugin(qpkg):
    for k in ("vlrlhf_b200.plugin", "vlrlhf_b200.plugin_qwen"):
        sys.modules.pop(k, None)
    import vlrlhf_b200
    for a in ("plugin", "plugin_qwen"):
        if hasattr(vlrlhf_b200, a):
            delattr(vlrlhf_b200, a)
    plugin = importlib.import_module("vlrlhf_b200.plugin")
    pq = importlib.import_module("vlrlhf_b200.plugin_qwen")
    yield plugin, pq
    for k in ("vlrlhf_b200.plugin", "vlrlhf_b200.plugin_qwen"):
        sys.modules.pop(k, None)
    for a in ("plugin", "plugin_qwen"):
        if hasattr(vlrlhf_b200, a):
            delattr(vlrlhf_b200, a)


def _spell(path: str, qcfg, total: int):
    b = list(path.encode("utf-8"))
    return [qcfg.image_start_id] + b + [qcfg.image_start_id + 2] * (total - len(b)) + [qcfg.image_start_id + 1]


def test_image_paths_from_ids_like_reference(qplugin):
    plugin, pq = qplugin
    qcfg = Q.TINY_QWEN
    ids = torch.full((3, 40), 7, dtype=torch.int64)
    paths = ["/tmp/a.png", "d/i_01.jpg", "x"]
    for i, p in enumerate(paths):
        blk = _spell(p, qcfg, qcfg.n_queries)
        ids[i, 2 + i:2 + i + len(blk)] = torch.tensor(blk)
    got = pq.image_paths_from_ids(ids, qcfg.image_start_id)
    # the reference's loop (modeling_qwen.py:524-534)
    bos, eos = torch.where(ids == qcfg.image_start_id), torch.where(ids == qcfg.image_start_id + 1)
    want = []
    for i, a, b in torch.stack((bos[0], bos[1], eos[1]), dim=1):
        image = ids[i][a + 1:b - 1].tolist()
        image = image[: image.index(qcfg.image_start_id + 2)]
        want.append(bytes(image).decode("utf-8"))
    assert got == want == paths


def test_qwen_from_pretrained_adapters_and_plugin_forward(qpkg, qplugin, tmp_path, monkeypatch):
    pytest.importorskip("safetensors")
    from safetensors.torch import save_file
    import json
    config, EQ, host, ops = qpkg
    plugin, pq = qplugin
    qcfg = Q.TINY_QWEN
    w, lora = Q.make_weights(qcfg, 0)
    src = tmp_path / "ckpt"
    src.mkdir()
    save_file({k: v.to(torch.bfloat16).contiguous() for k, v in w.items()}, str(src / "model.safetensors"))
    hf_cfg = dict(model_type="qwen", vocab_size=qcfg.vocab, hidden_size=qcfg.hidden, num_hidden_layers=qcfg.layers,
                  num_attention_heads=qcfg.heads, kv_channels=qcfg.head_dim, intermediate_size=2 * qcfg.ff, seq_length=2048,
                  layer_norm_epsilon=qcfg.rms_eps, rotary_emb_base=qcfg.rope_theta,
                  visual=dict(heads=qcfg.v_heads, image_size=qcfg.image_size, image_start_id=qcfg.image_start_id,
                              layers=qcfg.v_layers, mlp_ratio=qcfg.v_mlp / qcfg.v_width, output_dim=qcfg.hidden,
                              patch_size=qcfg.patch_size, width=qcfg.v_width, n_queries=qcfg.n_queries))
    (src / "config.json").write_text(json.dumps(hf_cfg))
    model = pq.B200QwenVLForRL.from_pretrained(str(src), torch_dtype=torch.bfloat16, device="cpu", lora_r=qcfg.lora_r,
                                               lora_alpha=qcfg.lora_alpha)
    mc = model.cfg
    for f in ("hidden", "layers", "heads", "ff", "vocab", "v_width", "v_layers", "v_heads", "v_mlp", "n_queries", "image_size",
              "image_start_id", "lora_r", "lora_alpha"):
        assert getattr(mc, f) == getattr(config.TINY_QWEN, f), f
    st = model.engine.hf_state("ref")
    for k, v in w.items():
        if not k.startswith("transformer.visual."):
            assert torch.equal(st[k].float().reshape(v.shape), v), k
    # frozen base, trainable adapters with gradient views
    trainable = {n for n, p in model._hf.items() if p.requires_grad}
    assert trainable == set(lora)
    # peft init (B = 0): the policy equals the reference
    batch = Q.make_batch(qcfg, 2, 48, 24, seed=1)
    cb = host.concatenated_inputs(batch)
    a = model.engine.prepare_inputs(cb["concatenated_input_ids"], cb["concatenated_attention_mask"], cb["concatenated_labels"],
                                    cb["concatenated_img_input_dict"]["pixel_values"])
    out = model.engine.step(*a, train=False)
    assert torch.allclose(out.policy_logps, out.ref_logps, rtol=0, atol=1e-3)
    with torch.no_grad():
        want = torch.cat(Q.concatenated_forward(qcfg, w, None, batch)[:2])
    np.testing.assert_allclose(out.ref_logps.numpy(), want.numpy(), rtol=1e-3)
    # adapters: write the synthetic ones in PEFT naming, load, save, reload
    views = model.engine.lora_views(model.engine.policy)
    for k, v in lora.items():
        views[k].copy_(v.to(torch.bfloat16))
    files = model.save_pretrained(str(tmp_path / "adapter"))
    assert files == ["adapter_model.safetensors", "adapter_config.json"]
    acfg = json.loads((tmp_path / "adapter" / "adapter_config.json").read_text())
    assert acfg["r"] == qcfg.lora_r and acfg["peft_type"] == "LORA" and sorted(acfg["target_modules"]) == sorted(model.default_lora_target)
    from safetensors import safe_open
    with safe_open(str(tmp_path / "adapter" / "adapter_model.safetensors"), "pt") as f:
        keys = set(f.keys())
        assert keys == {f"base_model.model.{k}.weight" for k in lora}
        k0 = "base_model.model.transformer.h.1.mlp.w1.lora_B.weight"
        assert torch.equal(f.get_tensor(k0).float(), lora["transformer.h.1.mlp.w1.lora_B"])
    model.reset_adapters()
    assert float(views["transformer.h.0.attn.c_attn.lora_B"].abs().max()) == 0.0
    model.load_adapter(str(tmp_path / "adapter"))
    for k, v in lora.items():
        assert torch.equal(views[k].float(), v), k
    # LoraConfig validation of the trainer hook
    from types import SimpleNamespace
    pq.check_peft_config(model, SimpleNamespace(r=qcfg.lora_r, lora_alpha=qcfg.lora_alpha, target_modules=["c_attn", "attn.c_proj", "w1", "w2"]))
    for bad in (None, SimpleNamespace(r=4, lora_alpha=qcfg.lora_alpha, target_modules=["c_attn", "attn.c_proj", "w1", "w2"]),
                SimpleNamespace(r=qcfg.lora_r, lora_alpha=qcfg.lora_alpha, target_modules=["c_attn"])):
        with pytest.raises(ValueError):
            pq.check_peft_config(model, bad)
    # the trainer-level call on a reference-format batch (no img_input_dict: images are named in the token stream)
    Image = pytest.importorskip("PIL.Image")
    from oracle import image_restate as IR
    imgs = [IR.synthetic_image(90, 120, 1), IR.synthetic_image(130, 80, 2)]
    b2 = {k: v.clone() for k, v in batch.items() if k != "img_input_dict"}
    monkeypatch.chdir(tmp_path)  # short relative names: the tiny config has only 16 placeholder tokens to spell a path
    for i, im in enumerate(imgs):
        path = f"q{i}.png"
        Image.fromarray(im).save(path)
        blk = torch.tensor(_spell(path, qcfg, qcfg.n_queries))
        for side in ("chosen", "rejected"):
            b2[f"{side}_input_ids"][i, 1:1 + len(blk)] = blk
    model._preprocessor = lambda arrs: torch.from_numpy(np.stack([IR.square_preprocess(x, qcfg.image_size) for x in arrs]))
    trainer = SimpleNamespace(loss_type="sigmoid", is_encoder_decoder=False, label_pad_token_id=-100, padding_value=0)
    with torch.no_grad():
        pc, pr, _, _ = plugin.concatenated_forward(trainer, model, b2)
        rc, rr, _, _ = plugin.concatenated_forward(trainer, plugin.RefView(model), b2)
    b3 = dict(b2)
    b3["img_input_dict"] = {"pixel_values": model._preprocessor(imgs)}
    with torch.no_grad():
        w_pc, w_pr, _, _, _ = Q.concatenated_forward(qcfg, w, lora, b3)
        w_rc, w_rr, _, _, _ = Q.concatenated_forward(qcfg, w, None, b3)
    np.testing.assert_allclose(torch.cat([pc, pr]).numpy(), torch.cat([w_pc, w_pr]).numpy(), rtol=1e-3)
    np.testing.assert_allclose(torch.cat([rc, rr]).numpy(), torch.cat([w_rc, w_rr]).numpy(), rtol=1e-3)


@pytest.mark.parametrize("loss_type,ckpt", [("sigmoid", False), ("ddpo", True)])
def test_qwen_packed_step_equals_padded_step(qpkg, loss_type, ckpt):
    """TrainConfig.pack_sequences on the Qwen-VL engine (S == L: the attended tokens survive): same log-probs, loss, rewards
    and adapter gradients as the padded step."""
    res = []
    for pack in (False, True):
        eng, qcfg, d, batch = _setup(qpkg, "g9_qwen_tiny", loss_type=loss_type, pack_sequences=pack, activation_checkpointing=ckpt)
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


@pytest.mark.parametrize("loss_type,ckpt", [("sigmoid", False), ("ddpo", True)])
def test_qwen_shared_prefix_step_equals_padded_step(qpkg, loss_type, ckpt):
    """TrainConfig.share_prefix on the Qwen-VL engine: the prompt incl. its 256 image placeholder tokens (S == L: a common token
    is a common row) is laid out once per pair: log-probs, loss and rewards of the padded step, adapter gradients equal up to the
    accumulation order."""
    res = []
    for share in (False, True):
        eng, qcfg, d, batch = _setup(qpkg, "g9_qwen_tiny", loss_type=loss_type, share_prefix=share, activation_checkpointing=ckpt)
        metrics = eng.train_step(batch, train=True)
        m = eng._saved["m"]
        assert m.shared == share
        if share:
            assert m.shared_rows >= int(d["n_pairs"]) * qcfg.n_queries and m.T < m.n_seq * m.S - m.shared_rows + 1
        res.append((metrics, eng.grads.clone()))
    (m0, g0), (m1, g1) = res
    for k in ("loss", "rewards/chosen", "rewards/rejected", "rewards/margins", "logps/chosen", "logps/rejected"):
        assert abs(m0[k] - m1[k]) <= 5e-5 * max(1.0, abs(m0[k])), (k, m0[k], m1[k])
    g0, g1 = g0.float(), g1.float()
    assert float(g0.abs().sum()) > 0 and float((g0 - g1).norm() / g0.norm()) < 2e-2
