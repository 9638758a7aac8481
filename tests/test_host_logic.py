"""CPU tests of the host logic: the engine's orchestration (run over tests/mock_ops.py, a plain-PyTorch
stand-in for the C ABI) against the oracle + golden vectors, the host-side DDPO mask, concatenated_inputs,
and the 2-rank gloo data-parallel path.  No CUDA kernel runs here."""
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
    """vlrlhf_b200 with `ops` replaced by the PyTorch mock (CPU)."""
    import vlrlhf_b200  # noqa: F401
    from tests import mock_ops
    saved = {k: sys.modules.get(k) for k in ("vlrlhf_b200.ops", "vlrlhf_b200.engine")}
    sys.modules["vlrlhf_b200.ops"] = mock_ops
    sys.modules.pop("vlrlhf_b200.engine", None)
    engine = importlib.import_module("vlrlhf_b200.engine")
    from vlrlhf_b200 import config, host
    yield config, engine, host, mock_ops
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


def _setup(cpu_pkg, tag="g4_tiny", loss_type="sigmoid", with_optimizer=False):
    config, engine, host, ops = cpu_pkg
    name, rcfg, npairs, tl, pl = {"g4_tiny": ("TINY", R.TINY, 2, 24, 8), "g4_small": ("SMALL", R.SMALL, 2, 96, 24),
                                  "g6_next_tiny": ("TINY_NEXT", R.TINY_NEXT, 3, 24, 8),
                                  "g6_next_small": ("SMALL_NEXT", R.SMALL_NEXT, 2, 96, 24)}[tag]
    d = np.load(os.path.join(G, tag + ".npz"))
    sizes = [tuple(x) for x in d["image_sizes"].tolist()] if "image_sizes" in d.files else None
    eng = engine.LlavaDPOEngine(getattr(config, name), config.TrainConfig(loss_type=loss_type, learning_rate=1e-3),
                                device="cpu", with_optimizer=with_optimizer)
    eng.init_synthetic(int(d["seed"]))
    batch = R.make_batch(rcfg, npairs, tl, pl, int(d["seed"]), ddpo_like=True, image_sizes=sizes)
    cb = host.concatenated_inputs(batch)
    return eng, rcfg, d, batch, cb


def test_config_mirrors_oracle_specs(cpu_pkg):
    config, engine, host, ops = cpu_pkg
    for a, b in ((config.TINY, R.TINY), (config.SMALL, R.SMALL), (config.LLAVA15_7B, R.LLAVA15_7B),
                 (config.TINY_NEXT, R.TINY_NEXT), (config.SMALL_NEXT, R.SMALL_NEXT),
                 (config.LLAVANEXT_MISTRAL_7B, R.LLAVANEXT_MISTRAL_7B)):
        assert config.weight_specs(a) == R.weight_specs(b)
        assert a.n_patches == b.n_patches and a.v_used_layers == b.v_used_layers
    assert config.tensor_seed("x.y", 3) == R.tensor_seed("x.y", 3)


def test_concatenated_inputs_matches_oracle(cpu_pkg):
    config, engine, host, ops = cpu_pkg
    batch = R.make_batch(R.TINY, 3, 20, 6, seed=1)
    batch["rejected_input_ids"] = batch["rejected_input_ids"][:, :17]  # ragged chosen/rejected lengths
    batch["rejected_attention_mask"] = batch["rejected_attention_mask"][:, :17]
    batch["rejected_labels"] = batch["rejected_labels"][:, :17]
    a, b = host.concatenated_inputs(batch, padding_value=7), R.concatenated_inputs(batch, padding_value=7)
    for k in ("concatenated_input_ids", "concatenated_attention_mask", "concatenated_labels"):
        assert torch.equal(a[k], b[k])
    assert torch.equal(a["concatenated_img_input_dict"]["pixel_values"], b["concatenated_img_input_dict"]["pixel_values"])
    with pytest.raises(ValueError):
        host.concatenated_inputs({**batch, "img_input_dict": {"x": 3}})


def test_get_diff_ids_golden(cpu_pkg):
    config, engine, host, ops = cpu_pkg
    d = np.load(os.path.join(G, "g3_ddpo.npz"))
    for c in sorted({k[:-2] for k in d.files if k.endswith("_a")}):
        ia, ib = host.get_diff_ids(d[c + "_a"].tolist(), d[c + "_b"].tolist(), 3)
        assert ia == d[c + "_ia"].tolist() and ib == d[c + "_ib"].tolist(), c


def test_native_matching_blocks_equal_difflib(cpu_pkg):
    """vlb200_host_matching_blocks (C++ restatement of difflib.SequenceMatcher incl. autojunk) == CPython difflib,
    bit-exact, on random edits, popular elements (image/prompt zeros) and the n >= 200 autojunk threshold."""
    import difflib
    import random
    config, engine, host, ops = cpu_pkg
    rnd = random.Random(0)
    for trial in range(200):
        n = rnd.choice([0, 1, 5, 30, 199, 200, 201, 400, 1600])
        vocab = rnd.choice([2, 5, 50, 1000])
        a = [rnd.randint(0, vocab) for _ in range(n)]
        b = list(a)
        for _ in range(rnd.randint(0, 10)):
            if b:
                s0 = rnd.randrange(len(b))
                b[s0:s0 + rnd.randint(1, 8)] = [rnd.randint(0, vocab) for _ in range(rnd.randint(0, 8))]
        if rnd.random() < 0.3:
            b = [rnd.randint(0, vocab) for _ in range(max(0, n + rnd.randint(-20, 20)))]
        if rnd.random() < 0.5:
            a, b = [0] * rnd.randint(0, 300) + a, [0] * rnd.randint(0, 300) + b
        want = [tuple(x) for x in difflib.SequenceMatcher(None, a, b).get_matching_blocks()]
        assert host.matching_blocks_native(a, b) == want, (trial, len(a), len(b))
    d = np.load(os.path.join(G, "g3_ddpo.npz"))  # the reference's get_diff_ids cases, through the native matcher
    for c in sorted({k[:-2] for k in d.files if k.endswith("_a")}):
        a, b = d[c + "_a"].tolist(), d[c + "_b"].tolist()
        assert host.matching_blocks_native(a, b) == [tuple(x) for x in difflib.SequenceMatcher(None, a, b).get_matching_blocks()]


def test_native_ddpo_row_weights_equal_python_mirror(cpu_pkg):
    config, engine, host, ops = cpu_pkg
    for seed in range(3):
        cfg = R.SMALL
        cb = R.concatenated_inputs(R.make_batch(cfg, 3, 120, 10, seed=seed, ddpo_like=True))
        ids, lb = cb["concatenated_input_ids"], cb["concatenated_labels"]
        assert torch.equal(host.ddpo_row_weights(ids, lb, cfg.image_token_index, cfg.n_patches),
                           host.ddpo_row_weights_native(ids, lb, cfg.image_token_index, cfg.n_patches))
        cfg, sizes = R.SMALL_NEXT, [(112, 112), (90, 300), (200, 100)]
        cb = R.concatenated_inputs(R.make_batch(cfg, 3, 80, 8, seed=seed, ddpo_like=True, image_sizes=sizes))
        ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
        plan = host.anyres_pack_index(sizes, cfg.image_grid_pinpoints, cfg.image_size, cfg.patch_size)
        S = host.next_merged_len(ids, am, plan.feature_lens, cfg.image_token_index)
        w1 = host.ddpo_row_weights(ids, lb, cfg.image_token_index, plan.feature_lens * 2, attention_mask=am, merged_len=S)
        w2 = host.ddpo_row_weights_native(ids, lb, cfg.image_token_index, plan.feature_lens * 2, attention_mask=am, merged_len=S)
        assert torch.equal(w1, w2) and int(w1.sum()) > 0
    with pytest.raises(ValueError):
        host.ddpo_row_weights_native(ids[:3], lb[:3], cfg.image_token_index, [1, 1, 1])


def test_ddpo_row_weights_equal_reference_mask(cpu_pkg):
    config, engine, host, ops = cpu_pkg
    cfg = R.TINY
    batch = R.make_batch(cfg, 3, 40, 8, seed=2, ddpo_like=True)
    cb = R.concatenated_inputs(batch)
    ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    w = host.ddpo_row_weights(ids, lb, cfg.image_token_index, cfg.n_patches)
    # reference semantics: mask over the merged shifted labels (trainer.py:161-184)
    emb = torch.zeros(cfg.vocab, 8)
    img = torch.ones(6, cfg.n_patches, 8)
    _, _, fl, _, _ = R.merge_input_ids_with_image_features(cfg, img, torch.nn.functional.embedding(ids, emb) + 1, ids, am, lb)
    shift = fl[:, 1:].clone()
    shift[shift == -100] = 0
    mask = R.ddpo_shared_mask(shift)
    m = ops.llava_merge_index(ids, am, lb, cfg.n_patches, 3, 1, cfg.image_token_index, cfg.pad_token_id)
    rows = m.row_of_text.view(6, -1).long() - (torch.arange(6) * m.S)[:, None]
    want = torch.gather(mask, 1, rows.clamp(min=0)).to(torch.uint8)
    tgt = m.target.view(6, -1)
    assert torch.equal(w[tgt >= 0], want[tgt >= 0])
    assert int(w.sum()) > 0


@pytest.mark.parametrize("tag", ["g4_tiny", "g4_small"])
def test_engine_forward_parity_cpu_mock(cpu_pkg, tag):
    config, engine, host, ops = cpu_pkg
    eng, rcfg, d, batch, cb = _setup(cpu_pkg, tag)
    wp, wr = R.make_policy_and_ref(rcfg, int(d["seed"]))
    pol = eng.hf_state("policy")
    for k, v in wp.items():
        if k in pol:
            assert torch.equal(pol[k].float().reshape(v.shape), v), k
    ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    a = eng.prepare_inputs(ids, am, lb, cb["concatenated_img_input_dict"]["pixel_values"])
    out = eng.step(*a[:4], train=False)
    np.testing.assert_allclose(out.policy_logps.numpy(), d["policy_logps"], rtol=1e-3)
    np.testing.assert_allclose(out.ref_logps.numpy(), d["ref_logps"], rtol=1e-3)
    # DDPO path
    wt = host.ddpo_row_weights(ids, lb, rcfg.image_token_index, rcfg.n_patches)
    out = eng.step(*a[:4], ddpo_weight=wt.reshape(-1), train=False)
    np.testing.assert_allclose(out.policy_logps.numpy(), d["policy_logps_ddpo"], rtol=1e-3, atol=1e-2)


def test_engine_backward_parity_cpu_mock(cpu_pkg):
    config, engine, host, ops = cpu_pkg
    eng, rcfg, d, batch, cb = _setup(cpu_pkg, "g4_tiny")
    ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    a = eng.prepare_inputs(ids, am, lb, cb["concatenated_img_input_dict"]["pixel_values"])
    eng.step(*a[:4], train=True)
    got = {k: v.float() for k, v in eng.hf_state("grad").items()}
    wp, wr = R.make_policy_and_ref(rcfg, int(d["seed"]))
    leaves = {k: v.clone().requires_grad_(True) for k, v in wp.items() if not k.startswith("vision_tower.")}
    w = dict(wp)
    w.update(leaves)
    loss, _, _ = R.get_batch_loss_metrics(rcfg, w, wr, batch)
    loss.backward()
    for k, leaf in leaves.items():
        g, want = got[k].reshape(leaf.shape), leaf.grad
        rel = (g - want).norm().item() / max(want.norm().item(), 1e-12)
        assert rel < 5e-2, f"{k}: rel {rel}"


# ------------------------------------------------------------------------------------------
# LLaVA-Next: anyres packing index, merge, DDPO mask and the engine over the mock ops
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rcfg", [R.TINY_NEXT, R.SMALL_NEXT, R.LLAVANEXT_MISTRAL_7B])
def test_anyres_pack_index_equals_reference_packing(cpu_pkg, rcfg):
    """pack_image_features restated with tensors (oracle) on features that carry their own row number == the
    integer index the product builds (bit-exact), over square / wide / tall / extreme aspect ratios."""
    config, engine, host, ops = cpu_pkg
    P = rcfg.n_patches
    s = rcfg.image_size
    sizes = [(s, s), (20, 50), (60, 25), (100, 100), (37, 211), (500, 90), (333, 334), (640, 480), (480, 640), (1000, 300),
             (17, 17), (2 * s, 2 * s), (s, 3 * s), (3 * s, s)]
    plan = host.anyres_pack_index(sizes, rcfg.image_grid_pinpoints, rcfg.image_size, rcfg.patch_size)
    crops = [R.image_size_to_num_patches(x, rcfg.image_grid_pinpoints, rcfg.image_size) for x in sizes]
    assert crops == plan.crops
    feats = torch.arange(sum(crops) * P, dtype=torch.float32).reshape(sum(crops), P, 1)
    packed, lens = R.pack_image_features(rcfg, list(torch.split(feats, crops, 0)), sizes,
                                         torch.tensor([float(plan.n_crop_rows)]))
    assert lens.tolist() == plan.feature_lens and plan.total_feats == int(lens.sum())
    assert torch.equal(packed.flatten().to(torch.int32), plan.pack_index)
    nl = plan.pack_index == plan.n_crop_rows
    assert torch.equal(plan.scatter_index[~nl], plan.pack_index[~nl]) and bool((plan.scatter_index[nl] == -1).all())
    assert torch.equal(plan.newline_rows.long(), torch.nonzero(nl).flatten())
    assert plan.scatter_index[~nl].unique().numel() == int((~nl).sum())  # every crop row is used at most once


def test_next_merge_index_equals_reference_merge(cpu_pkg):
    config, engine, host, ops = cpu_pkg
    cfg = R.TINY_NEXT
    sizes = [(28, 28), (20, 50), (60, 25)]
    batch = R.make_batch(cfg, 3, 30, 8, seed=5, image_sizes=sizes)
    cb = R.concatenated_inputs(batch)
    ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    plan = host.anyres_pack_index(sizes, cfg.image_grid_pinpoints, cfg.image_size, cfg.patch_size)
    S = host.next_merged_len(ids, am, plan.feature_lens, cfg.image_token_index)
    m = ops.llavanext_merge_index(ids, am, lb, plan.feat_off, plan.total_feats, S, 3, 1, cfg.image_token_index)
    assert int(m.status) == 0
    lens2 = torch.tensor(plan.feature_lens * 2)
    d = 4
    feats = torch.arange(1, 2 * plan.total_feats + 1, dtype=torch.float32)[:, None].expand(-1, d) * -1.0
    # the reference sees the duplicated image batch [v, v]: feature rows of the second copy differ only by offset
    emb_w = torch.arange(cfg.vocab, dtype=torch.float32)[:, None].expand(-1, d) + 1.0
    ids0 = ids.clone(); ids0[ids == cfg.image_token_index] = 0
    fe, fm, fl, pos, imap = R.next_merge_input_ids_with_image_features(
        cfg, feats, lens2, torch.nn.functional.embedding(ids0, emb_w), ids, am, lb)
    assert fe.shape[1] == S == m.S
    assert torch.equal(m.labels, fl) and torch.equal(m.mask.long(), fm)
    assert torch.equal(m.pos.view(6, S).long(), pos)
    src = m.src_map.view(6, S).long()
    INT_MIN = -(2 ** 31)
    assert torch.equal(src < 0, imap | (src == INT_MIN)) and torch.equal((src < 0) & (src != INT_MIN), imap)
    # text rows hold the token's embedding row; image rows hold packed feature row k (second copy: k + total)
    want = fe[..., 0]
    got = torch.where(src >= 0, src.float() + 1.0, torch.zeros_like(want))
    k = (-1 - src).clamp(min=0)
    rep = (torch.arange(6) // 3)[:, None]
    got = torch.where(imap, -(k + rep * plan.total_feats + 1).float(), got)
    assert torch.equal(got, want)
    assert torch.equal(m.seqlens.long(), fm.sum(-1))
    # img_rows inverse map
    rows = m.img_pos.view(2, plan.total_feats).long()
    assert torch.equal(m.src_map.long()[rows], (-1 - torch.arange(plan.total_feats))[None].expand(2, -1))
    # wrong image count / left padding are flagged like the reference's ValueErrors
    bad = ids.clone(); bad[0, 3] = cfg.image_token_index
    assert int(ops.llavanext_merge_index(bad, am, lb, plan.feat_off, plan.total_feats, S + 20, 3, 1,
                                         cfg.image_token_index).status) == 2
    am2 = am.clone(); am2[1, 0] = 0; am2[:, -1] = 0
    assert int(ops.llavanext_merge_index(ids, am2, lb, plan.feat_off, plan.total_feats, S, 3, 1,
                                         cfg.image_token_index).status) == 3


def test_next_ddpo_row_weights_equal_reference_mask(cpu_pkg):
    config, engine, host, ops = cpu_pkg
    cfg = R.SMALL_NEXT
    sizes = [(112, 112), (90, 300), (200, 100)]
    batch = R.make_batch(cfg, 3, 60, 8, seed=2, ddpo_like=True, image_sizes=sizes)
    cb = R.concatenated_inputs(batch)
    ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    plan = host.anyres_pack_index(sizes, cfg.image_grid_pinpoints, cfg.image_size, cfg.patch_size)
    S = host.next_merged_len(ids, am, plan.feature_lens, cfg.image_token_index)
    w = host.ddpo_row_weights(ids, lb, cfg.image_token_index, plan.feature_lens * 2, attention_mask=am, merged_len=S)
    feats = torch.ones(2 * plan.total_feats, 2)
    ids0 = ids.clone(); ids0[ids == cfg.image_token_index] = 0
    _, _, fl, _, _ = R.next_merge_input_ids_with_image_features(cfg, feats, torch.tensor(plan.feature_lens * 2),
                                                                torch.ones(*ids.shape, 2), ids, am, lb)
    shift = fl[:, 1:].clone()
    shift[shift == -100] = 0
    mask = R.ddpo_shared_mask(shift)
    m = ops.llavanext_merge_index(ids, am, lb, plan.feat_off, plan.total_feats, S, 3, 1, cfg.image_token_index)
    rows = m.row_of_text.view(6, -1).long() - (torch.arange(6) * m.S)[:, None]
    want = torch.gather(mask, 1, rows.clamp(min=0)).to(torch.uint8)
    tgt = m.target.view(6, -1)
    assert torch.equal(w[tgt >= 0], want[tgt >= 0])
    assert int(w.sum()) > 0


@pytest.mark.parametrize("tag", ["g6_next_tiny", "g6_next_small"])
def test_next_engine_forward_parity_cpu_mock(cpu_pkg, tag):
    config, engine, host, ops = cpu_pkg
    eng, rcfg, d, batch, cb = _setup(cpu_pkg, tag)
    wp, wr = R.make_policy_and_ref(rcfg, int(d["seed"]))
    pol = eng.hf_state("policy")
    assert "image_newline" in pol
    for k, v in wp.items():
        if k in pol:
            assert torch.equal(pol[k].float().reshape(v.shape), v), k
    ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    img = cb["concatenated_img_input_dict"]
    a = eng.prepare_inputs(ids, am, lb, img["pixel_values"], None, img["image_sizes"])
    assert len(a) == 6
    out = eng.step(*a, train=False)
    assert np.array_equal(eng_labels(eng, a), d["labels"])
    np.testing.assert_allclose(out.policy_logps.numpy(), d["policy_logps"], rtol=1e-3)
    np.testing.assert_allclose(out.ref_logps.numpy(), d["ref_logps"], rtol=1e-3)
    wt = eng.ddpo_weights(ids, am, lb, img["image_sizes"])
    a = eng.prepare_inputs(ids, am, lb, img["pixel_values"], wt, img["image_sizes"])
    out = eng.step(*a, train=False)
    np.testing.assert_allclose(out.policy_logps.numpy(), d["policy_logps_ddpo"], rtol=1e-3, atol=1e-2)
    # the flat-crop form of pixel_values (4-D) is accepted too
    crops = a[5].crops
    flat = torch.cat([pv[:c] for pv, c in zip(batch["img_input_dict"]["pixel_values"], crops)], 0)
    b = eng.prepare_inputs(ids, am, lb, flat, wt, batch["img_input_dict"]["image_sizes"])
    assert torch.equal(b[3], a[3])
    with pytest.raises(ValueError):
        eng.prepare_inputs(ids, am, lb, flat[:-1], wt, batch["img_input_dict"]["image_sizes"])
    with pytest.raises(ValueError):
        eng.prepare_inputs(ids, am, lb, flat, wt)


def eng_labels(eng, a):
    from tests import mock_ops
    plan = a[5]
    m = mock_ops.llavanext_merge_index(a[0], a[1], a[2], plan.feat_off, plan.total_feats, plan.merged_len, len(plan.crops), 1,
                                       eng.cfg.image_token_index)
    return m.labels.numpy()


def test_next_engine_backward_parity_cpu_mock(cpu_pkg):
    config, engine, host, ops = cpu_pkg
    eng, rcfg, d, batch, cb = _setup(cpu_pkg, "g6_next_tiny")
    metrics = eng.train_step(batch, train=True)
    got = {k: v.float() for k, v in eng.hf_state("grad").items()}
    wp, wr = R.make_policy_and_ref(rcfg, int(d["seed"]))
    leaves = {k: v.clone().requires_grad_(True) for k, v in wp.items() if not k.startswith("vision_tower.")}
    w = dict(wp)
    w.update(leaves)
    loss, want_metrics, _ = R.get_batch_loss_metrics(rcfg, w, wr, batch)
    loss.backward()
    assert "image_newline" in leaves and leaves["image_newline"].grad.abs().max() > 0
    for k, leaf in leaves.items():
        g, want = got[k].reshape(leaf.shape), leaf.grad
        rel = (g - want).norm().item() / max(want.norm().item(), 1e-12)
        assert rel < 5e-2, f"{k}: rel {rel}"
    assert abs(metrics["loss"] - float(loss.detach())) < 2e-3
    for k in ("rewards/chosen", "rewards/rejected", "logps/chosen", "logps/rejected", "logits/chosen", "logits/rejected"):
        assert abs(metrics[k] - float(want_metrics[k])) <= 2e-3 * max(1.0, abs(float(want_metrics[k]))), k


def test_precomputed_reference_logps_skip_the_ref_pass(cpu_pkg):
    """TRL precompute_ref_log_probs: `reference_*_logps` in the batch replace the no-grad reference forward."""
    config, engine, host, ops = cpu_pkg
    eng, rcfg, d, batch, cb = _setup(cpu_pkg, "g4_tiny")
    base = eng.train_step(batch, train=False)
    n0 = ops.launch_count()
    eng.train_step(batch, train=False)
    full = ops.launch_count() - n0
    b2 = dict(batch)
    b2["reference_chosen_logps"] = torch.from_numpy(d["ref_logps"][:2])
    b2["reference_rejected_logps"] = torch.from_numpy(d["ref_logps"][2:])
    n0 = ops.launch_count()
    got = eng.train_step(b2, train=False)
    assert ops.launch_count() - n0 < 0.75 * full  # the reference decoder pass is gone
    for k in ("loss", "rewards/chosen", "rewards/rejected", "rewards/margins"):
        assert abs(got[k] - base[k]) < 2e-3, k
    with pytest.raises(ValueError):
        a = eng.prepare_inputs(*(cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels")),
                               cb["concatenated_img_input_dict"]["pixel_values"])
        eng.step(*a, train=False, ref_logps=torch.zeros(3))
    # the producer half (trl compute_reference_log_probs): one no-grad reference pass, values == the step's own reference pass
    rc, rr = eng.compute_reference_log_probs(batch)
    assert rc.device.type == "cpu" and rc.shape == rr.shape == (2,)
    np.testing.assert_allclose(torch.cat([rc, rr]).numpy(), d["ref_logps"], rtol=1e-3)
    b3 = dict(batch, reference_chosen_logps=rc, reference_rejected_logps=rr)
    again = eng.train_step(b3, train=False)
    for k in ("loss", "rewards/chosen", "rewards/rejected", "rewards/margins", "rewards/accuracies"):
        assert again[k] == base[k], k


def test_activation_checkpointing_gives_identical_gradients(cpu_pkg):
    config, engine, host, ops = cpu_pkg
    grads = []
    for ckpt in (False, True):
        eng, rcfg, d, batch, cb = _setup(cpu_pkg, "g4_tiny")
        eng.tc.activation_checkpointing = ckpt
        ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
        out = eng.step(*eng.prepare_inputs(ids, am, lb, cb["concatenated_img_input_dict"]["pixel_values"]), train=True)
        grads.append((eng.grads.clone(), out.policy_logps.clone()))
        saved = [k for k in eng._bufs if k.startswith("a.qkv")]
        assert (len(saved) == 0) == ckpt  # per-layer activations are not kept when checkpointing
    assert torch.equal(grads[0][0], grads[1][0]) and torch.equal(grads[0][1], grads[1][1])


def test_engine_optimizer_cpu_mock(cpu_pkg):
    config, engine, host, ops = cpu_pkg
    eng, rcfg, d, batch, cb = _setup(cpu_pkg, "g4_tiny", with_optimizer=True)
    ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    a = eng.prepare_inputs(ids, am, lb, cb["concatenated_img_input_dict"]["pixel_values"])
    l0 = float(eng.step(*a[:4], train=True).stats[0])
    for _ in range(3):
        l1 = float(eng.step(*a[:4], train=True).stats[0])
    assert l1 < l0
    assert torch.equal(eng.params, eng.master.to(torch.bfloat16))


# ------------------------------------------------------------------------------------------
# plugin API (the reference's three VLDPOTrainer override points) over the mock ops
# ------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def cpu_plugin(cpu_pkg):
    sys.modules.pop("vlrlhf_b200.plugin", None)
    return importlib.import_module("vlrlhf_b200.plugin")


def _ref_fns():
    from oracle import ref_shim
    if ref_shim.reference_available():
        T, _, _ = ref_shim.reference_symbols()
        return T.get_batch_logps, T.dpo_loss, "reference"
    return (lambda lg, lb, **kw: R.get_batch_logps(lg, lb, **{k: v for k, v in kw.items() if k != "is_encoder_decoder"}),
            lambda s, a, b, c, d: R.dpo_loss(a, b, c, d, s.beta, s.label_smoothing, s.loss_type, s.reference_free), "oracle")


def test_plugin_get_batch_logps_like_reference(cpu_plugin):
    ref_logps, _, _ = _ref_fns()
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(4, 30, 97, generator=g) * 2
    labels = torch.randint(1, 97, (4, 30), generator=g)
    labels[:, :7] = -100
    labels[1, 22:] = -100
    labels[2:] = labels[:2]
    labels[2, 12:15] = torch.tensor([5, 6, 7])  # a modified span so the DDPO mask is non-trivial
    for kw in (dict(), dict(average_log_prob=True), dict(mask_shared_tokens=True)):
        a = logits.clone().requires_grad_(True)
        b = logits.clone().requires_grad_(True)
        got = cpu_plugin.get_batch_logps(a, labels, **kw)
        want = ref_logps(b, labels, **kw)
        np.testing.assert_allclose(got.detach().numpy(), want.detach().numpy(), rtol=1e-5, atol=1e-4)
        w = torch.tensor([1.0, -2.0, 0.5, 3.0])
        (got * w).sum().backward()
        (want * w).sum().backward()
        np.testing.assert_allclose(a.grad.numpy(), b.grad.numpy(), rtol=2e-2, atol=2e-3)  # bf16 dlogits
    with pytest.raises(ValueError):
        cpu_plugin.get_batch_logps(logits, labels[:, :-1])


def test_plugin_dpo_loss_like_reference(cpu_plugin):
    from types import SimpleNamespace
    _, ref_loss, _ = _ref_fns()
    g = torch.Generator().manual_seed(1)
    pc, pr = -torch.rand(5, generator=g) * 300, -torch.rand(5, generator=g) * 300
    rc, rr = pc + torch.randn(5, generator=g) * 3, pr + torch.randn(5, generator=g) * 3
    for lt in ("sigmoid", "ddpo", "hinge", "ipo", "kto_pair"):
        for ls in (0.0, 0.1):
            s = SimpleNamespace(beta=0.1, label_smoothing=ls, loss_type=lt, reference_free=False,
                                accelerator=SimpleNamespace(device="cpu"))
            a, b = pc.clone().requires_grad_(True), pr.clone().requires_grad_(True)
            c, d = pc.clone().requires_grad_(True), pr.clone().requires_grad_(True)
            l1, cr1, rr1 = cpu_plugin.dpo_loss(s, a, b, rc, rr)
            l2, cr2, rr2 = ref_loss(s, c, d, rc, rr)
            np.testing.assert_allclose(l1.detach().numpy(), l2.detach().numpy(), rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(cr1.numpy(), cr2.numpy(), rtol=1e-6)
            l1.mean().backward()
            l2.mean().backward()
            np.testing.assert_allclose(a.grad.numpy(), c.grad.numpy(), rtol=1e-4, atol=1e-7)
            np.testing.assert_allclose(b.grad.numpy(), d.grad.numpy(), rtol=1e-4, atol=1e-7)
    with pytest.raises(ValueError):
        cpu_plugin.dpo_loss(SimpleNamespace(beta=0.1, label_smoothing=0, loss_type="x", reference_free=False), pc, pr, rc, rr)


@pytest.mark.parametrize("pack", [False, True])
def test_plugin_concatenated_forward_and_module(cpu_plugin, cpu_pkg, pack):
    from types import SimpleNamespace
    config, engine, host, ops = cpu_pkg
    d = np.load(os.path.join(G, "g4_tiny.npz"))
    model = cpu_plugin.B200LlavaForRL(config.TINY, config.TrainConfig(pack_sequences=pack), device="cpu", with_optimizer=False)
    model.engine.init_synthetic(int(d["seed"]))
    names = dict(model.hf_named_parameters())
    assert "language_model.model.layers.0.self_attn.q_proj.weight" in names
    assert not names["vision_tower.vision_model.embeddings.class_embedding"].requires_grad
    batch = R.make_batch(R.TINY, 2, 24, 8, int(d["seed"]), ddpo_like=True)
    trainer = SimpleNamespace(loss_type="sigmoid", is_encoder_decoder=False, label_pad_token_id=-100, padding_value=0,
                              beta=0.1, label_smoothing=0.0, reference_free=False)
    pc, pr, _, _ = cpu_plugin.concatenated_forward(trainer, model, batch)
    with torch.no_grad():
        rc, rr, _, _ = cpu_plugin.concatenated_forward(trainer, cpu_plugin.RefView(model), batch)
    np.testing.assert_allclose(torch.cat([pc, pr]).detach().numpy(), d["policy_logps"], rtol=1e-3)
    np.testing.assert_allclose(torch.cat([rc, rr]).numpy(), d["ref_logps"], rtol=1e-3)
    losses, cr, rj = cpu_plugin.dpo_loss(trainer, pc, pr, rc, rr)
    np.testing.assert_allclose(losses.detach().numpy(), d["sigmoid_losses"], atol=2e-3)
    losses.mean().backward()  # autograd -> engine backward -> .grad views of the flat gradient buffer
    gq = names["language_model.model.layers.0.self_attn.q_proj.weight"].grad
    assert gq is not None and float(gq.float().abs().sum()) > 0
    assert model.engine._saved["m"].packed == pack
    assert gq.data_ptr() == model.engine.hf_state("grad")["language_model.model.layers.0.self_attn.q_proj.weight"].data_ptr()


def test_train_step_metrics_match_oracle(cpu_pkg):
    """engine.train_step returns TRL's metric keys; values (incl. the logits/* means computed without logits) vs the oracle."""
    config, engine, host, ops = cpu_pkg
    eng, rcfg, d, batch, cb = _setup(cpu_pkg, "g4_tiny")
    got = eng.train_step(batch, train=True)
    wp, wr = R.make_policy_and_ref(rcfg, int(d["seed"]))
    with torch.no_grad():
        loss, metrics, aux = R.get_batch_loss_metrics(rcfg, wp, wr, batch)
    assert abs(got["loss"] - float(loss.detach())) < 2e-3
    for k in ("rewards/chosen", "rewards/rejected", "rewards/margins", "logps/chosen", "logps/rejected"):
        assert abs(got[k] - float(metrics[k])) < 2e-3 * max(1.0, abs(float(metrics[k]))), k
    assert got["rewards/accuracies"] == float(metrics["rewards/accuracies"])
    for k in ("logits/chosen", "logits/rejected"):
        assert abs(got[k] - float(metrics[k])) < 2e-3, (k, got[k], float(metrics[k]))


# ------------------------------------------------------------------------------------------
# collator (base/collator.py:26-68 + Llava/__init__.py:435-443) against the reference's own class
# ------------------------------------------------------------------------------------------
def test_collator_matches_reference_collator(tmp_path):
    from oracle import ref_shim, image_restate as IR
    ref_shim.install()
    from vlrlhf.base.collator import VLDPODataCollatorWithPadding
    import vlrlhf_b200  # noqa: F401
    from vlrlhf_b200.collator import B200DPODataCollatorWithPadding, load_rgb
    Image = pytest.importorskip("PIL.Image")
    rs = np.random.RandomState(0)
    feats = []
    for i, (pl, cl, rl) in enumerate([(5, 12, 9), (8, 7, 14), (3, 20, 20)]):
        path = str(tmp_path / f"im{i}.png")
        Image.fromarray(IR.synthetic_image(60 + 10 * i, 90 - 7 * i, i)).save(path)
        f = {"prompt": f"p{i}", "img_path": path, "reference_chosen_logps": -1.5 * i, "reference_rejected_logps": -2.0 - i}
        for name, n in (("prompt", pl), ("chosen", cl), ("rejected", rl)):
            f[f"{name}_input_ids"] = rs.randint(3, 300, n).tolist()
            f[f"{name}_attention_mask"] = [1] * n
            if name != "prompt":
                f[f"{name}_labels"] = [-100] * pl + f[f"{name}_input_ids"][pl:]
        feats.append(f)
    want = VLDPODataCollatorWithPadding(pad_token_id=301, label_pad_token_id=-100)(feats)
    fake_pre = lambda imgs: torch.from_numpy(np.stack([IR.clip_preprocess(im, 48, 48) for im in imgs]))  # noqa: E731
    got = B200DPODataCollatorWithPadding(pad_token_id=301, label_pad_token_id=-100, preprocessor=fake_pre)(feats)
    assert set(want) | {"img_input_dict"} == set(got)
    for k, v in want.items():
        if isinstance(v, torch.Tensor):
            assert torch.equal(got[k], v) and got[k].dtype == v.dtype, k
        else:
            assert got[k] == v, k
    # prompt is left padded, responses right padded
    assert got["prompt_input_ids"][2, 0] == 301 and got["chosen_input_ids"][1, -1] == 301
    # image step: decoded RGB of the same files through the (here CPU stand-in) preprocessor
    assert got["img_input_dict"]["pixel_values"].shape == (3, 3, 48, 48)
    assert np.array_equal(load_rgb(feats[1]["img_path"]), IR.synthetic_image(70, 83, 1))
    with pytest.raises(ValueError):
        B200DPODataCollatorWithPadding(is_encoder_decoder=True)


# ------------------------------------------------------------------------------------------
# HF checkpoint interchange (f-4): from_pretrained / save_pretrained of the plugin model
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("family", ["llava", "llava_next"])
def test_hf_checkpoint_roundtrip(cpu_pkg, cpu_plugin, tmp_path, family):
    transformers = pytest.importorskip("transformers")
    pytest.importorskip("safetensors")
    from oracle import make_fixtures as MF
    config, engine, host, ops = cpu_pkg
    rcfg = R.TINY if family == "llava" else R.TINY_NEXT
    if family == "llava":
        hf = transformers.LlavaForConditionalGeneration(MF.hf_config(rcfg))
    else:
        hf = transformers.LlavaNextForConditionalGeneration(MF.hf_config_next(rcfg))
    hf = hf.to(torch.bfloat16)
    src = str(tmp_path / "src")
    hf.save_pretrained(src, safe_serialization=True)
    model = cpu_plugin.B200LlavaForRL.from_pretrained(src, torch_dtype=torch.bfloat16, device="cpu", with_optimizer=True)
    want_cfg = getattr(config, "TINY" if family == "llava" else "TINY_NEXT")
    for f in ("hidden", "layers", "heads", "kv_heads", "ff", "vocab", "v_hidden", "v_layers", "image_size", "patch_size",
              "image_token_index", "family", "image_grid_pinpoints", "rope_theta", "rms_eps", "vision_feature_layer"):
        assert getattr(model.cfg, f) == getattr(want_cfg, f), f
    from vlrlhf_b200 import checkpoint
    sd = {checkpoint.legacy_name(k): v for k, v in hf.state_dict().items()}
    pol, ref = model.engine.hf_state("policy"), model.engine.hf_state("ref")
    assert len(pol) > 30
    for k, t in pol.items():
        assert torch.equal(t.reshape(sd[k].shape), sd[k]), k
        if not k.startswith("vision_tower."):
            assert torch.equal(ref[k].reshape(sd[k].shape), sd[k]), k
    assert torch.equal(model.engine.master, model.engine.params.float())
    unused = set(sd) - set(pol)
    assert unused and all(("encoder.layers" in k) or ("post_layernorm" in k) for k in unused)
    # one optimizer step, then export: the saved policy is complete and carries the update
    eng = model.engine
    eng.tc.learning_rate = 1e-3
    sizes = [(28, 28), (20, 50)] if family == "llava_next" else None
    eng.train_step(R.make_batch(rcfg, 2, 24, 8, 0, image_sizes=sizes), train=True)
    out = str(tmp_path / "out")
    files = model.save_pretrained(out, max_shard_size=200_000)  # force several shards + an index
    assert len(files) > 1 and os.path.exists(os.path.join(out, "model.safetensors.index.json"))
    saved = dict(checkpoint.iter_checkpoint(out))
    assert set(saved) == set(sd)  # 4.41 names, nothing missing
    changed = 0
    for k, t in saved.items():
        assert t.dtype == torch.bfloat16 and tuple(t.shape) == tuple(sd[k].shape), k
        if k in pol and not k.startswith("vision_tower."):
            assert torch.equal(t, eng.hf_state("policy")[k].reshape(t.shape))
            changed += int(not torch.equal(t, sd[k]))
        else:
            assert torch.equal(t, sd[k]), k  # frozen tower + unused tensors written back unchanged
    assert changed > 10
    # the reference's loader (HF from_pretrained) accepts the directory
    cls = transformers.LlavaForConditionalGeneration if family == "llava" else transformers.LlavaNextForConditionalGeneration
    back, info = cls.from_pretrained(out, torch_dtype=torch.bfloat16, output_loading_info=True)
    assert not info["missing_keys"] and not info["unexpected_keys"], info
    k = "language_model.model.layers.1.mlp.down_proj.weight"
    assert torch.equal({checkpoint.legacy_name(n): v for n, v in back.state_dict().items()}[k], saved[k])
    with pytest.raises(ValueError):
        checkpoint.config_from_hf({"model_type": "qwen"})


# ------------------------------------------------------------------------------------------
# packed rows (TrainConfig.pack_sequences) over the mock ops
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,ckpt,loss_type", [("g4_tiny", False, "sigmoid"), ("g4_tiny", True, "ddpo"),
                                                ("g6_next_tiny", False, "sigmoid"), ("g6_next_tiny", True, "ddpo"),
                                                ("g4_tiny", False, "kto_pair")])
def test_packed_step_equals_padded_step(cpu_pkg, tag, ckpt, loss_type):
    """Dropping the padding rows changes nothing a DPO step returns except the `logits/*` means: log-probs, loss, rewards and
    every gradient are those of the padded batch (and so within 1e-3 of the reference fixtures)."""
    config, engine, host, ops = cpu_pkg
    res = []
    for pack in (False, True):
        eng, rcfg, d, batch, cb = _setup(cpu_pkg, tag, loss_type=loss_type)
        eng.tc.pack_sequences, eng.tc.activation_checkpointing = pack, ckpt
        n_pad = int((batch["chosen_attention_mask"] == 0).sum() + (batch["rejected_attention_mask"] == 0).sum())
        assert n_pad > 0  # the fixture batch is ragged
        metrics = eng.train_step(batch, train=True)
        m = eng._saved["m"]
        assert m.packed == pack
        if pack:
            assert m.T < m.n_seq * m.S and m.T == sum(eng.host_seq_lens(cb["concatenated_input_ids"], cb["concatenated_attention_mask"],
                                                                          batch["img_input_dict"].get("image_sizes")))
            assert eng._bufs["x.0"].shape[0] == m.T and eng._stores["x.0"].numel() == m.n_seq * m.S * eng.cfg.hidden
        res.append((metrics, eng.grads.clone()))
    (m0, g0), (m1, g1) = res
    for k in m0:
        if not k.startswith("logits/"):
            assert m0[k] == m1[k], k
    assert torch.equal(g0, g1)
    key = "policy_logps_ddpo" if loss_type == "ddpo" else "policy_logps"
    np.testing.assert_allclose([m1["logps/chosen"], m1["logps/rejected"]],
                               [d[key][:len(d[key]) // 2].mean(), d[key][len(d[key]) // 2:].mean()], rtol=1e-3, atol=1e-2)


def test_packed_rows_shrink_and_grow_between_batches(cpu_pkg):
    """Workspaces are views of storage reserved at the padded size: a longer batch after a shorter one reuses it."""
    config, engine, host, ops = cpu_pkg
    eng, rcfg, d, batch, cb = _setup(cpu_pkg, "g4_tiny")
    eng.tc.pack_sequences = True
    want = eng.train_step(batch, train=False)
    store = eng._stores["s.h"]
    short = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
    for side in ("chosen", "rejected"):
        short[f"{side}_attention_mask"][:, -6:] = 0
        short[f"{side}_labels"][:, -6:] = -100
    eng.train_step(short, train=False)
    assert eng._stores["s.h"] is store and eng._bufs["s.h"].shape[0] < want_rows(eng, cb)
    again = eng.train_step(batch, train=False)
    assert eng._stores["s.h"] is store
    assert again == want


def want_rows(eng, cb):
    return sum(eng.host_seq_lens(cb["concatenated_input_ids"], cb["concatenated_attention_mask"]))


def test_merged_seq_lens_equal_merge_index_and_reject_left_padding(cpu_pkg):
    config, engine, host, ops = cpu_pkg
    eng, rcfg, d, batch, cb = _setup(cpu_pkg, "g4_tiny")
    ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    m = ops.llava_merge_index(ids, am, lb, rcfg.n_patches, ids.shape[0] // 2, 1, rcfg.image_token_index, rcfg.pad_token_id)
    assert host.merged_seq_lens(ids, am, rcfg.image_token_index, rcfg.n_patches) == m.seqlens.tolist()
    left = am.clone()
    left[0, 0] = 0
    with pytest.raises(ValueError):
        host.merged_seq_lens(ids, left, rcfg.image_token_index, rcfg.n_patches)
    with pytest.raises(ValueError):
        ops.pack_merge_rows(m, [1] * (m.n_seq + 1))
    cut = am.clone()
    cut[1, 1:] = 0   # the <image> placeholder (position 1) falls outside the attended prefix
    with pytest.raises(ValueError):
        host.merged_seq_lens(ids, cut, rcfg.image_token_index, rcfg.n_patches)


def _left_pad(batch, pad_id):
    """the same collated batch with every sequence's padding moved to the LEFT"""
    out = dict(batch)
    for side in ("chosen", "rejected"):
        ids, am, lb = (batch[f"{side}_{k}"].clone() for k in ("input_ids", "attention_mask", "labels"))
        for b in range(ids.shape[0]):
            n = int(am[b].sum())
            k = ids.shape[1] - n
            ids[b] = torch.cat([torch.full((k,), pad_id, dtype=ids.dtype), ids[b, :n]])
            lb[b] = torch.cat([torch.full((k,), -100, dtype=lb.dtype), lb[b, :n]])
            am[b] = torch.cat([torch.zeros(k, dtype=am.dtype), am[b, :n]])
        out[f"{side}_input_ids"], out[f"{side}_attention_mask"], out[f"{side}_labels"] = ids, am, lb
    return out


@pytest.mark.parametrize("pack", [False, True])
def test_left_padded_batch_gives_the_right_padded_results(cpu_pkg, pack):
    """f-2: train_step moves the attended tokens to the front (host.right_pad_valid_tokens); nothing it returns depends on
    where the padding sat (the `logits/*` means aside, which cover the padding positions' own logits)."""
    config, engine, host, ops = cpu_pkg
    res = []
    for left in (False, True):
        eng, rcfg, d, batch, cb = _setup(cpu_pkg, "g4_tiny")
        eng.tc.pack_sequences = pack
        b = _left_pad(batch, eng.tc.padding_value) if left else batch
        if left:
            assert int(b["rejected_attention_mask"][:, 0].min()) == 0   # really left-padded
        res.append((eng.train_step(b, train=True), eng.grads.clone()))
    (m0, g0), (m1, g1) = res
    for k in m0:
        if not k.startswith("logits/"):
            assert m0[k] == m1[k], k
    assert torch.equal(g0, g1)
    ids = torch.tensor([[7, 8, 0, 9], [1, 2, 3, 4]])
    am = torch.tensor([[1, 1, 0, 1], [1, 1, 1, 1]])
    lb = torch.tensor([[-100, 8, -100, 9], [-100, 2, 3, 4]])
    i2, a2, l2 = host.right_pad_valid_tokens(ids, am, lb, 0, -100)    # interleaved padding: order of the valid tokens kept
    assert i2.tolist() == [[7, 8, 9, 0], [1, 2, 3, 4]] and a2.tolist() == [[1, 1, 1, 0], [1, 1, 1, 1]]
    assert l2.tolist() == [[-100, 8, 9, -100], [-100, 2, 3, 4]]
    same = host.right_pad_valid_tokens(i2, a2, l2)
    assert same[0] is i2 and same[1] is a2 and same[2] is l2          # prefix masks pass through untouched
    with pytest.raises(ValueError):                                   # the reference's DDPO diff depends on the padding side
        host.right_pad_valid_tokens(ids, am, lb, 0, -100, loss_type="ddpo")
    assert host.right_pad_valid_tokens(i2, a2, l2, loss_type="ddpo")[0] is i2


def test_pack_sequences_launcher_switch(cpu_pkg, monkeypatch):
    """VLB200_PACK_SEQUENCES=1 reaches engines built without a TrainConfig (the way the reference's dpo.py builds the model);
    an explicit TrainConfig keeps its own setting."""
    config, engine, host, ops = cpu_pkg
    monkeypatch.setenv("VLB200_PACK_SEQUENCES", "1")
    assert engine.LlavaDPOEngine(config.TINY, None, device="cpu", with_optimizer=False).tc.pack_sequences
    assert not engine.LlavaDPOEngine(config.TINY, config.TrainConfig(), device="cpu", with_optimizer=False).tc.pack_sequences
    monkeypatch.delenv("VLB200_PACK_SEQUENCES")
    assert not engine.LlavaDPOEngine(config.TINY, None, device="cpu", with_optimizer=False).tc.pack_sequences


def _two_image_batch(rcfg, seed=3, n_pairs=2, L=24, pl=8):
    """every sequence holds TWO <image> placeholders (the second one inside the shared prompt); pixel_values pair-major"""
    b = R.make_batch(rcfg, n_pairs, L, pl, seed, ddpo_like=True)
    for side in ("chosen", "rejected"):
        b[f"{side}_input_ids"][:, 4] = rcfg.image_token_index
    g = torch.Generator().manual_seed(seed)
    b["img_input_dict"] = {"pixel_values": torch.randn(n_pairs * 2, 3, rcfg.image_size, rcfg.image_size, generator=g)}
    return b


@pytest.mark.parametrize("pack,loss_type", [(False, "sigmoid"), (True, "ddpo")])
def test_two_images_per_sequence_match_the_oracle(cpu_pkg, pack, loss_type):
    """f-2 (multi-image, uniform count): k <image> placeholders per sequence, pixel_values [B*k, ...] -- the merge places the k
    feature blocks (Llava/__init__.py:44-47,60-94), chosen and rejected share the pair's images; against the oracle's
    restatement of the reference merge (itself checked on the live reference, tests/test_oracle_vs_reference.py)."""
    config, engine, host, ops = cpu_pkg
    batch = _two_image_batch(R.TINY)
    eng = engine.LlavaDPOEngine(config.TINY, config.TrainConfig(loss_type=loss_type, learning_rate=1e-3, pack_sequences=pack),
                                device="cpu", with_optimizer=False)
    eng.init_synthetic(0)
    assert eng.images_per_sequence(batch) == 2
    got = eng.train_step(batch, train=True)
    assert eng._saved["m"].S == 24 + 2 * (R.TINY.n_patches - 1) and eng._saved["m"].imgs_per_seq == 2
    wp, wr = R.make_policy_and_ref(R.TINY, 0)
    with torch.no_grad():
        loss, metrics, _ = R.get_batch_loss_metrics(R.TINY, wp, wr, batch, loss_type=loss_type)
    assert abs(got["loss"] - float(loss)) < 2e-3
    for k in ("rewards/chosen", "rewards/rejected", "logps/chosen", "logps/rejected"):
        assert abs(got[k] - float(metrics[k])) < 2e-3 * max(1.0, abs(float(metrics[k]))), k
    grads = {k: v.float() for k, v in eng.hf_state("grad").items()}
    leaves = {k: v.clone().requires_grad_(True) for k, v in wp.items() if not k.startswith("vision_tower.")}
    w = dict(wp)
    w.update(leaves)
    loss, _, _ = R.get_batch_loss_metrics(R.TINY, w, wr, batch, loss_type=loss_type)
    loss.backward()
    for k, leaf in leaves.items():
        rel = (grads[k].reshape(leaf.shape) - leaf.grad).norm().item() / max(leaf.grad.norm().item(), 1e-12)
        assert rel < 5e-2, f"{k}: rel {rel}"
    bad = dict(batch, img_input_dict={"pixel_values": batch["img_input_dict"]["pixel_values"][:3]})
    with pytest.raises(ValueError):
        eng.train_step(bad, train=False)            # 3 images do not split over 2 pairs
