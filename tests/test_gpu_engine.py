"""GPU parity of the whole DPO step (engine.py through the C ABI) against the oracle and the golden vectors
minted from the reference's LlavaForRL.forward / get_batch_logps / dpo_loss.

Tolerances: north_star asks <= 1e-3 relative on per-pair log-probs and loss vs the fp32 oracle (bf16 compute).
"""
import os

import numpy as np
import pytest
import torch

from oracle import restate as R
from tests import parity_log

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def pkg():
    import vlrlhf_b200  # noqa: F401
    from vlrlhf_b200 import config, engine, host, ops
    return config, engine, host, ops


CASES = {"g4_tiny": ("TINY", R.TINY, 2, 24, 8), "g4_small": ("SMALL", R.SMALL, 2, 96, 24),
         # LLaVA-Next (models/LlavaNext): anyres crops with unpadding, image_newline, GQA 4/2, rope theta 1e6
         "g6_next_tiny": ("TINY_NEXT", R.TINY_NEXT, 3, 24, 8), "g6_next_small": ("SMALL_NEXT", R.SMALL_NEXT, 2, 96, 24)}
ALL_TAGS = list(CASES)


def build(pkg, tag, loss_type="sigmoid", with_optimizer=True):
    config, engine, host, ops = pkg
    name, rcfg, npairs, tl, pl = CASES[tag]
    d = np.load(os.path.join(G, tag + ".npz"))
    seed = int(d["seed"])
    eng = engine.LlavaDPOEngine(getattr(config, name), config.TrainConfig(loss_type=loss_type, learning_rate=1e-3),
                                with_optimizer=with_optimizer)
    eng.init_synthetic(seed)
    sizes = [tuple(x) for x in d["image_sizes"].tolist()] if "image_sizes" in d.files else None
    batch = R.make_batch(rcfg, npairs, tl, pl, seed, ddpo_like=True, image_sizes=sizes)
    cb = host.concatenated_inputs(batch)
    return eng, rcfg, d, batch, cb


def stage(eng, host, cb, rcfg, ddpo=False):
    """-> the argument tuple of eng.step (5 entries, 6 with the anyres plan for LLaVA-Next)."""
    ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    img = cb["concatenated_img_input_dict"]
    wt = eng.ddpo_weights(ids, am, lb, img.get("image_sizes")) if ddpo else None
    return eng.prepare_inputs(ids, am, lb, img["pixel_values"], wt, img.get("image_sizes"))


@pytest.mark.parametrize("tag", ALL_TAGS)
def test_synthetic_weights_bit_exact(pkg, tag):
    eng, rcfg, d, batch, cb = build(pkg, tag, with_optimizer=False)
    wp, wr = R.make_policy_and_ref(rcfg, int(d["seed"]))
    pol, ref = eng.hf_state("policy"), eng.hf_state("ref")
    n_checked = 0
    for k, v in wp.items():
        if k not in pol:
            assert "encoder.layers" in k  # vision layers above vision_feature_layer are not materialised
            continue
        assert torch.equal(pol[k].float().cpu().reshape(v.shape), v), k
        if not k.startswith("vision_tower."):
            assert torch.equal(ref[k].float().cpu().reshape(v.shape), wr[k]), k
        n_checked += 1
    assert n_checked > 20


@pytest.mark.parametrize("tag", ALL_TAGS)
def test_forward_logps_and_loss_parity(pkg, tag):
    config, engine, host, ops = pkg
    eng, rcfg, d, batch, cb = build(pkg, tag, with_optimizer=False)
    a = stage(eng, host, cb, rcfg)
    out = eng.step(*a, train=False)
    if len(a) == 6:  # LLaVA-Next: the merged labels are bit-exact with the reference's (integer work)
        m = ops.llavanext_merge_index(a[0], a[1], a[2], a[5].feat_off, a[5].total_feats, a[5].merged_len, len(a[5].crops),
                                      1, rcfg.image_token_index)
        assert int(m.status) == 0
        assert np.array_equal(m.labels.cpu().numpy(), d["labels"])
        src = m.src_map.view(m.n_seq, m.S)
        assert np.array_equal(((src < 0) & (src != -(2 ** 31))).cpu().numpy(), d["image_position_map"])
    # golden = reference LlavaForRL.forward + get_batch_logps + dpo_loss (fp32 CPU); log-probs 1e-3 relative, losses and
    # rewards to the absolute error measured on a B200 x2 (tests/parity_log.py, profiles/parity_r2.md)
    pol, ref = parity_log.check_step(tag, out, d)
    slack = 0.1 * 1e-3 * np.abs(d["policy_logps"]).max() * 4
    # same comparison against the oracle restatement run here (CPU fp32)
    wp, wr = R.make_policy_and_ref(rcfg, int(d["seed"]))
    with torch.no_grad():
        loss, metrics, aux = R.get_batch_loss_metrics(rcfg, wp, wr, batch)
    want = torch.cat([aux["policy_chosen_logps"], aux["policy_rejected_logps"]]).numpy()
    np.testing.assert_allclose(pol, want, rtol=1e-3)
    m = eng._saved["m"] if hasattr(eng, "_saved") else None
    # other loss types on the same log-probs
    for lt in ("ipo", "hinge", "kto_pair"):
        losses, cr, rr, stats, _ = ops.dpo_loss(out.policy_logps, out.ref_logps, 0.1, 0.0, lt, want_grad=False)
        np.testing.assert_allclose(losses.cpu().numpy(), d[f"{lt}_losses"], atol=max(slack * (20 if lt == "ipo" else 1), 1e-4),
                                   rtol=2e-2 if lt == "ipo" else 0)


@pytest.mark.parametrize("tag", ALL_TAGS)
def test_ddpo_parity(pkg, tag):
    config, engine, host, ops = pkg
    eng, rcfg, d, batch, cb = build(pkg, tag, loss_type="ddpo", with_optimizer=False)
    a = stage(eng, host, cb, rcfg, ddpo=True)
    assert int(a[4].sum()) > 0
    out = eng.step(*a, train=False)
    np.testing.assert_allclose(out.policy_logps.cpu().numpy(), d["policy_logps_ddpo"], rtol=1e-3, atol=1e-2)
    np.testing.assert_allclose(out.ref_logps.cpu().numpy(), d["ref_logps_ddpo"], rtol=1e-3, atol=1e-2)
    np.testing.assert_allclose(out.losses.cpu().numpy(), d["ddpo_losses"], atol=5e-3)


@pytest.mark.parametrize("tag", ALL_TAGS)
def test_backward_matches_oracle_autograd(pkg, tag):
    config, engine, host, ops = pkg
    eng, rcfg, d, batch, cb = build(pkg, tag, with_optimizer=False)
    eng.step(*stage(eng, host, cb, rcfg), train=True)
    torch.cuda.synchronize()
    got = {k: v.float().cpu() for k, v in eng.hf_state("grad").items()}
    wp, wr = R.make_policy_and_ref(rcfg, int(d["seed"]))
    leaves = {k: v.clone().requires_grad_(True) for k, v in wp.items() if not k.startswith("vision_tower.")}
    w = dict(wp)
    w.update(leaves)
    loss, _, _ = R.get_batch_loss_metrics(rcfg, w, wr, batch)
    loss.backward()
    worst = 0.0
    for k, leaf in leaves.items():
        want = leaf.grad
        g = got[k].reshape(want.shape)
        assert torch.isfinite(g).all(), k
        denom = want.norm().item()
        rel = (g - want).norm().item() / max(denom, 1e-12)
        worst = max(worst, rel)
        # bf16 activations + bf16 gradient storage: a few percent of relative L2 noise per tensor
        assert rel < 6e-2, f"{k}: rel l2 err {rel:.4g} (|want|={denom:.3g})"
        cos = torch.nn.functional.cosine_similarity(g.flatten(), want.flatten(), dim=0).item()
        assert cos > 0.998, f"{k}: cosine {cos}"
    print(f"[{tag}] worst per-tensor gradient rel-l2 error {worst:.4g}")


def test_optimizer_step_reduces_loss_and_tracks_master(pkg):
    config, engine, host, ops = pkg
    eng, rcfg, d, batch, cb = build(pkg, "g4_tiny")
    ids, am, lb, px, _ = stage(eng, host, cb, rcfg)
    losses = []
    for _ in range(4):
        out = eng.step(ids, am, lb, px, train=True)
        losses.append(float(out.stats[0]))
    assert losses[-1] < losses[0], losses
    eng.wait_optimizer()  # the optimizer runs on a side stream; order this stream after it before peeking at its state
    assert torch.equal(eng.params, eng.master.to(torch.bfloat16))
    assert torch.isfinite(eng.master).all()
    assert eng.opt_step == 4 and float(eng.grad_sumsq) > 0
    # the reference copy and the vision tower are never touched by the step
    wp, wr = R.make_policy_and_ref(rcfg, int(d["seed"]))
    ref = eng.hf_state("ref")
    k = "language_model.model.layers.0.mlp.down_proj.weight"
    assert torch.equal(ref[k].float().cpu(), wr[k])


@pytest.mark.parametrize("tag", ["g4_small", "g6_next_small"])
def test_train_step_metric_keys_and_values(pkg, tag):
    """The public end-to-end call (host batch in, TRL metric dict out) incl. the logits/* means computed without logits."""
    config, engine, host, ops = pkg
    eng, rcfg, d, batch, cb = build(pkg, tag)
    got = eng.train_step(batch, train=True)
    wp, wr = R.make_policy_and_ref(rcfg, int(d["seed"]))
    with torch.no_grad():
        loss, metrics, aux = R.get_batch_loss_metrics(rcfg, wp, wr, batch)
    assert set(got) >= {"loss", "rewards/chosen", "rewards/rejected", "rewards/accuracies", "rewards/margins",
                        "logps/chosen", "logps/rejected", "logits/chosen", "logits/rejected"}
    assert abs(got["loss"] - float(loss)) < 5e-2
    for k in ("logps/chosen", "logps/rejected"):
        assert abs(got[k] / float(metrics[k]) - 1) < 1e-3, k
    for k in ("logits/chosen", "logits/rejected"):
        assert abs(got[k] - float(metrics[k])) < 5e-3, (k, got[k], float(metrics[k]))
    assert got["grad_norm"] > 0


def test_config1_7b_shapes_logps_and_loss_parity(pkg):
    """BASELINE.json configs[0]: LLaVA-1.5-7B shapes, 2 pairs, text 128 (703 merged), loss/logprob parity against the
    reference's LlavaForRL.forward + get_batch_logps + dpo_loss run in fp32 on CPU (tests/golden/g5_config1_7b.npz).
    7B-shape weights are regenerated on the GPU by the bit-exact hash twin, so nothing travels."""
    config, engine, host, ops = pkg
    d = np.load(os.path.join(G, "g5_config1_7b.npz"))
    rcfg = R.LLAVA15_7B
    eng = engine.LlavaDPOEngine(config.LLAVA15_7B, config.TrainConfig(), with_optimizer=False)
    eng.init_synthetic(int(d["seed"]))
    batch = R.make_batch(rcfg, int(d["n_pairs"]), int(d["text_len"]), int(d["prompt_len"]), int(d["seed"]))
    cb = host.concatenated_inputs(batch)
    ids, am, lb, px, _ = stage(eng, host, cb, rcfg)
    out = eng.step(ids, am, lb, px, train=False)
    pol, ref = out.policy_logps.cpu().numpy(), out.ref_logps.cpu().numpy()
    if "policy_logps_refdtype_bf16" in d.files:   # the noise floor of the reference's own bf16 deployment, for the table
        parity_log.record("g5_config1_7b", "REFERENCE bf16 vs its fp32 (policy_logps)", d["policy_logps_refdtype_bf16"],
                          d["policy_logps"], note="the reference's own bf16 path, not ours")
    parity_log.check_step("g5_config1_7b", out, d, rtol=parity_log.RTOL_7B)
    if "policy_logps_refdtype_bf16" in d.files:   # closer to the fp32 run than the reference's own bf16 execution
        ours = np.abs(pol / d["policy_logps"] - 1).max()
        theirs = np.abs(d["policy_logps_refdtype_bf16"] / d["policy_logps"] - 1).max()
        assert ours < 0.5 * theirs, (ours, theirs)
    del eng
    torch.cuda.empty_cache()


def test_merge_validity_errors(pkg):
    config, engine, host, ops = pkg
    eng, rcfg, d, batch, cb = build(pkg, "g4_tiny", with_optimizer=False)
    ids = cb["concatenated_input_ids"].clone()
    ids[0, 3] = rcfg.image_token_index  # a second image token in one sequence only
    with pytest.raises(ValueError, match="number of image tokens"):   # host batches are refused before any device work
        eng.prepare_inputs(ids, cb["concatenated_attention_mask"], cb["concatenated_labels"],
                           cb["concatenated_img_input_dict"]["pixel_values"])
    # device-resident batches skip the host check: the merge kernel's status word carries the same verdict
    a, b, c, px, _ = eng.prepare_inputs(ids.cuda(), cb["concatenated_attention_mask"].cuda(), cb["concatenated_labels"].cuda(),
                                        cb["concatenated_img_input_dict"]["pixel_values"].cuda())
    m = ops.llava_merge_index(a, b, c, rcfg.n_patches, px.shape[0], 1, rcfg.image_token_index, rcfg.pad_token_id)
    with pytest.raises(ValueError):
        eng.check_merge_status(m)


def test_next_merge_kernels_match_host_mock(pkg):
    """Integer kernels of the LLaVA-Next merge == the plain-Python mirror (tests/mock_ops.py), bit-exact; the backward
    gather-sum over the two sequences that share an image == index_add."""
    config, engine, host, ops = pkg
    from tests import mock_ops
    cfg = R.SMALL_NEXT
    sizes = [(112, 112), (90, 300), (200, 100), (50, 50)]
    batch = R.make_batch(cfg, 4, 50, 8, seed=9, image_sizes=sizes)
    cb = host.concatenated_inputs(batch)
    ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    plan = host.anyres_pack_index(sizes, cfg.image_grid_pinpoints, cfg.image_size, cfg.patch_size)
    S = host.next_merged_len(ids, am, plan.feature_lens, cfg.image_token_index)
    want = mock_ops.llavanext_merge_index(ids, am, lb, plan.feat_off, plan.total_feats, S, 4, 1, cfg.image_token_index)
    got = ops.llavanext_merge_index(ids.cuda(), am.cuda(), lb.cuda(), plan.feat_off.cuda(), plan.total_feats, S, 4, 1,
                                    cfg.image_token_index)
    assert int(got.status) == 0
    for k in ("src_map", "labels", "mask", "pos", "seqlens", "img_pos", "target"):
        assert torch.equal(getattr(got, k).cpu(), getattr(want, k)), k
    live = want.target >= 0
    assert torch.equal(got.row_of_text.cpu()[live], want.row_of_text[live])
    d = 64
    g = torch.Generator().manual_seed(0)
    dx = torch.randn(8 * S, d, generator=g).to(torch.bfloat16)
    de_w = torch.zeros(cfg.vocab, d)
    di_w = torch.zeros(plan.total_feats, d, dtype=torch.bfloat16)
    mock_ops.llavanext_merge_bwd(want, dx, de_w, di_w)
    de_g = torch.zeros(cfg.vocab, d, device="cuda")
    di_g = torch.zeros(plan.total_feats, d, dtype=torch.bfloat16, device="cuda")
    ops.llavanext_merge_bwd(got, dx.cuda(), de_g, di_g)
    assert torch.equal(di_g.cpu(), di_w)
    torch.testing.assert_close(de_g.cpu(), de_w, rtol=1e-5, atol=1e-5)
    # the reference's ValueErrors (LlavaNext/__init__.py:75-79, 60-71) surface through the status word
    bad = ids.clone(); bad[0, 3] = cfg.image_token_index
    st = ops.llavanext_merge_index(bad.cuda(), am.cuda(), lb.cuda(), plan.feat_off.cuda(), plan.total_feats, S + 400, 4, 1,
                                   cfg.image_token_index)
    eng = engine.LlavaDPOEngine(config.TINY_NEXT, config.TrainConfig(), with_optimizer=False)
    with pytest.raises(ValueError):
        eng.check_merge_status(st)


@pytest.mark.parametrize("tag", ["g4_small", "g6_next_small"])
def test_activation_checkpointing_bit_identical(pkg, tag):
    """TrainConfig.activation_checkpointing keeps only the fp32 layer inputs and recomputes each layer in backward:
    every kernel is deterministic, so gradients and log-probs are bit-identical to the keep-everything schedule."""
    config, engine, host, ops = pkg
    res = []
    for ckpt in (False, True):
        eng, rcfg, d, batch, cb = build(pkg, tag, with_optimizer=False)
        eng.tc.activation_checkpointing = ckpt
        out = eng.step(*stage(eng, host, cb, rcfg), train=True)
        torch.cuda.synchronize()
        res.append((eng.grads.clone(), out.policy_logps.clone()))
        assert (not any(k.startswith("a.qkv") for k in eng._bufs)) == ckpt
    assert torch.equal(res[0][1], res[1][1])
    assert torch.equal(res[0][0], res[1][0])


def test_config4_next7b_shapes_ddpo_parity(pkg):
    """BASELINE.json configs[3] at parity size: LLaVA-Next-Mistral-7B shapes (CLIP-L/336 anyres crops, Mistral decoder
    GQA 32/8, ff 14336, theta 1e6), 1 pair, text 96, DDPO token weights -- against the reference's LlavaNextForRL.forward
    + get_batch_logps(mask_shared_tokens) + dpo_loss run in fp32 on CPU (tests/golden/g8_config4_next7b.npz)."""
    config, engine, host, ops = pkg
    path = os.path.join(G, "g8_config4_next7b.npz")
    if not os.path.exists(path):
        pytest.skip("g8 fixture not generated yet (oracle/make_fixtures.py --config4)")
    d = np.load(path)
    rcfg = R.LLAVANEXT_MISTRAL_7B
    sizes = [tuple(x) for x in d["image_sizes"].tolist()]
    eng = engine.LlavaDPOEngine(config.LLAVANEXT_MISTRAL_7B, config.TrainConfig(loss_type="ddpo"), with_optimizer=False)
    eng.init_synthetic(int(d["seed"]))
    batch = R.make_batch(rcfg, int(d["n_pairs"]), int(d["text_len"]), int(d["prompt_len"]), int(d["seed"]), ddpo_like=True,
                         image_sizes=sizes)
    cb = host.concatenated_inputs(batch)
    for ddpo, key in ((False, "policy_logps"), (True, "policy_logps_ddpo")):
        a = stage(eng, host, cb, rcfg, ddpo=ddpo)
        out = eng.step(*a, train=False)
        pol, ref = out.policy_logps.cpu().numpy(), out.ref_logps.cpu().numpy()
        rkey = key.replace("policy", "ref")
        print("config4", key, pol, "golden", d[key], "rel", np.abs(pol / d[key] - 1), "ref rel", np.abs(ref / d[rkey] - 1))
        if not ddpo:
            np.testing.assert_allclose(pol, d[key], rtol=parity_log.RTOL_7B)
            np.testing.assert_allclose(ref, d[rkey], rtol=parity_log.RTOL_7B)
        else:
            # DDPO log-probs are a 0/1-weighted SUBSET of the same per-token terms (a few dozen of the ~70 labelled
            # tokens here): their absolute error is bounded by the budget the 1e-3 relative bound gives the full sum
            assert (np.abs(pol - d[key]) <= 1e-3 * np.abs(d["policy_logps"])).all()
            assert (np.abs(ref - d[rkey]) <= 1e-3 * np.abs(d["ref_logps"])).all()
            assert np.abs(pol / d[key] - 1).max() < 5e-3
    b = parity_log.LOSS_ABS_BOUNDS.get("g8_config4_next7b", 0.1 * 1e-3 * np.abs(d["policy_logps_ddpo"]).max() * 4)
    parity_log.record("g8_config4_next7b", "ddpo_losses", out.losses.cpu().numpy(), d["ddpo_losses"], bound_abs=b)
    del eng
    torch.cuda.empty_cache()


def test_config2_full_length_7b_parity(pkg):
    """BASELINE.json configs[1] -- the headline bench shape -- at its FULL sequence length: LLaVA-1.5-7B shapes, ONE pair,
    text 1024 -> 1599 merged rows, against the reference's LlavaForRL.forward + get_batch_logps + dpo_loss run in fp32 on the
    CPU (tests/golden/g14_config2_full_7b.npz, oracle/make_fixtures.py --config2).  ~900 labelled tokens per sequence."""
    config, engine, host, ops = pkg
    path = os.path.join(G, "g14_config2_full_7b.npz")
    if not os.path.exists(path):
        pytest.skip("g14 fixture not generated yet (oracle/make_fixtures.py --config2)")
    d = np.load(path)
    rcfg = R.LLAVA15_7B
    eng = engine.LlavaDPOEngine(config.LLAVA15_7B, config.TrainConfig(), with_optimizer=False)
    eng.init_synthetic(int(d["seed"]))
    batch = R.make_batch(rcfg, int(d["n_pairs"]), int(d["text_len"]), int(d["prompt_len"]), int(d["seed"]))
    cb = host.concatenated_inputs(batch)
    ids, am, lb, px, _ = stage(eng, host, cb, rcfg)
    assert ids.shape[1] == 1024
    out = eng.step(ids, am, lb, px, train=False)
    assert eng._bufs["s.x0"].shape[0] == 2 * 1599
    parity_log.check_step("g14_config2_full_7b", out, d, rtol=parity_log.RTOL_7B)
    # the same batch as packed rows (f-2): identical log-probs to the padded layout, still within the bound
    eng.tc.pack_sequences = True
    seq_lens = eng.host_seq_lens(cb["concatenated_input_ids"], cb["concatenated_attention_mask"])
    outp = eng.step(ids, am, lb, px, train=False, seq_lens=seq_lens)
    parity_log.record("g14_config2_full_7b", "packed policy_logps", outp.policy_logps.cpu().numpy(), d["policy_logps"], bound_rel=1e-3)
    del eng
    torch.cuda.empty_cache()


@pytest.mark.parametrize("name", ["SMALL", "SMALL_NEXT"])
def test_hf_checkpoint_roundtrip_on_gpu(pkg, tmp_path, name):
    """save_pretrained -> from_pretrained through the plugin model: same parameters, same log-probs (f-4)."""
    config, engine, host, ops = pkg
    from vlrlhf_b200 import checkpoint, plugin
    cfg = getattr(config, name)
    rcfg = getattr(R, name)
    a = plugin.B200LlavaForRL(cfg, config.TrainConfig(), with_optimizer=False)
    a.engine.init_synthetic(3)
    a.hf_config_dict = checkpoint.hf_config_dict(cfg)
    assert checkpoint.config_from_hf(a.hf_config_dict) == cfg
    a.save_pretrained(str(tmp_path), max_shard_size=4 << 20)
    b = plugin.B200LlavaForRL.from_pretrained(str(tmp_path), torch_dtype=torch.bfloat16, with_optimizer=False)
    assert b.cfg == cfg
    assert torch.equal(a.engine.params, b.engine.params) and torch.equal(a.engine.vparams, b.engine.vparams)
    assert torch.equal(b.engine.ref_params, b.engine.params[: b.engine.ref_params.numel()])  # reference copy = initial policy
    sizes = [(112, 112), (90, 300)] if cfg.family == "llava_next" else None
    batch = R.make_batch(rcfg, 2, 48, 12, seed=5, image_sizes=sizes)
    cb = host.concatenated_inputs(batch)
    outs = [m.engine.step(*stage(m.engine, host, cb, rcfg), train=False).policy_logps for m in (a, b)]
    assert torch.equal(outs[0], outs[1])


# ------------------------------------------------------------------ packed rows
@pytest.mark.parametrize("tag,ddpo", [("g4_small", False), ("g4_small", True), ("g6_next_small", False)])
def test_packed_step_equals_padded_step(pkg, tag, ddpo):
    """TrainConfig.pack_sequences drops the padding rows of the ragged fixture batch: the log-probs are those of the padded
    step (every surviving row goes through the same arithmetic), the fixtures still hold to 1e-3, the gradients agree to
    accumulation order (the weight gradients contract over the rows), and nothing non-finite leaks in from unwritten rows."""
    config, engine, host, ops = pkg
    res = []
    for pack in (False, True):
        eng, rcfg, d, batch, cb = build(pkg, tag, loss_type="ddpo" if ddpo else "sigmoid", with_optimizer=False)
        eng.tc.pack_sequences = pack
        args = stage(eng, host, cb, rcfg, ddpo=ddpo)
        lens = eng.host_seq_lens(cb["concatenated_input_ids"], cb["concatenated_attention_mask"],
                                 cb["concatenated_img_input_dict"].get("image_sizes")) if pack else None
        eng.step(*args, train=True, seq_lens=lens)   # allocates the workspaces
        if pack:   # poison them, storage beyond the packed rows included: what a packed step does not write it must not read
            for t in eng._stores.values():
                if t.is_floating_point():
                    t.fill_(float("nan"))
        out = eng.step(*args, train=True, seq_lens=lens)
        m = eng._saved["m"]
        assert m.packed == pack
        if pack:
            assert m.T == sum(lens) < m.n_seq * m.S
        torch.cuda.synchronize()
        res.append((out.policy_logps.clone(), out.ref_logps.clone(), out.losses.clone(), eng.grads.clone().float()))
    (p0, r0, l0, g0), (p1, r1, l1, g1) = res
    for a, b in ((p0, p1), (r0, r1), (l0, l1)):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-4)
    key = "policy_logps_ddpo" if ddpo else "policy_logps"
    np.testing.assert_allclose(p1.cpu().numpy(), d[key], rtol=1e-3, atol=1e-2 if ddpo else 0)
    assert torch.isfinite(g1).all()
    # weight gradients contract over the rows: other rows share a 64-row k-block / a 16-row MMA step once the padding is gone,
    # so their fp32 partial sums round differently and a few per cent of the bf16 results land one ulp away
    rel = ((g0 - g1).norm() / g0.norm()).item()
    cos = (torch.dot(g0, g1) / (g0.norm() * g1.norm())).item()
    assert rel < 1e-2 and cos > 0.9999, f"gradient rel l2 {rel}, cosine {cos}"


def test_packed_step_without_host_lengths_reads_them_back(pkg):
    """step() without seq_lens falls back to reading the merge kernel's seqlens from the device (one synchronisation)."""
    config, engine, host, ops = pkg
    eng, rcfg, d, batch, cb = build(pkg, "g4_tiny", with_optimizer=False)
    want = eng.step(*stage(eng, host, cb, rcfg), train=False).policy_logps.clone()
    eng.tc.pack_sequences = True
    got = eng.step(*stage(eng, host, cb, rcfg), train=False).policy_logps
    assert eng._stores
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-4)
