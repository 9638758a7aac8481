"""GPU tests written after the round's GPU budget was spent: they have run against the CPU mirror of the ops (tests/mock_ops.py)
only.  The file sorts last among the GPU tests on purpose, so that under `pytest -x` a failure here cannot hide the results of
the validated parity tests.

* BASELINE.json configs[1] at its full size (7B shapes, 4 pairs, text 1024) through size-independent properties;
* packed rows (TrainConfig.pack_sequences) on the Qwen-VL and XC2 engines -- the kernels involved (vlb200_pack_merge_rows,
  the var-len attention entry points, gather / scatter-add rows) are the ones the LLaVA packed tests already exercise on a GPU;
* two <image> placeholders per sequence (the multi-slot path of vlb200_llava_merge_index / _merge_bwd).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import vlrlhf_b200  # noqa: F401
    from vlrlhf_b200 import config, engine, host, ops
    return config, engine, host, ops


def test_config2_full_size_properties(pkg):
    """BASELINE.json configs[1] at its FULL size (LLaVA-1.5-7B shapes, 4 pairs, text 1024 -> 1599 merged rows per sequence,
    T = 12 792 rows): no CPU oracle finishes that inside a test, so the forward half of the step is checked through
    properties that hold at any size:
      (i)   reference == policy  =>  equal log-probs, every loss == ln 2, rewards and margins == 0;
      (ii)  sequences are independent units: swapping chosen and rejected swaps the log-probs;
      (iii) dropping the padding rows (TrainConfig.pack_sequences) changes no log-prob;
      (iv)  log-probs are finite sums of log-probabilities (< 0) over exactly the labelled tokens."""
    from vlrlhf_b200 import synthetic
    config, engine, host, ops = pkg
    cfg = config.LLAVA15_7B
    eng = engine.LlavaDPOEngine(cfg, config.TrainConfig(), with_optimizer=False)
    eng.init_synthetic(0, ref_alpha=0.0)
    assert torch.equal(eng.ref_params, eng.params[: eng.ref_params.numel()])
    batch = synthetic.make_batch(cfg, 4, 1024, 128, seed=1000)
    cb = host.concatenated_inputs(batch)
    ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    px = batch["img_input_dict"]["pixel_values"]
    out = eng.step(*eng.prepare_inputs(ids, am, lb, px), train=False)
    pol, ref = out.policy_logps.clone(), out.ref_logps.clone()
    assert torch.equal(pol, ref)                                                         # (i)
    np.testing.assert_allclose(out.losses.cpu().numpy(), np.log(2.0), rtol=0, atol=1e-6)
    assert float(out.chosen_rewards.abs().max()) == 0.0 and float(out.rejected_rewards.abs().max()) == 0.0
    n_lab = (lb != -100).sum(-1).float().cuda()
    assert torch.isfinite(pol).all() and bool((pol < 0).all())                           # (iv)
    per_tok = (-pol / n_lab).cpu().numpy()
    assert (per_tok > 1.0).all() and (per_tok < 40.0).all()   # random weights: around ln V = 10.4 nats per labelled token
    swapped = {k: v for k, v in batch.items()}                                           # (ii)
    for k in ("input_ids", "attention_mask", "labels"):
        swapped[f"chosen_{k}"], swapped[f"rejected_{k}"] = batch[f"rejected_{k}"], batch[f"chosen_{k}"]
    cs = host.concatenated_inputs(swapped)
    out_s = eng.step(*eng.prepare_inputs(cs["concatenated_input_ids"], cs["concatenated_attention_mask"],
                                         cs["concatenated_labels"], px), train=False)
    torch.testing.assert_close(out_s.policy_logps, torch.cat([pol[4:], pol[:4]]), rtol=1e-6, atol=1e-3)
    eng.tc.pack_sequences = True                                                         # (iii)
    lens = eng.host_seq_lens(ids, am)
    assert sum(lens) < 8 * 1599 and max(lens) == 1599
    out_p = eng.step(*eng.prepare_inputs(ids, am, lb, px), train=False, seq_lens=lens)
    torch.testing.assert_close(out_p.policy_logps, pol, rtol=1e-6, atol=1e-3)
    del eng
    torch.cuda.empty_cache()


def _packed_vs_padded(build_engine, batch, loss_type):
    res = []
    for pack in (False, True):
        eng = build_engine(pack)
        eng.train_step(batch, train=True)             # allocates the workspaces
        if pack:   # what a packed step does not write it must not read: poison the decoder's workspaces (saved activations,
            # scratch, backward buffers, residual stream), storage beyond the packed rows included; the tower's buffers are
            # left alone (Qwen's resampler keeps a constant query table in one of them)
            for name, t in eng._stores.items():
                if name.split(".")[0] in ("a", "s", "b", "x") and t.is_floating_point():
                    t.fill_(float("nan"))
        metrics = eng.train_step(batch, train=True)
        torch.cuda.synchronize()
        m = eng._saved["m"]
        assert m.packed == pack and (not pack or m.T < m.n_seq * m.S)
        res.append((metrics, eng.grads.clone().float()))
        del eng
    (m0, g0), (m1, g1) = res
    for k in m0:
        if not k.startswith("logits/") and k != "grad_norm":
            assert abs(m0[k] - m1[k]) <= 1e-5 * max(1.0, abs(m0[k])), (k, m0[k], m1[k])
    assert torch.isfinite(g1).all()
    rel = ((g0 - g1).norm() / g0.norm()).item()
    cos = (torch.dot(g0, g1) / (g0.norm() * g1.norm())).item()
    assert rel < 1e-2 and cos > 0.9999, f"gradient rel l2 {rel}, cosine {cos}"


@pytest.mark.parametrize("loss_type", ["sigmoid", "ddpo"])
def test_qwen_packed_step_equals_padded_step(loss_type):
    import vlrlhf_b200  # noqa: F401
    from oracle import qwen_restate as Q
    from vlrlhf_b200 import config, engine_qwen
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "g9_qwen_small.npz"))
    batch = Q.make_batch(Q.SMALL_QWEN, int(d["n_pairs"]), int(d["text_len"]), int(d["prompt_len"]), int(d["seed"]), ddpo_like=True)

    def build(pack):
        eng = engine_qwen.QwenVLDPOEngine(config.SMALL_QWEN, config.TrainConfig(loss_type=loss_type, learning_rate=1e-3,
                                                                                  pack_sequences=pack), with_optimizer=False)
        eng.init_synthetic(int(d["seed"]))
        return eng

    _packed_vs_padded(build, batch, loss_type)


@pytest.mark.parametrize("loss_type", ["kto_pair", "ddpo"])
def test_xc2_packed_step_equals_padded_step(loss_type):
    import vlrlhf_b200  # noqa: F401
    from oracle import restate as R
    from oracle import xc2_restate as X
    from vlrlhf_b200 import config, engine_xc2
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "g10_xc2_small.npz"))
    batch = R.make_batch(X.SMALL_XC2, int(d["n_pairs"]), int(d["text_len"]), int(d["prompt_len"]), int(d["seed"]), ddpo_like=True)

    def build(pack):
        eng = engine_xc2.XC2DPOEngine(config.SMALL_XC2, config.TrainConfig(loss_type=loss_type, learning_rate=1e-3,
                                                                            pack_sequences=pack), with_optimizer=False)
        eng.init_synthetic(int(d["seed"]))
        return eng

    _packed_vs_padded(build, batch, loss_type)


def test_two_images_per_sequence_on_the_gpu(pkg):
    """f-2 (multi-image, uniform count): the merge-index kernel's multi-slot path against its plain-Python mirror, and the
    engine's step on a two-image batch against the oracle."""
    from oracle import restate as R
    from tests import mock_ops
    config, engine, host, ops = pkg
    rcfg = R.TINY
    batch = R.make_batch(rcfg, 2, 24, 8, 3, ddpo_like=True)
    for side in ("chosen", "rejected"):
        batch[f"{side}_input_ids"][:, 4] = rcfg.image_token_index
    batch["img_input_dict"] = {"pixel_values": torch.randn(4, 3, rcfg.image_size, rcfg.image_size,
                                                          generator=torch.Generator().manual_seed(3))}
    cb = host.concatenated_inputs(batch)
    ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    args = (rcfg.n_patches, 2, 2, rcfg.image_token_index, rcfg.pad_token_id)
    m = ops.llava_merge_index(ids.cuda(), am.cuda(), lb.cuda(), *args)
    want = mock_ops.llava_merge_index(ids, am, lb, *args)
    assert int(m.status.item()) == 0 and m.S == want.S == 24 + 2 * (rcfg.n_patches - 1)
    for k in ("src_map", "pos", "seqlens", "img_pos", "row_of_text", "target", "labels", "mask"):
        assert torch.equal(getattr(m, k).cpu().reshape(-1).long(), getattr(want, k).reshape(-1).long()), k
    for pack in (False, True):
        eng = engine.LlavaDPOEngine(config.TINY, config.TrainConfig(learning_rate=1e-3, pack_sequences=pack), with_optimizer=False)
        eng.init_synthetic(0)
        got = eng.train_step(batch, train=True)
        wp, wr = R.make_policy_and_ref(rcfg, 0)
        with torch.no_grad():
            loss, metrics, _ = R.get_batch_loss_metrics(rcfg, wp, wr, batch)
        assert abs(got["loss"] - float(loss)) < 2e-3
        for k in ("rewards/chosen", "rewards/rejected", "logps/chosen", "logps/rejected"):
            assert abs(got[k] - float(metrics[k])) < 2e-3 * max(1.0, abs(float(metrics[k]))), k
        assert torch.isfinite(eng.grads.float()).all()


def _shared_vs_padded(build_engine, batch, loss_type, grad_rel=2e-2):
    """TrainConfig.share_prefix on an adapter engine: the padded step's log-probs / losses (<= 2e-4 relative: another row
    grouping per attention tile), adapter gradients equal up to the accumulation order."""
    from tests import parity_log
    res = []
    for share in (False, True):
        eng = build_engine(share)
        eng.train_step(batch, train=True)             # allocates the workspaces
        if share:   # what a shared step does not write it must not read (see _packed_vs_padded)
            for name, t in eng._stores.items():
                if name.split(".")[0] in ("a", "s", "b", "x") and t.is_floating_point():
                    t.fill_(float("nan"))
        metrics = eng.train_step(batch, train=True)
        torch.cuda.synchronize()
        m = eng._saved["m"]
        assert m.shared == share and (not share or (m.shared_rows > 0 and m.T < m.n_seq * m.S - m.shared_rows + 1))
        res.append((metrics, eng.grads.clone().float()))
        del eng
    (m0, g0), (m1, g1) = res
    # log-probs: another row grouping per attention tile / k-block (bf16 operands): <= 3e-4 relative, as on the LLaVA engines
    # (tests/test_gpu_share_prefix.py); losses and rewards are beta x differences of those log-probs: absolute bound
    for k in ("logps/chosen", "logps/rejected"):
        assert abs(m0[k] - m1[k]) <= 3e-4 * max(1.0, abs(m0[k])), (k, m0[k], m1[k])
    for k in ("loss", "rewards/chosen", "rewards/rejected", "rewards/margins"):
        assert abs(m0[k] - m1[k]) <= 0.1 * 3e-4 * 4 * max(abs(m0["logps/chosen"]), abs(m0["logps/rejected"])), (k, m0[k], m1[k])
    print(f"[shared vs padded {loss_type}] " + ", ".join(f"{k} {m0[k]:.6g}/{m1[k]:.6g}" for k in ("loss", "logps/chosen", "logps/rejected")))
    assert torch.isfinite(g1).all()
    rel = ((g0 - g1).norm() / g0.norm()).item()
    cos = (torch.dot(g0, g1) / (g0.norm() * g1.norm())).item()
    print(f"[shared vs padded {loss_type}] gradient rel l2 {rel:.3e}, cosine {cos:.6f}")
    assert rel < grad_rel and cos > 0.9995, f"gradient rel l2 {rel}, cosine {cos}"


@pytest.mark.parametrize("loss_type", ["sigmoid", "ddpo"])
def test_qwen_shared_prefix_step_equals_padded_step(loss_type):
    import vlrlhf_b200  # noqa: F401
    from oracle import qwen_restate as Q
    from vlrlhf_b200 import config, engine_qwen
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "g9_qwen_small.npz"))
    batch = Q.make_batch(Q.SMALL_QWEN, int(d["n_pairs"]), int(d["text_len"]), int(d["prompt_len"]), int(d["seed"]), ddpo_like=True)

    def build(share):
        eng = engine_qwen.QwenVLDPOEngine(config.SMALL_QWEN, config.TrainConfig(loss_type=loss_type, learning_rate=1e-3,
                                                                                  share_prefix=share), with_optimizer=False)
        eng.init_synthetic(int(d["seed"]))
        return eng

    _shared_vs_padded(build, batch, loss_type)


@pytest.mark.parametrize("loss_type", ["kto_pair", "ddpo"])
def test_xc2_shared_prefix_step_equals_padded_step(loss_type):
    import vlrlhf_b200  # noqa: F401
    from oracle import restate as R
    from oracle import xc2_restate as X
    from vlrlhf_b200 import config, engine_xc2
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "g10_xc2_small.npz"))
    batch = R.make_batch(X.SMALL_XC2, int(d["n_pairs"]), int(d["text_len"]), int(d["prompt_len"]), int(d["seed"]), ddpo_like=True)

    def build(share):
        eng = engine_xc2.XC2DPOEngine(config.SMALL_XC2, config.TrainConfig(loss_type=loss_type, learning_rate=1e-3,
                                                                            share_prefix=share), with_optimizer=False)
        eng.init_synthetic(int(d["seed"]))
        return eng

    _shared_vs_padded(build, batch, loss_type)


@pytest.mark.parametrize("family", ["qwen", "xc2"])
def test_7b_shape_fixtures_with_shared_prefix(family):
    """Configs 3 and 5 at 7B shapes with the pair's prompt + image prefix laid out once: log-probs against the reference's
    fp32 run (g12 / g13), same bound as the padded layout."""
    import vlrlhf_b200  # noqa: F401
    from oracle import restate as R
    from tests import parity_log
    from vlrlhf_b200 import config, host
    G = os.path.join(os.path.dirname(__file__), "golden")
    if family == "qwen":
        from oracle import qwen_restate as Q
        from vlrlhf_b200 import engine_qwen
        tag, rcfg = "g12_config3_qwen7b", Q.QWEN_VL_CHAT
        mk = lambda: engine_qwen.QwenVLDPOEngine(config.QWEN_VL_CHAT, config.TrainConfig(share_prefix=True), with_optimizer=False)
        make_batch = Q.make_batch
    else:
        from oracle import xc2_restate as X
        from vlrlhf_b200 import engine_xc2
        tag, rcfg = "g13_config5_xc2_7b", X.XC2_VL_7B
        mk = lambda: engine_xc2.XC2DPOEngine(config.XC2_VL_7B, config.TrainConfig(share_prefix=True), with_optimizer=False)
        make_batch = R.make_batch
    path = os.path.join(G, tag + ".npz")
    if not os.path.exists(path):
        pytest.skip(f"{tag} fixture not generated")
    d = np.load(path)
    eng = mk()
    eng.init_synthetic(int(d["seed"]))
    batch = make_batch(rcfg, int(d["n_pairs"]), int(d["text_len"]), int(d["prompt_len"]), int(d["seed"]), ddpo_like=True)
    cb = host.concatenated_inputs(batch)
    ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    px = cb["concatenated_img_input_dict"]["pixel_values"]
    plan = eng.host_row_plan(ids, am)
    assert plan["prefix_rows"][0] > 0
    out = eng.step(*eng.prepare_inputs(ids, am, lb, px), train=False, **plan)
    parity_log.check_step(tag + " share_prefix", out, d, rtol=parity_log.RTOL_7B)
    del eng
    torch.cuda.empty_cache()
