"""GPU parity of the Qwen-VL + LoRA DPO step (engine_qwen.py through the C ABI) against the fixtures minted from the
reference's vendored QWenLMHeadModel / VisionTransformer (tests/golden/g9_qwen_*.npz) and the oracle's autograd."""
import os

import numpy as np
import pytest
import torch

from oracle import qwen_restate as Q
from oracle import restate as R
from tests import parity_log

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
CASES = {"g9_qwen_tiny": ("TINY_QWEN", Q.TINY_QWEN), "g9_qwen_small": ("SMALL_QWEN", Q.SMALL_QWEN),
         # BASELINE.json configs[2] at 7B shapes (ViT-bigG 48 x 1664 / dh 104, V 151936, LoRA r 64): forward parity only
         "g12_config3_qwen7b": ("QWEN_VL_CHAT", Q.QWEN_VL_CHAT)}
SMALL_TAGS = ["g9_qwen_tiny", "g9_qwen_small"]


@pytest.fixture(scope="module")
def pkg():
    import vlrlhf_b200  # noqa: F401
    from vlrlhf_b200 import config, engine_qwen, host, ops
    return config, engine_qwen, host, ops


def build(pkg, tag, loss_type="sigmoid", with_optimizer=False, **tc):
    config, EQ, host, ops = pkg
    name, qcfg = CASES[tag]
    if not os.path.exists(os.path.join(G, tag + ".npz")):
        pytest.skip(f"{tag} fixture not generated yet (oracle/make_fixtures.py --config3)")
    d = np.load(os.path.join(G, tag + ".npz"))
    eng = EQ.QwenVLDPOEngine(getattr(config, name), config.TrainConfig(loss_type=loss_type, learning_rate=1e-3, **tc),
                             with_optimizer=with_optimizer)
    eng.init_synthetic(int(d["seed"]))
    batch = Q.make_batch(qcfg, int(d["n_pairs"]), int(d["text_len"]), int(d["prompt_len"]), int(d["seed"]), ddpo_like=True)
    return eng, qcfg, d, batch


def test_lora_shaped_gemms(pkg):
    """The skinny GEMMs of the adapter path (N or M or K = r, 2r) on the tcgen05 kernels vs fp32 matmul."""
    config, EQ, host, ops = pkg
    g = torch.Generator(device="cuda").manual_seed(0)
    rn = lambda *s: (torch.randn(*s, device="cuda", generator=g) * 0.5).to(torch.bfloat16)  # noqa: E731
    T, d, r, ff = 1000, 512, 16, 1024
    x, A, B = rn(T, d), rn(r, d), rn(3 * d, r)
    t32 = ops.gemm(x, A, out_dtype=torch.float32)                                   # N = r
    torch.testing.assert_close(t32, x.float() @ A.float().t(), rtol=1e-3, atol=1e-2)
    ts = torch.empty(T, r, dtype=torch.bfloat16, device="cuda")
    ops.cast_f32_to_bf16(t32.view(-1), ts.view(-1), 0.5)
    u = ops.gemm(ts, B)                                                             # K = r
    torch.testing.assert_close(u.float(), ts.float() @ B.float().t(), rtol=2e-2, atol=2e-2)
    dy = rn(T, 3 * d)
    dB = ops.gemm(dy, ts, a_kmajor=False, b_kmajor=False)                           # [3d, r]: N = r, MN-major operands
    torch.testing.assert_close(dB.float(), dy.float().t() @ ts.float(), rtol=2e-2, atol=0.3)
    dt = ops.gemm(dy, B, b_kmajor=False, out_dtype=torch.float32)                   # [T, r]
    torch.testing.assert_close(dt, dy.float() @ B.float(), rtol=1e-3, atol=5e-2)
    dtb = dt.to(torch.bfloat16)
    dA = ops.gemm(dtb, x, a_kmajor=False, b_kmajor=False)                           # [r, d]: M = r
    torch.testing.assert_close(dA.float(), dtb.float().t() @ x.float(), rtol=2e-2, atol=1.0)
    base = rn(T, d)
    acc = base.clone()
    ops.gemm(dtb, A, b_kmajor=False, out=acc, accumulate=True)                      # K = r, accumulate into bf16
    torch.testing.assert_close(acc.float(), base.float() + dtb.float() @ A.float(), rtol=2e-2, atol=0.1)
    # column-sliced fp32 outputs / bf16 operands (the fused gate|up adapters)
    gu, tsg = rn(T, 2 * ff), rn(T, 2 * r)
    out32 = torch.zeros(T, 2 * r, device="cuda")
    B2 = rn(ff, r)
    ops.gemm(gu[:, :ff], B2, b_kmajor=False, out=out32[:, :r])
    torch.testing.assert_close(out32[:, :r], gu[:, :ff].float() @ B2.float(), rtol=1e-3, atol=0.1)
    assert float(out32[:, r:].abs().max()) == 0.0
    dB2 = ops.gemm(gu[:, ff:], tsg[:, r:], a_kmajor=False, b_kmajor=False)
    torch.testing.assert_close(dB2.float(), gu[:, ff:].float().t() @ tsg[:, r:].float(), rtol=2e-2, atol=0.5)


def test_qwen_merge_kernel_matches_mock(pkg):
    config, EQ, host, ops = pkg
    from tests import mock_ops
    qcfg = Q.SMALL_QWEN
    batch = Q.make_batch(qcfg, 3, 140, 72, seed=3)
    cb = R.concatenated_inputs(batch)
    ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    want = mock_ops.qwen_merge_index(ids, am, lb, qcfg.n_queries, 3, 1, qcfg.image_start_id)
    got = ops.qwen_merge_index(ids.cuda(), am.cuda(), lb.cuda(), qcfg.n_queries, 3, 1, qcfg.image_start_id)
    assert int(got.status) == 0
    for k in ("src_map", "labels", "mask", "pos", "seqlens", "row_of_text", "target"):
        assert torch.equal(getattr(got, k).cpu(), getattr(want, k)), k
    bad = ids.clone(); bad[1, 2 + qcfg.n_queries] = 7
    st = ops.qwen_merge_index(bad.cuda(), am.cuda(), lb.cuda(), qcfg.n_queries, 3, 1, qcfg.image_start_id)
    eng = EQ.QwenVLDPOEngine(config.TINY_QWEN, config.TrainConfig(), with_optimizer=False)
    with pytest.raises(ValueError):
        eng.check_merge_status(st)


@pytest.mark.parametrize("tag", SMALL_TAGS)
def test_qwen_weights_bit_exact_and_visual_tower(pkg, tag):
    config, EQ, host, ops = pkg
    eng, qcfg, d, batch = build(pkg, tag)
    w, lora = Q.make_weights(qcfg, int(d["seed"]))
    st = eng.hf_state("policy")
    for k, v in lora.items():
        assert torch.equal(st[k].float().cpu(), v), k
    for k, v in w.items():
        if not k.startswith("transformer.visual."):
            assert torch.equal(st[k].float().cpu().reshape(v.shape), v), k
    with torch.no_grad():
        want = Q.visual_forward(qcfg, w, batch["img_input_dict"]["pixel_values"])
    got = eng.vision_features(batch["img_input_dict"]["pixel_values"].cuda()).float().cpu().view(want.shape)
    err = (got - want).abs().max().item() / want.abs().max().item()
    print(f"[{tag}] resampler output max err / max |x| = {err:.4g}")
    assert err < 3e-2


@pytest.mark.parametrize("tag", list(CASES))
def test_qwen_forward_logps_loss_and_ddpo_parity(pkg, tag):
    config, EQ, host, ops = pkg
    eng, qcfg, d, batch = build(pkg, tag)
    cb = host.concatenated_inputs(batch)
    ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    px = cb["concatenated_img_input_dict"]["pixel_values"]
    out = eng.step(*eng.prepare_inputs(ids, am, lb, px), train=False)
    parity_log.check_step(tag, out, d, rtol=parity_log.RTOL_7B if tag.startswith("g12") else 1e-3)
    wt = eng.ddpo_weights(ids, am, lb)
    out = eng.step(*eng.prepare_inputs(ids, am, lb, px, wt), train=False)
    # DDPO sums a subset of the same per-token terms: absolute error bounded by the full sum's 1e-3 budget
    for got, key, full in ((out.policy_logps, "policy_logps_ddpo", "policy_logps"), (out.ref_logps, "ref_logps_ddpo", "ref_logps")):
        err = np.abs(got.cpu().numpy() - d[key])
        parity_log.record(tag, key, got.cpu().numpy(), d[key], note="subset of the per-token terms; bound = 1e-3 x |full sum|")
        assert (err <= 1e-3 * np.abs(d[full])).all(), (key, err)
    if tag.startswith("g12"):
        del eng
        torch.cuda.empty_cache()


@pytest.mark.parametrize("tag", SMALL_TAGS)
def test_qwen_adapter_gradients_match_oracle_autograd(pkg, tag):
    config, EQ, host, ops = pkg
    res = {}
    for ckpt in (False, True):
        eng, qcfg, d, batch = build(pkg, tag, activation_checkpointing=ckpt)
        eng.train_step(batch, train=True)
        torch.cuda.synchronize()
        res[ckpt] = eng.grads.clone()
    assert torch.equal(res[False], res[True])
    got = {k: v.float().cpu() for k, v in eng.hf_state("grad").items()}
    w, lora = Q.make_weights(qcfg, int(d["seed"]))
    leaves = {k: v.clone().requires_grad_(True) for k, v in lora.items()}
    loss, _, _ = Q.get_batch_loss_metrics(qcfg, w, leaves, batch)
    loss.backward()
    worst = 0.0
    for k, leaf in leaves.items():
        g, want = got[k], leaf.grad
        assert torch.isfinite(g).all(), k
        rel = (g - want).norm().item() / max(want.norm().item(), 1e-12)
        worst = max(worst, rel)
        assert rel < 6e-2, f"{k}: rel l2 err {rel:.4g}"
        assert torch.nn.functional.cosine_similarity(g.flatten(), want.flatten(), dim=0).item() > 0.998, k
    print(f"[{tag}] worst adapter-gradient rel-l2 error {worst:.4g}")


def test_qwen_train_step_updates_only_adapters(pkg):
    config, EQ, host, ops = pkg
    eng, qcfg, d, batch = build(pkg, "g9_qwen_small", with_optimizer=True, weight_decay=0.05)
    base0, vis0 = eng.bparams.clone(), eng.vparams.clone()
    losses = [eng.train_step(batch, train=True)["loss"] for _ in range(4)]
    eng.wait_optimizer()
    assert losses[-1] < losses[0], losses
    assert torch.equal(eng.bparams, base0) and torch.equal(eng.vparams, vis0)
    assert torch.equal(eng.params, eng.master.to(torch.bfloat16)) and torch.isfinite(eng.master).all()
    w, lora = Q.make_weights(qcfg, int(d["seed"]))
    with torch.no_grad():
        loss, metrics, _ = Q.get_batch_loss_metrics(qcfg, w, lora, batch)
    eng2, _, _, _ = build(pkg, "g9_qwen_small", with_optimizer=True)
    got = eng2.train_step(batch, train=True)
    assert abs(got["loss"] - float(loss)) < 5e-2
    for k in ("logps/chosen", "logps/rejected"):
        assert abs(got[k] / float(metrics[k]) - 1) < 1e-3, k
    for k in ("logits/chosen", "logits/rejected"):
        assert abs(got[k] - float(metrics[k])) < 5e-3, k
