"""Shared-prefix rows on the CUDA kernels (TrainConfig.share_prefix):

  * vlb200_attn_fwd_tc_ctx / vlb200_attn_bwd_tc_ctx (context sequences: a suffix sees its pair's prefix, the prefix's dK/dV
    gather both suffixes' queries) against a plain fp32 PyTorch restatement, dh 128 / 64, GQA, ragged lengths;
  * vlb200_share_prefix_rows against its CPU mirror (tests/mock_ops.py), bit-exact (integer work);
  * the whole step: shared == padded log-probs / losses, gradients vs the oracle's autograd, the full-length 7B fixture,
    the Trainer-side boundary.
"""
import math
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


def ref_ctx_attention(q, k, v, lens, starts, ctx, H, KV, dh, scale):
    """fp32 restatement: sequence b's queries see all rows of sequence ctx[b], then their own rows causally."""
    out = torch.zeros(q.shape[0], H * dh, dtype=torch.float32, device=q.device)
    lse = {}
    g = H // KV
    for b in range(len(lens)):
        n, r0 = lens[b], starts[b]
        if n == 0:
            continue
        rows = torch.arange(r0, r0 + n, device=q.device)
        nc = 0
        rows_k = rows
        if ctx[b] >= 0:
            nc, c0 = lens[ctx[b]], starts[ctx[b]]
            rows_k = torch.cat([torch.arange(c0, c0 + nc, device=q.device), rows])
        qq = q[rows].view(n, H, dh).transpose(0, 1)
        kk = k[rows_k].view(-1, KV, dh).transpose(0, 1).repeat_interleave(g, 0)
        vv = v[rows_k].view(-1, KV, dh).transpose(0, 1).repeat_interleave(g, 0)
        s = qq @ kk.transpose(1, 2) * scale
        mask = torch.zeros(n, nc + n, dtype=torch.bool, device=q.device)
        mask[:, nc:] = torch.ones(n, n, dtype=torch.bool, device=q.device).triu(1)
        s = s.masked_fill(mask[None], float("-inf"))
        out[rows] = (torch.softmax(s, -1) @ vv).transpose(0, 1).reshape(n, H * dh)
        lse[b] = torch.logsumexp(s, -1)
    return out, lse


def layout(pre, suf_c, suf_r):
    """rows = [chosen suffixes | prefixes | rejected suffixes] -> (lens, starts, ctx, kids) of the 3*n attention sequences"""
    n = len(pre)
    lens = list(suf_c) + list(pre) + list(suf_r)
    starts = [0]
    for x in lens:
        starts.append(starts[-1] + x)
    ctx = [n + i if pre[i] > 0 else -1 for i in range(n)] + [-1] * n + [n + i if pre[i] > 0 else -1 for i in range(n)]
    kids = [-1] * (6 * n)
    for i in range(n):
        if pre[i] > 0:
            kids[2 * (n + i)], kids[2 * (n + i) + 1] = i, 2 * n + i
    return lens, starts, ctx, kids


@pytest.mark.parametrize("H,KV,dh,pre,suf_c,suf_r", [
    (2, 2, 128, [200, 64], [130, 1], [77, 300]),        # prefixes / suffixes across tile boundaries, a 1-row suffix
    (4, 2, 128, [703, 0, 129], [250, 90, 128], [40, 33, 127]),   # GQA; a pair that shares nothing; LLaVA-like prefix
    (4, 4, 64, [100, 260], [64, 65], [63, 200]),         # dh 64
])
def test_ctx_attention_fwd_bwd(pkg, H, KV, dh, pre, suf_c, suf_r):
    config, engine, host, ops = pkg
    torch.manual_seed(dh + H)
    dev = "cuda"
    lens, starts, ctx, kids = layout(pre, suf_c, suf_r)
    T, B, S = starts[-1], len(lens), max(lens) + 7
    ld = (H + 2 * KV) * dh
    qkv = (torch.randn(T, ld, device=dev) * 0.7).to(torch.bfloat16)
    q, k, v = qkv[:, :H * dh], qkv[:, H * dh:(H + KV) * dh], qkv[:, (H + KV) * dh:]
    scale = 1.0 / math.sqrt(dh)
    i32 = lambda x: torch.tensor(x, dtype=torch.int32, device=dev)  # noqa: E731
    lens_d, starts_d, ctx_d, kids_d = i32(lens), i32(starts), i32(ctx), i32(kids)
    out = torch.full((T, H * dh), float("nan"), dtype=torch.bfloat16, device=dev)
    lse = torch.zeros(B, H, S, dtype=torch.float32, device=dev)
    ops.attn_fwd_tc(q, k, v, out, lse, lens_d, B, S, H, KV, dh, True, scale, row_starts=starts_d, total_rows=T, ctx=ctx_d, kids=kids_d)
    qf, kf, vf = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    want, want_lse = ref_ctx_attention(qf, kf, vf, lens, starts, ctx, H, KV, dh, scale)
    assert torch.isfinite(out.float()).all()          # every row belongs to a sequence: all of them are written
    err = (out.float() - want.detach()).abs().max().item()
    assert err < 2e-2, err
    for b, l in want_lse.items():
        torch.testing.assert_close(lse[b, :, :lens[b]], l.detach(), rtol=1e-3, atol=1e-3)
    # the same sequences WITHOUT context must still equal the plain var-len kernel (ctx = -1 everywhere)
    out2 = torch.empty_like(out)
    lse2 = torch.zeros_like(lse)
    ops.attn_fwd_tc(q, k, v, out2, lse2, lens_d, B, S, H, KV, dh, True, scale, row_starts=starts_d, total_rows=T,
                    ctx=i32([-1] * B), kids=i32([-1] * (2 * B)))
    out3 = torch.empty_like(out)
    ops.attn_fwd_tc(q, k, v, out3, torch.zeros_like(lse), lens_d, B, S, H, KV, dh, True, scale, row_starts=starts_d, total_rows=T)
    assert torch.equal(out2, out3)
    # backward
    dout = (torch.randn(T, H * dh, device=dev) * 0.3).to(torch.bfloat16)
    (want * dout.float()).sum().backward()
    dqkv = torch.full_like(qkv, float("nan"))
    delta = torch.zeros(B, H, S, dtype=torch.float32, device=dev)
    ops.attn_bwd_tc(q, k, v, out, dout, lse, delta, dqkv[:, :H * dh], dqkv[:, H * dh:(H + KV) * dh], dqkv[:, (H + KV) * dh:],
                    lens_d, B, S, H, KV, dh, True, scale, row_starts=starts_d, total_rows=T, ctx=ctx_d, kids=kids_d)
    assert torch.isfinite(dqkv.float()).all()
    for name, got, w in (("dq", dqkv[:, :H * dh], qf.grad), ("dk", dqkv[:, H * dh:(H + KV) * dh], kf.grad),
                         ("dv", dqkv[:, (H + KV) * dh:], vf.grad)):
        rel = ((got.float() - w).norm() / w.norm()).item()
        print(f"[ctx attention H{H} KV{KV} dh{dh}] {name} rel-l2 {rel:.3e}")
        assert rel < 2e-2, (name, rel)
        # per segment too: a prefix's dK/dV sum three groups of queries -- a dropped group would hide in the global norm
        for b in range(B):
            if lens[b]:
                sl = slice(starts[b], starts[b] + lens[b])
                relb = ((got[sl].float() - w[sl]).norm() / w[sl].norm().clamp_min(1e-6)).item()
                assert relb < 4e-2, (name, b, relb)


def test_share_prefix_rows_kernel_matches_mirror(pkg):
    config, engine, host, ops = pkg
    from tests import mock_ops
    cfg = R.SMALL
    batch = R.make_batch(cfg, 3, 60, 12, seed=11, ddpo_like=True)
    batch["rejected_input_ids"] = batch["rejected_input_ids"].clone()
    batch["rejected_input_ids"][2, 0] = 9          # the third pair shares nothing
    cb = host.concatenated_inputs(batch)
    ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    lens = host.merged_seq_lens(ids, am, cfg.image_token_index, cfg.n_patches)
    pre = host.shared_prefix_rows(ids, am, cfg.image_token_index, cfg.n_patches)
    assert pre[2] == 0 and pre[0] > cfg.n_patches
    want = mock_ops.share_prefix_rows(mock_ops.llava_merge_index(ids, am, lb, cfg.n_patches, 3, 1, cfg.image_token_index,
                                                                  cfg.pad_token_id), lens, pre)
    got = ops.share_prefix_rows(ops.llava_merge_index(ids.cuda(), am.cuda(), lb.cuda(), cfg.n_patches, 3, 1, cfg.image_token_index,
                                                      cfg.pad_token_id), lens, pre)
    for k in ("src_map", "pos", "img_pos", "att_starts", "att_lens", "att_ctx", "att_kids"):
        assert torch.equal(getattr(got, k).cpu(), getattr(want, k)), k
    live = want.target >= 0
    assert torch.equal(got.row_of_text.cpu()[live], want.row_of_text[live])
    assert (got.T, got.chosen_rows, got.rejected_rows, got.n_attn_seq) == (want.T, want.chosen_rows, want.rejected_rows, 9)


def _run(pkg, cfg_name, batch, seed, mode, loss_type="sigmoid", **tc):
    config, engine, host, ops = pkg
    t = config.TrainConfig(pack_sequences=(mode == "packed"), share_prefix=(mode == "shared"), loss_type=loss_type,
                           learning_rate=1e-3, **tc)
    eng = engine.LlavaDPOEngine(getattr(config, cfg_name), t, with_optimizer=False)
    eng.init_synthetic(seed)
    cb = host.concatenated_inputs(batch)
    ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    wt = eng.ddpo_weights(ids, am, lb) if loss_type == "ddpo" else None
    plan = eng.host_row_plan(ids, am)
    out = eng.step(*eng.prepare_inputs(ids, am, lb, batch["img_input_dict"]["pixel_values"], wt), train=True, **plan)
    torch.cuda.synchronize()
    return eng, out


@pytest.mark.parametrize("tag,cfg_name,rcfg,shape", [("g4_tiny", "TINY", R.TINY, (2, 24, 8)), ("g4_small", "SMALL", R.SMALL, (2, 96, 24))])
@pytest.mark.parametrize("loss_type", ["sigmoid", "ddpo"])
def test_shared_step_equals_padded_step_and_oracle(pkg, tag, cfg_name, rcfg, shape, loss_type):
    d = np.load(os.path.join(G, tag + ".npz"))
    seed = int(d["seed"])
    batch = R.make_batch(rcfg, *shape, seed, ddpo_like=True)
    e0, o0 = _run(pkg, cfg_name, batch, seed, "padded", loss_type)
    e1, o1 = _run(pkg, cfg_name, batch, seed, "shared", loss_type)
    m = e1._saved["m"]
    assert m.shared and m.T < e0._saved["m"].T - m.shared_rows + 1 and m.shared_rows >= 2 * (shape[2] + rcfg.n_patches - 1)
    key = "policy_logps_ddpo" if loss_type == "ddpo" else "policy_logps"
    full = np.abs(d["policy_logps"])
    for got, k in ((o1.policy_logps, key), (o1.ref_logps, key.replace("policy", "ref"))):
        parity_log.record(f"{tag} share_prefix", k, got.cpu().numpy(), d[k])
        assert (np.abs(got.cpu().numpy() - d[k]) <= 1e-3 * full).all(), k           # vs the reference's own numbers
    parity_log.record(f"{tag} share_prefix", f"{loss_type} logps vs padded step", o1.policy_logps.cpu().numpy(),
                      o0.policy_logps.cpu().numpy(), bound_rel=2e-4)
    parity_log.record(f"{tag} share_prefix", f"{loss_type} losses vs padded step", o1.losses.cpu().numpy(), o0.losses.cpu().numpy(),
                      bound_abs=1.5e-2)
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
        r0 = float((g0[n].float().cpu().view(-1) - w).norm() / w.norm().clamp_min(1e-12))
        r1 = float((g1[n].float().cpu().view(-1) - w).norm() / w.norm().clamp_min(1e-12))
        print(f"[{tag} {loss_type}] {n}: rel-l2 vs oracle autograd padded {r0:.3e} shared {r1:.3e}")
        assert r1 < max(2.0 * r0, 3e-2), (n, r0, r1)


def test_shared_step_checkpointing_and_no_share_pair(pkg):
    config, engine, host, ops = pkg
    seed = 4
    batch = R.make_batch(R.SMALL, 3, 96, 24, seed, ddpo_like=True)
    batch["rejected_input_ids"] = batch["rejected_input_ids"].clone()
    batch["rejected_input_ids"][1, 0] = 7          # pair 1 shares nothing
    e0, o0 = _run(pkg, "SMALL", batch, seed, "padded")
    e1, o1 = _run(pkg, "SMALL", batch, seed, "shared")
    e2, o2 = _run(pkg, "SMALL", batch, seed, "shared", activation_checkpointing=True)
    m = e1._saved["m"]
    assert int(m.att_lens[3 + 1]) == 0 and int(m.att_ctx[1]) == -1
    np.testing.assert_allclose(o1.policy_logps.cpu().numpy(), o0.policy_logps.cpu().numpy(), rtol=2e-4)
    assert torch.equal(o1.policy_logps, o2.policy_logps) and torch.equal(e1.grads, e2.grads)   # recompute is bit-identical
    g0, g1 = e0.grads.float(), e1.grads.float()
    assert float((g0 - g1).norm() / g0.norm()) < 1.5e-2


def test_config2_full_length_7b_shared_prefix(pkg):
    """The headline shape with the prefix shared: 1 pair, text 1024 -> 1599 merged rows per sequence, 703 of them (prompt 128 +
    575 image rows) laid out once; log-probs against the reference's fp32 run (g14)."""
    config, engine, host, ops = pkg
    path = os.path.join(G, "g14_config2_full_7b.npz")
    if not os.path.exists(path):
        pytest.skip("g14 fixture not generated yet")
    d = np.load(path)
    rcfg = R.LLAVA15_7B
    eng = engine.LlavaDPOEngine(config.LLAVA15_7B, config.TrainConfig(share_prefix=True), with_optimizer=False)
    eng.init_synthetic(int(d["seed"]))
    batch = R.make_batch(rcfg, int(d["n_pairs"]), int(d["text_len"]), int(d["prompt_len"]), int(d["seed"]))
    cb = host.concatenated_inputs(batch)
    ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    plan = eng.host_row_plan(ids, am)
    assert plan["prefix_rows"][0] >= int(d["prompt_len"]) + rcfg.n_patches - 1
    out = eng.step(*eng.prepare_inputs(ids, am, lb, batch["img_input_dict"]["pixel_values"]), train=False, **plan)
    parity_log.check_step("g14_config2_full_7b share_prefix", out, d, rtol=parity_log.RTOL_7B)
    del eng
    torch.cuda.empty_cache()


def test_plugin_trainer_loop_with_shared_prefix(pkg):
    config, engine, host, ops = pkg
    from vlrlhf_b200 import plugin
    from tests import trl_loop
    seed = 5
    batches = [R.make_batch(R.SMALL, 2, 96, 24, 20 + i, ddpo_like=True) for i in range(4)]
    kw = dict(learning_rate=2e-3, adam_beta1=0.9, adam_beta2=0.98, adam_eps=1e-6, weight_decay=0.0, max_grad_norm=1.0)
    m1 = plugin.B200LlavaForRL(config.SMALL, config.TrainConfig(share_prefix=True))
    m1.engine.init_synthetic(seed)
    args = trl_loop.training_args(learning_rate=2e-3, adam_beta1=0.9, adam_beta2=0.98, adam_epsilon=1e-6, max_grad_norm=1.0,
                                  gradient_accumulation_steps=2)
    tr = plugin.make_trainer_class(trl_loop.StubDPOTrainer)(m1, None, args=args)
    tr.train_loop(batches)
    assert m1.engine._saved["m"].shared
    m2 = plugin.B200LlavaForRL(config.SMALL, config.TrainConfig(share_prefix=True, gradient_accumulation_steps=2, **kw))
    m2.engine.init_synthetic(seed)
    for i, b in enumerate(batches):
        got = m2.engine.train_step(b)
        assert abs(got["rewards/margins"] - tr.logged[i]["rewards/margins"]) <= 1e-5 * max(1.0, abs(got["rewards/margins"]))
    m2.engine.wait_optimizer()
    torch.cuda.synchronize()
    assert torch.equal(m1.engine.params, m2.engine.params)
