"""GPU parity of the InternLM-XComposer2-VL + PLoRA/LoRA step (engine_xc2.py through the C ABI) against the fixtures minted
from the reference's InternLMXC2ForRL (tests/golden/g10_xc2_*.npz) and the oracle's autograd."""
import os

import numpy as np
import pytest
import torch

from oracle import restate as R
from oracle import xc2_restate as X

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
CASES = {"g10_xc2_tiny": ("TINY_XC2", X.TINY_XC2), "g10_xc2_small": ("SMALL_XC2", X.SMALL_XC2)}


@pytest.fixture(scope="module")
def pkg():
    import vlrlhf_b200  # noqa: F401
    from vlrlhf_b200 import config, engine_xc2, host, ops
    return config, engine_xc2, host, ops


def build(pkg, tag, loss_type="sigmoid", with_optimizer=False, **tc):
    config, EX, host, ops = pkg
    name, xcfg = CASES[tag]
    d = np.load(os.path.join(G, tag + ".npz"))
    eng = EX.XC2DPOEngine(getattr(config, name), config.TrainConfig(loss_type=loss_type, learning_rate=1e-3, **tc),
                          with_optimizer=with_optimizer)
    eng.init_synthetic(int(d["seed"]))
    batch = R.make_batch(xcfg, int(d["n_pairs"]), int(d["text_len"]), int(d["prompt_len"]), int(d["seed"]), ddpo_like=True)
    return eng, xcfg, d, batch


def test_scatter_add_rows_kernel(pkg):
    config, EX, host, ops = pkg
    g = torch.Generator(device="cuda").manual_seed(0)
    n, T, cols = 300, 1000, 256
    src = torch.randn(n, cols, device="cuda", generator=g).to(torch.bfloat16)
    idx = torch.randperm(T, device="cuda", generator=g)[:n].to(torch.int32)
    idx[7] = -1
    for dt in (torch.bfloat16, torch.float32):
        full = torch.randn(T, 2 * cols, device="cuda", generator=g).to(dt)
        want = full.clone()
        ok = idx >= 0
        want[idx[ok].long(), cols:] = (want[idx[ok].long(), cols:].float() + 0.5 * src[ok].float()).to(dt)
        ops.scatter_add_rows(src, idx, full[:, cols:], 0.5)   # destination = a column slice
        assert torch.equal(full, want), dt


@pytest.mark.parametrize("tag", list(CASES))
def test_xc2_weights_and_forward_parity(pkg, tag):
    config, EX, host, ops = pkg
    eng, xcfg, d, batch = build(pkg, tag)
    w, lora = X.make_weights(xcfg, int(d["seed"]))
    st = eng.hf_state("policy")
    for k, v in list(lora.items()) + [(k, v) for k, v in w.items() if not k.startswith("vit.")]:
        assert torch.equal(st[k].float().cpu().reshape(v.shape), v), k
    cb = host.concatenated_inputs(batch)
    ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    px = cb["concatenated_img_input_dict"]["pixel_values"]
    out = eng.step(*eng.prepare_inputs(ids, am, lb, px), train=False)
    pol, ref = out.policy_logps.cpu().numpy(), out.ref_logps.cpu().numpy()
    print(f"[{tag}] policy rel err", np.abs(pol / d["policy_logps"] - 1), "ref rel err", np.abs(ref / d["ref_logps"] - 1))
    np.testing.assert_allclose(pol, d["policy_logps"], rtol=1e-3)
    np.testing.assert_allclose(ref, d["ref_logps"], rtol=1e-3)
    slack = 0.1 * 1e-3 * np.abs(d["policy_logps"]).max() * 4
    np.testing.assert_allclose(out.losses.cpu().numpy(), d["sigmoid_losses"], atol=slack)
    wt = eng.ddpo_weights(ids, am, lb)
    out = eng.step(*eng.prepare_inputs(ids, am, lb, px, wt), train=False)
    for got, key, full in ((out.policy_logps, "policy_logps_ddpo", "policy_logps"), (out.ref_logps, "ref_logps_ddpo", "ref_logps")):
        assert (np.abs(got.cpu().numpy() - d[key]) <= 1e-3 * np.abs(d[full])).all(), key
    eng.tc.loss_type = "kto_pair"
    out = eng.step(*eng.prepare_inputs(ids, am, lb, px), train=False)
    np.testing.assert_allclose(out.losses.cpu().numpy(), d["kto_pair_losses"], atol=slack)


@pytest.mark.parametrize("tag", list(CASES))
def test_xc2_adapter_gradients_match_oracle_autograd(pkg, tag):
    config, EX, host, ops = pkg
    res = {}
    for ckpt in (False, True):
        eng, xcfg, d, batch = build(pkg, tag, loss_type="kto_pair", activation_checkpointing=ckpt)
        eng.train_step(batch, train=True)
        torch.cuda.synchronize()
        res[ckpt] = eng.grads.clone()
    assert torch.equal(res[False], res[True])
    got = {k: v.float().cpu() for k, v in eng.hf_state("grad").items()}
    w, lora = X.make_weights(xcfg, int(d["seed"]))
    leaves = {k: v.clone().requires_grad_(True) for k, v in lora.items()}
    loss, _, _ = X.get_batch_loss_metrics(xcfg, w, leaves, batch, loss_type="kto_pair")
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


def test_xc2_train_step_updates_only_adapters(pkg):
    config, EX, host, ops = pkg
    eng, xcfg, d, batch = build(pkg, "g10_xc2_small", loss_type="kto_pair", with_optimizer=True, weight_decay=0.1)
    base0, vis0 = eng.bparams.clone(), eng.vparams.clone()
    losses = [eng.train_step(batch, train=True)["loss"] for _ in range(4)]
    eng.wait_optimizer()
    assert losses[-1] < losses[0], losses
    assert torch.equal(eng.bparams, base0) and torch.equal(eng.vparams, vis0)
    assert torch.equal(eng.params, eng.master.to(torch.bfloat16)) and torch.isfinite(eng.master).all()
