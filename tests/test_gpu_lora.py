"""GPU parity of the LLaVA-1.5 / LLaVA-Next + LoRA step (engine_lora.py through the C ABI) against the fixtures minted from
the reference's LlavaForRL / LlavaNextForRL with hand-applied peft-style adapters (tests/golden/g11_*.npz) and the oracle's
autograd."""
import os

import numpy as np
import pytest
import torch

from oracle import lora_restate as LR
from oracle import restate as R
from tests import parity_log

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
CASES = {"g11_lora_tiny": ("TINY_LORA", LR.TINY_LORA), "g11_lora_small": ("SMALL_LORA", LR.SMALL_LORA),
         "g11_next_lora_tiny": ("TINY_NEXT_LORA", LR.TINY_NEXT_LORA), "g11_next_lora_small": ("SMALL_NEXT_LORA", LR.SMALL_NEXT_LORA)}


@pytest.fixture(scope="module")
def pkg():
    import vlrlhf_b200  # noqa: F401
    from vlrlhf_b200 import config, engine_lora, host, ops
    return config, engine_lora, host, ops


def build(pkg, tag, loss_type="sigmoid", with_optimizer=False, **tc):
    config, EL, host, ops = pkg
    name, cfg = CASES[tag]
    d = np.load(os.path.join(G, tag + ".npz"))
    eng = EL.LlavaLoRADPOEngine(getattr(config, name), config.TrainConfig(loss_type=loss_type, learning_rate=1e-3, **tc),
                                with_optimizer=with_optimizer)
    eng.init_synthetic(int(d["seed"]))
    sizes = [tuple(int(x) for x in s) for s in d["image_sizes"]] if "image_sizes" in d.files else None
    batch = R.make_batch(cfg, int(d["n_pairs"]), int(d["text_len"]), int(d["prompt_len"]), int(d["seed"]), ddpo_like=True,
                         image_sizes=sizes)
    return eng, cfg, d, batch


@pytest.mark.parametrize("tag", list(CASES))
def test_lora_weights_and_forward_parity(pkg, tag):
    config, EL, host, ops = pkg
    eng, cfg, d, batch = build(pkg, tag)
    w, lora = LR.make_weights(cfg, int(d["seed"]))
    st = eng.hf_state("policy")
    for k, v in list(lora.items()) + [(k, v) for k, v in w.items() if k in st]:
        assert torch.equal(st[k].float().cpu().reshape(v.shape), v), k
    cb = host.concatenated_inputs(batch)
    ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    img = cb["concatenated_img_input_dict"]
    px, sizes = img["pixel_values"], img.get("image_sizes")
    out = eng.step(*eng.prepare_inputs(ids, am, lb, px, None, sizes), train=False)
    parity_log.check_step(tag, out, d)
    wt = eng.ddpo_weights(ids, am, lb, sizes)
    out = eng.step(*eng.prepare_inputs(ids, am, lb, px, wt, sizes), train=False)
    for got, key, full in ((out.policy_logps, "policy_logps_ddpo", "policy_logps"), (out.ref_logps, "ref_logps_ddpo", "ref_logps")):
        assert (np.abs(got.cpu().numpy() - d[key]) <= 1e-3 * np.abs(d[full])).all(), key
    eng.tc.loss_type = "kto_pair"
    out = eng.step(*eng.prepare_inputs(ids, am, lb, px, None, sizes), train=False)
    b = parity_log.LOSS_ABS_BOUNDS.get(tag, 0.1 * 1e-3 * np.abs(d["policy_logps"]).max() * 4)
    parity_log.record(tag, "kto_pair_losses", out.losses.cpu().numpy(), d["kto_pair_losses"], bound_abs=b)


@pytest.mark.parametrize("tag,loss_type", [("g11_lora_tiny", "sigmoid"), ("g11_lora_small", "kto_pair"),
                                           ("g11_next_lora_tiny", "ddpo"), ("g11_next_lora_small", "sigmoid")])
def test_lora_adapter_gradients_match_oracle_autograd(pkg, tag, loss_type):
    config, EL, host, ops = pkg
    res = {}
    for ckpt in (False, True):
        eng, cfg, d, batch = build(pkg, tag, loss_type=loss_type, activation_checkpointing=ckpt)
        eng.train_step(batch, train=True)
        torch.cuda.synchronize()
        res[ckpt] = eng.grads.clone()
    assert torch.equal(res[False], res[True])
    got = {k: v.float().cpu() for k, v in eng.hf_state("grad").items()}
    w, lora = LR.make_weights(cfg, int(d["seed"]))
    leaves = {k: v.clone().requires_grad_(True) for k, v in lora.items()}
    loss, _, _ = LR.get_batch_loss_metrics(cfg, w, leaves, batch, loss_type=loss_type)
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


def test_lora_train_step_updates_only_adapters(pkg):
    config, EL, host, ops = pkg
    eng, cfg, d, batch = build(pkg, "g11_lora_small", with_optimizer=True, weight_decay=0.1)
    base0, vis0 = eng.bparams.clone(), eng.vparams.clone()
    losses = [eng.train_step(batch, train=True)["loss"] for _ in range(4)]
    eng.wait_optimizer()
    assert losses[-1] < losses[0], losses
    assert torch.equal(eng.bparams, base0) and torch.equal(eng.vparams, vis0)
    assert torch.equal(eng.params, eng.master.to(torch.bfloat16)) and torch.isfinite(eng.master).all()
