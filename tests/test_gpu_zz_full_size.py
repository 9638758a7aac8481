"""BASELINE.json configs[1] at its full size, through size-independent properties (this file sorts last among the GPU tests
on purpose: it builds a 7B engine and was added after the round's GPU budget was spent -- it has run against the CPU mock of
the ops only -- so under `pytest -x` it cannot hide the results of the parity tests)."""
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
