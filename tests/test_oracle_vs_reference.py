"""The oracle restatement against the LIVE reference (build container only: skipped where /root/reference is absent, e.g. on
the GPU box -- there the committed vectors in tests/golden/, minted from these same reference functions by
oracle/make_fixtures.py, are the pin).

The reference's own code runs here through oracle/ref_shim.py (stub trl / peft / accelerate / deepspeed modules, nothing of
the reference is copied): VLDPOTrainer.get_batch_logps, VLDPOTrainer.dpo_loss, diff_lib.get_diff_ids and LlavaForRL /
LlavaNextForRL forward incl. _merge_input_ids_with_image_features.
"""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import ref_shim, restate as R

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    T, get_diff_ids, _ = ref_shim.reference_symbols()
    return T, get_diff_ids


def _labels(g, n_seq, S, V):
    lb = torch.randint(3, V, (n_seq, S), generator=g)
    for b in range(n_seq):
        lb[b, : int(torch.randint(1, S // 2, (1,), generator=g))] = -100          # prompt
        lb[b, S - int(torch.randint(0, S // 4, (1,), generator=g)):] = -100        # right padding
    return lb


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("average", [False, True])
def test_get_batch_logps_equals_reference(ref, dtype, average):
    T, _ = ref
    g = torch.Generator().manual_seed(11)
    logits = (torch.randn(4, 37, 211, generator=g) * 3).to(dtype)
    lb = _labels(g, 4, 37, 211)
    want = T.get_batch_logps(logits, lb, average_log_prob=average, label_pad_token_id=-100, is_encoder_decoder=False)
    got = R.get_batch_logps(logits, lb, average_log_prob=average)
    assert got.dtype == want.dtype and torch.equal(got, want)
    with pytest.raises(ValueError):
        T.get_batch_logps(logits[:, :-1], lb)
    with pytest.raises(ValueError):
        R.get_batch_logps(logits[:, :-1], lb)


def test_get_batch_logps_ddpo_mask_equals_reference(ref):
    """mask_shared_tokens (base/trainer.py:169-184): chosen/rejected differ by a few substituted spans."""
    T, _ = ref
    g = torch.Generator().manual_seed(5)
    S, V = 120, 97
    chosen = _labels(g, 2, S, V)
    rejected = chosen.clone()
    for b in range(2):
        for _ in range(3):
            a = int(torch.randint(S // 2, S - 12, (1,), generator=g))
            n = int(torch.randint(1, 8, (1,), generator=g))
            rejected[b, a:a + n] = torch.where(rejected[b, a:a + n] == -100, rejected[b, a:a + n],
                                               torch.randint(3, V, (n,), generator=g))
    lb = torch.cat([chosen, rejected])
    logits = torch.randn(4, S, V, generator=g)
    want = T.get_batch_logps(logits, lb, mask_shared_tokens=True)
    got = R.get_batch_logps(logits, lb, mask_shared_tokens=True)
    assert torch.equal(got, want)
    assert not torch.equal(want, T.get_batch_logps(logits, lb))     # the mask removes shared tokens


@pytest.mark.parametrize("loss_type", ["sigmoid", "ddpo", "hinge", "ipo", "kto_pair"])
@pytest.mark.parametrize("ls,reference_free", [(0.0, False), (0.1, False), (0.0, True)])
def test_dpo_loss_equals_reference(ref, loss_type, ls, reference_free):
    T, _ = ref
    g = torch.Generator().manual_seed(3)
    pc, pr, rc, rr = (torch.randn(5, generator=g) * 4 - 60 for _ in range(4))
    me = SimpleNamespace(beta=0.1, label_smoothing=ls, loss_type=loss_type, reference_free=reference_free,
                         accelerator=SimpleNamespace(device="cpu"))
    want = T.dpo_loss(me, pc, pr, rc, rr)
    got = R.dpo_loss(pc, pr, rc, rr, 0.1, ls, loss_type, reference_free)
    for a, b in zip(got, want):
        assert a.shape == b.shape
        torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-6)
    with pytest.raises(ValueError):
        T.dpo_loss(SimpleNamespace(**{**me.__dict__, "loss_type": "nope"}), pc, pr, rc, rr)
    with pytest.raises(ValueError):
        R.dpo_loss(pc, pr, rc, rr, 0.1, ls, "nope", reference_free)


def test_get_diff_ids_equals_reference(ref):
    _, get_diff_ids = ref
    rng = np.random.RandomState(0)
    for case in range(60):
        n = int(rng.randint(5, 400))
        a = rng.randint(0, 12 if case % 3 == 0 else 400, size=n).tolist()   # small alphabets trigger autojunk (n >= 200)
        b = list(a)
        for _ in range(int(rng.randint(0, 6))):
            i = int(rng.randint(0, len(b)))
            k = int(rng.randint(0, 9))
            b[i:i + k] = rng.randint(0, 400, size=int(rng.randint(0, 9))).tolist()
        want = get_diff_ids(a, b, min_match_size=3)
        got = R.get_diff_ids(a, b, 3)
        assert (list(got[0]), list(got[1])) == (list(want[0]), list(want[1])), case


@pytest.mark.parametrize("name,sizes", [("TINY", None), ("TINY_NEXT", [(28, 28), (20, 50), (60, 25)])])
def test_llava_forward_and_merge_equal_reference(name, sizes):
    """LlavaForRL / LlavaNextForRL.forward (merge included) on seeded weights and a ragged collated batch: merged labels and
    image-position map bit-equal, per-sequence log-probs of the reference's logits == the oracle's concatenated_forward."""
    from oracle import make_fixtures as MF
    cfg = getattr(R, name)
    seed = 4
    batch = R.make_batch(cfg, 3, 24, 8, seed, ddpo_like=True, image_sizes=sizes)
    model = MF.build_reference_model(cfg, MF.streamed_weights(cfg, seed, "policy"))
    want_logps, out = MF.reference_concatenated_forward(model, cfg, batch, "sigmoid")
    w, _ = R.make_policy_and_ref(cfg, seed)
    cb = R.concatenated_inputs(batch, -100, 0)
    with torch.no_grad():
        pc, pr, pcl, prl = R.concatenated_forward(cfg, w, batch)
        _, labels, imap = R.model_forward(cfg, w, cb["concatenated_input_ids"], cb["concatenated_attention_mask"],
                                          cb["concatenated_labels"], **cb["concatenated_img_input_dict"])
    assert torch.equal(labels, out.labels)
    assert torch.equal(imap, out.image_position_map)
    np.testing.assert_allclose(torch.cat([pc, pr]).numpy(), want_logps.numpy(), rtol=2e-5, atol=2e-4)
    np.testing.assert_allclose(torch.cat([pcl, prl]).numpy(), out.logits.float().numpy(), rtol=1e-3, atol=2e-4)


@pytest.mark.parametrize("name,sizes", [("TINY", None), ("TINY_NEXT", [(28, 28), (20, 50), (60, 25)])])
def test_left_padded_reference_equals_right_padded_path(name, sizes):
    """f-2 (left padding): the reference's LlavaForRL / LlavaNextForRL on a LEFT-padded batch gives, per sequence, the
    log-probs of the right-padded form -- which is what host.right_pad_valid_tokens hands to the engine."""
    from oracle import make_fixtures as MF
    import vlrlhf_b200  # noqa: F401
    from vlrlhf_b200 import host
    cfg, seed = getattr(R, name), 9
    batch = R.make_batch(cfg, 3, 24, 8, seed, ddpo_like=True, image_sizes=sizes)
    left = dict(batch)
    for side in ("chosen", "rejected"):
        ids, am, lb = (batch[f"{side}_{k}"].clone() for k in ("input_ids", "attention_mask", "labels"))
        for b in range(ids.shape[0]):
            n = int(am[b].sum())
            k = ids.shape[1] - n
            ids[b] = torch.cat([torch.zeros(k, dtype=ids.dtype), ids[b, :n]])
            lb[b] = torch.cat([torch.full((k,), -100, dtype=lb.dtype), lb[b, :n]])
            am[b] = torch.cat([torch.zeros(k, dtype=am.dtype), am[b, :n]])
        left[f"{side}_input_ids"], left[f"{side}_attention_mask"], left[f"{side}_labels"] = ids, am, lb
    assert int(left["rejected_attention_mask"][:, 0].min()) == 0
    model = MF.build_reference_model(cfg, MF.streamed_weights(cfg, seed, "policy"))
    want_left, _ = MF.reference_concatenated_forward(model, cfg, left, "sigmoid")
    want_right, _ = MF.reference_concatenated_forward(model, cfg, batch, "sigmoid")
    np.testing.assert_allclose(want_left.numpy(), want_right.numpy(), rtol=2e-5, atol=2e-4)   # the reference itself agrees
    cb = R.concatenated_inputs(left, -100, 0)
    ids, am, lb = host.right_pad_valid_tokens(cb["concatenated_input_ids"], cb["concatenated_attention_mask"],
                                              cb["concatenated_labels"], 0, -100)
    cr = R.concatenated_inputs(batch, -100, 0)
    assert torch.equal(am, cr["concatenated_attention_mask"]) and torch.equal(lb, cr["concatenated_labels"])
    att = am == 1
    assert torch.equal(ids[att], cr["concatenated_input_ids"][att])
    w, _ = R.make_policy_and_ref(cfg, seed)
    with torch.no_grad():
        logits, labels, _ = R.model_forward(cfg, w, ids, am, lb, **cb["concatenated_img_input_dict"])
        got = R.get_batch_logps(logits, labels)
    np.testing.assert_allclose(got.numpy(), want_left.numpy(), rtol=2e-5, atol=2e-4)


def test_reference_ddpo_depends_on_the_padding_side(ref):
    """Why host.right_pad_valid_tokens refuses loss_type='ddpo' on left-padded batches: the reference's own get_batch_logps
    with mask_shared_tokens gives different results for the same pairs padded on the other side (its diff runs over the
    padded label sequences, masked positions rewritten to token 0, base/trainer.py:166,177-180)."""
    T, _ = ref
    batch = R.make_batch(R.TINY, 3, 24, 8, 0, ddpo_like=True)
    cb = R.concatenated_inputs(batch, -100, 0)
    right, am = cb["concatenated_labels"].clone(), cb["concatenated_attention_mask"]
    V = 50
    right[right >= V] = right[right >= V] % (V - 3) + 3
    n_seq, L = right.shape
    logits_r = torch.randn(n_seq, L, V, generator=torch.Generator().manual_seed(0))
    left, logits_l = torch.full_like(right, -100), torch.zeros_like(logits_r)
    for b in range(n_seq):
        n = int(am[b].sum())
        left[b, L - n:], logits_l[b, L - n:] = right[b, :n], logits_r[b, :n]
    torch.testing.assert_close(T.get_batch_logps(logits_r, right), T.get_batch_logps(logits_l, left))   # plain sum: no
    ddpo_r = T.get_batch_logps(logits_r, right, mask_shared_tokens=True)
    ddpo_l = T.get_batch_logps(logits_l, left, mask_shared_tokens=True)
    assert float((ddpo_r - ddpo_l).abs().max()) > 1.0                                                   # DDPO: yes
    torch.testing.assert_close(R.get_batch_logps(logits_l, left, mask_shared_tokens=True), ddpo_l)      # (the oracle follows)


@pytest.mark.parametrize("minter", ["g1_logps", "g2_loss", "g3_ddpo", "g4", "g6_next", "g7_clip_preprocess", "g9_qwen",
                                    "g10_xc2", "g11_lora"])
def test_committed_fixtures_regenerate_from_the_reference(minter, tmp_path, monkeypatch):
    """Every committed vector in tests/golden/ (the 7B-shape ones aside: minutes and ~30 GB) comes out of the reference's own
    functions again, value for value: LlavaForRL, LlavaNextForRL, the vendored QWenLMHeadModel and InternLMXC2ForRL with
    hand-applied peft-style adapters, get_batch_logps, dpo_loss, get_diff_ids, Pillow + CLIPImageProcessor."""
    import os
    from oracle import make_fixtures as MF
    golden = MF.GOLDEN
    monkeypatch.setattr(MF, "GOLDEN", str(tmp_path))
    if minter == "g4":
        MF.g45_llava("g4_tiny", R.TINY, 2, 24, 8, 0, ddpo=True)
        MF.g45_llava("g4_small", R.SMALL, 2, 96, 24, 0, ddpo=True)
    else:
        getattr(MF, minter)()
    made = sorted(os.listdir(tmp_path))
    assert made
    for f in made:
        new, old = np.load(os.path.join(tmp_path, f), allow_pickle=True), np.load(os.path.join(golden, f), allow_pickle=True)
        assert set(new.files) == set(old.files), f
        for k in old.files:
            a, b = new[k], old[k]
            assert a.shape == b.shape and a.dtype == b.dtype, (f, k)
            if a.dtype.kind in "fc":
                np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-6, err_msg=f"{f}:{k}")
            else:
                assert np.array_equal(a, b), (f, k)


def test_two_images_per_sequence_reference_equals_oracle():
    """f-2 (multi-image): the reference's LlavaForRL with two <image> placeholders per sequence (pixel_values [2*B*2, ...] after
    concatenated_inputs' [v, v] duplication) against the oracle's restated merge: merged labels / image map bit-equal,
    log-probs equal."""
    from oracle import make_fixtures as MF
    cfg, seed = R.TINY, 3
    batch = R.make_batch(cfg, 2, 24, 8, seed, ddpo_like=True)
    for side in ("chosen", "rejected"):
        batch[f"{side}_input_ids"][:, 4] = cfg.image_token_index
    batch["img_input_dict"] = {"pixel_values": torch.randn(4, 3, cfg.image_size, cfg.image_size,
                                                          generator=torch.Generator().manual_seed(seed))}
    model = MF.build_reference_model(cfg, MF.streamed_weights(cfg, seed, "policy"))
    want, out = MF.reference_concatenated_forward(model, cfg, batch, "sigmoid")
    assert out.labels.shape[1] == 24 + 2 * (cfg.n_patches - 1)
    w, _ = R.make_policy_and_ref(cfg, seed)
    cb = R.concatenated_inputs(batch, -100, 0)
    with torch.no_grad():
        pc, pr, _, _ = R.concatenated_forward(cfg, w, batch)
        _, labels, imap = R.model_forward(cfg, w, cb["concatenated_input_ids"], cb["concatenated_attention_mask"],
                                          cb["concatenated_labels"], **cb["concatenated_img_input_dict"])
    assert torch.equal(labels, out.labels) and torch.equal(imap, out.image_position_map)
    np.testing.assert_allclose(torch.cat([pc, pr]).numpy(), want.numpy(), rtol=2e-5, atol=2e-4)
