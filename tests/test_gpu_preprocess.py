"""GPU parity of the CLIP image preprocessing kernels (vlb200_clip_preprocess_u8) -- bit-exact against the oracle and
the fixture minted from Pillow + transformers' PIL-backend CLIP processor (the reference collator's code path,
models/Llava/__init__.py:435-443)."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import image_restate as IR

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden", "g7_clip_preprocess.npz")


@pytest.fixture(scope="module")
def P():
    import vlrlhf_b200  # noqa: F401
    from vlrlhf_b200 import preprocess
    return preprocess


def test_preprocess_bit_exact_against_fixture_and_oracle(P):
    d = np.load(G)
    pre = P.ClipPreprocessor()
    imgs = [IR.synthetic_image(h, w, i) for i, (h, w) in enumerate(IR.G7_SIZES)]
    out = pre(imgs)
    assert out.shape == (len(imgs), 3, 336, 336) and out.dtype == torch.float32
    got = out.cpu().numpy()
    for i, img in enumerate(imgs):
        digest = np.frombuffer(hashlib.sha256(np.ascontiguousarray(got[i]).tobytes()).digest(), dtype=np.uint8)
        want = IR.clip_preprocess(img)
        assert np.array_equal(got[i], want), (i, np.abs(got[i] - want).max())
        assert np.array_equal(digest, d[f"sha256_{i}"]), i


def test_preprocess_edge_sizes_and_dtypes(P):
    rs = np.random.RandomState(3)
    pre = P.ClipPreprocessor()
    pre16 = P.ClipPreprocessor(out_dtype=torch.bfloat16)
    for h, w in [(336, 336), (337, 336), (50, 70), (1365, 2048), (2000, 340), (336, 1500)]:
        img = rs.randint(0, 256, (h, w, 3), dtype=np.uint8)
        want = IR.clip_preprocess(img)
        got = pre([img])[0].cpu().numpy()
        assert np.array_equal(got, want), (h, w)
        got16 = pre16([torch.from_numpy(img).cuda()])[0].float().cpu()
        assert torch.equal(got16, torch.from_numpy(want).to(torch.bfloat16).float()), (h, w)
    with pytest.raises(ValueError):
        pre([np.zeros((10, 10), dtype=np.uint8)])
    # a small-size processor (224) on the same kernels
    pre224 = P.ClipPreprocessor(size=224, crop=224)
    img = rs.randint(0, 256, (300, 500, 3), dtype=np.uint8)
    assert np.array_equal(pre224([img])[0].cpu().numpy(), IR.clip_preprocess(img, 224, 224))


def test_preprocessed_pixels_drive_the_engine(P):
    """uint8 images -> GPU preprocess -> train_step: the collator-to-step path with no CPU pixel work."""
    from vlrlhf_b200 import config, engine
    from oracle import restate as R
    cfg = config.ModelConfig(image_size=336, patch_size=14, v_hidden=128, v_layers=3, v_heads=2, v_ff=256, hidden=128,
                             layers=2, heads=2, kv_heads=2, ff=256, vocab=320, image_token_index=300, pad_token_id=301)
    rcfg = R.LlavaCfg(image_size=336, patch_size=14, v_hidden=128, v_layers=3, v_heads=2, v_ff=256, hidden=128, layers=2,
                      heads=2, kv_heads=2, ff=256, vocab=320, image_token_index=300, pad_token_id=301)
    eng = engine.LlavaDPOEngine(cfg, config.TrainConfig(), with_optimizer=False)
    eng.init_synthetic(0)
    batch = R.make_batch(rcfg, 2, 24, 8, seed=0)
    imgs = [IR.synthetic_image(300, 400, 1), IR.synthetic_image(500, 350, 2)]
    batch["img_input_dict"]["pixel_values"] = P.ClipPreprocessor()(imgs)
    got = eng.train_step(batch, train=False)
    batch["img_input_dict"]["pixel_values"] = torch.from_numpy(np.stack([IR.clip_preprocess(i) for i in imgs]))
    wp, wr = R.make_policy_and_ref(rcfg, 0)
    with torch.no_grad():
        loss, metrics, _ = R.get_batch_loss_metrics(rcfg, wp, wr, batch)
    assert abs(got["logps/chosen"] / float(metrics["logps/chosen"]) - 1) < 1e-3
    assert abs(got["logps/rejected"] / float(metrics["logps/rejected"]) - 1) < 1e-3


def test_collator_end_to_end_on_gpu(P, tmp_path):
    """PNG files -> host decode -> uint8 H2D -> GPU preprocess: pixel_values bit-equal to the CPU processor path."""
    Image = pytest.importorskip("PIL.Image")
    from vlrlhf_b200.collator import B200DPODataCollatorWithPadding
    feats = []
    for i, (h, w) in enumerate([(400, 600), (700, 500)]):
        path = str(tmp_path / f"im{i}.png")
        Image.fromarray(IR.synthetic_image(h, w, i)).save(path)
        feats.append({"img_path": path, "chosen_input_ids": [1, 2, 3 + i], "chosen_attention_mask": [1, 1, 1],
                      "chosen_labels": [-100, 2, 3 + i], "rejected_input_ids": [1, 2], "rejected_attention_mask": [1, 1],
                      "rejected_labels": [-100, 2], "prompt_input_ids": [1], "prompt_attention_mask": [1]})
    batch = B200DPODataCollatorWithPadding(pad_token_id=0, preprocessor=P.ClipPreprocessor())(feats)
    pv = batch["img_input_dict"]["pixel_values"]
    assert pv.is_cuda and pv.shape == (2, 3, 336, 336)
    for i, (h, w) in enumerate([(400, 600), (700, 500)]):
        assert np.array_equal(pv[i].cpu().numpy(), IR.clip_preprocess(IR.synthetic_image(h, w, i)))
    assert batch["chosen_input_ids"].shape == (2, 3) and batch["rejected_labels"][0].tolist() == [-100, 2]


def test_square_transform_on_gpu(P):
    """Qwen-VL (448) / InternLM-XC2 (490) transform: Resize((s, s)) + ToTensor + Normalize, bit-exact."""
    for size, (h, w), seed in ((448, (300, 500), 1), (448, (700, 333), 2), (490, (490, 490), 3), (112, (60, 45), 4)):
        img = IR.synthetic_image(h, w, seed)
        got = P.ClipPreprocessor(size=size, square=True)([img])[0].cpu().numpy()
        assert got.shape == (3, size, size)
        assert np.array_equal(got, IR.square_preprocess(img, size)), (size, h, w)


def test_anyres_preprocessor_on_gpu(P):
    """LLaVA-Next image processor on the GPU: base view + zero-padded canvas cells, bit-exact; ragged view counts are
    zero-padded like _pad_for_batching; the crops drive the anyres engine path unchanged."""
    pins = [[336, 672], [672, 336], [672, 672], [1008, 336], [336, 1008]]
    pre = P.AnyresPreprocessor(pins)
    shapes = [(336, 336), (400, 640), (900, 300), (150, 1000), (50, 60), (1200, 1300)]
    imgs = [IR.synthetic_image(h, w, s) for s, (h, w) in enumerate(shapes)]
    out, sizes = pre(imgs)
    assert sizes.tolist() == [list(s) for s in shapes]
    for i, img in enumerate(imgs):
        want = IR.anyres_preprocess(img, pins)
        got = out[i, :want.shape[0]].cpu().numpy()
        assert np.array_equal(got, want), (shapes[i], np.abs(got - want).max())
        assert float(out[i, want.shape[0]:].abs().max()) == 0.0 if want.shape[0] < out.shape[1] else True
    from vlrlhf_b200 import host
    crops = [host.anyres_num_crops(s, pins, 336) for s in shapes]
    assert crops == [IR.anyres_preprocess(im, pins).shape[0] for im in imgs]
