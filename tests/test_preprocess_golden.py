"""Image side of the DPO collator -- CPU tests.

* the oracle (oracle/image_restate.py: Pillow's 8-bit bicubic resampler + transformers' crop/rescale/normalize
  restated in numpy) against the committed fixture minted from Pillow + transformers' PIL-backend CLIP processor
  (tests/golden/g7_clip_preprocess.npz) and, when Pillow is importable, against Pillow itself on fresh sizes;
* the product's host tables (vlrlhf_b200/preprocess.py, vectorised) == the oracle's loop restatement, bit-exact.
The CUDA kernels themselves are checked in tests/test_gpu_preprocess.py.
"""
import hashlib
import os

import numpy as np
import pytest

from oracle import image_restate as IR

G = os.path.join(os.path.dirname(__file__), "golden", "g7_clip_preprocess.npz")


def test_oracle_matches_golden_fixture():
    d = np.load(G)
    sizes = [tuple(x) for x in d["sizes"].tolist()]
    assert sizes == IR.G7_SIZES
    for i, (h, w) in enumerate(sizes):
        img = IR.synthetic_image(h, w, i)
        pv = IR.clip_preprocess(img)
        assert pv.dtype == np.float32 and pv.shape == (3, 336, 336)
        digest = np.frombuffer(hashlib.sha256(np.ascontiguousarray(pv).tobytes()).digest(), dtype=np.uint8)
        assert np.array_equal(digest, d[f"sha256_{i}"]), (i, h, w)  # bit-exact float32 output
        if f"pixel_values_{i}" in d.files:
            assert np.array_equal(pv, d[f"pixel_values_{i}"])
            nh, nw = IR.shortest_edge_size(h, w, 336)
            assert np.array_equal(IR.pil_bicubic_resize_u8(img, (nh, nw)), d[f"resized_u8_{i}"])


def test_oracle_matches_pillow_on_fresh_sizes():
    Image = pytest.importorskip("PIL.Image")
    rs = np.random.RandomState(7)
    for (h, w), (oh, ow) in [((123, 457), (336, 1248)), ((900, 350), (864, 336)), ((64, 64), (336, 336)),
                             ((336, 500), (336, 500)), ((1200, 1600), (336, 448)), ((10, 3000), (5, 7))]:
        img = rs.randint(0, 256, (h, w, 3), dtype=np.uint8)
        want = np.array(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
        assert np.array_equal(IR.pil_bicubic_resize_u8(img, (oh, ow)), want), ((h, w), (oh, ow))


def test_product_host_tables_equal_oracle():
    import vlrlhf_b200  # noqa: F401
    from vlrlhf_b200 import preprocess as P
    for i, o in [(640, 448), (480, 336), (336, 336), (300, 336), (1000, 1120), (2048, 504), (70, 470), (337, 337), (50, 336),
                 (3, 7), (4000, 336), (97, 336), (211, 730)]:
        ks, b, kk = IR.precompute_coeffs(i, 0.0, float(i), o)
        k2, b2, c2 = P.resample_tables(i, o)
        assert ks == k2 and np.array_equal(b, b2) and np.array_equal(IR.normalize_coeffs_8bpc(kk), c2), (i, o)
    for hw in [(480, 640), (640, 480), (336, 336), (1000, 300), (200, 333), (1365, 2048), (337, 336), (97, 211)]:
        nh, nw = IR.shortest_edge_size(*hw, 336)
        g = P.resize_geometry(*hw, 336, 336)
        assert (nh, nw) == g[:2] and g[2] == (nh - 336) // 2 and g[3] == (nw - 336) // 2
    with pytest.raises(ValueError):
        P.resize_geometry(100, 100, 224, 336)


def test_square_transform_matches_torchvision():
    """Qwen-VL / XC2 transform: the oracle == torchvision Resize((s, s), BICUBIC) + ToTensor + Normalize, bit for bit."""
    tv = pytest.importorskip("torchvision.transforms")
    Image = pytest.importorskip("PIL.Image")
    from torchvision.transforms import InterpolationMode
    for size, (h, w), seed in ((448, (300, 500), 1), (448, (700, 333), 2), (490, (490, 490), 3), (112, (60, 45), 4)):
        img = IR.synthetic_image(h, w, seed)
        t = tv.Compose([tv.Resize((size, size), interpolation=InterpolationMode.BICUBIC), tv.ToTensor(),
                        tv.Normalize(IR.OPENAI_CLIP_MEAN, IR.OPENAI_CLIP_STD)])
        want = t(Image.fromarray(img)).numpy()
        assert np.array_equal(IR.square_preprocess(img, size), want), (size, h, w)


def test_anyres_oracle_matches_hf_llava_next_processor():
    """oracle.anyres_preprocess == transformers' PIL-backend LlavaNextImageProcessor (the 4.41 slow processor's code path)
    on square / wide / tall / tiny / huge images, bit for bit; image_sizes pass through."""
    Image = pytest.importorskip("PIL.Image")
    mod = pytest.importorskip("transformers.models.llava_next.image_processing_pil_llava_next")
    pins = [[336, 672], [672, 336], [672, 672], [1008, 336], [336, 1008]]
    proc = mod.LlavaNextImageProcessorPil(size={"shortest_edge": 336}, crop_size={"height": 336, "width": 336},
                                          image_grid_pinpoints=pins)
    for seed, (h, w) in enumerate([(336, 336), (400, 640), (900, 300), (150, 1000), (50, 60), (1200, 1300), (672, 671)]):
        img = IR.synthetic_image(h, w, seed)
        o = proc(images=[Image.fromarray(img)], return_tensors="np")
        want = o["pixel_values"][0]
        got = IR.anyres_preprocess(img, pins)
        assert got.shape == want.shape, (h, w, got.shape, want.shape)
        assert np.array_equal(got, want), (h, w, np.abs(got - want).max())
        assert tuple(o["image_sizes"][0]) == (h, w)
