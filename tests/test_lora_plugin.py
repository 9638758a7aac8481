"""Plugin side of the LoRA-trained families (vlrlhf_b200/plugin_lora.py) over the mock ops -- CPU tests: checkpoints in,
PEFT-format adapters out, merge (merge_peft_model.py), LoraConfig validation, launcher flags, trainer-level calls."""
import importlib
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import lora_restate as LR
from oracle import restate as R
from oracle import xc2_restate as X


@pytest.fixture(scope="module")
def lplug():
    import vlrlhf_b200
    from tests import mock_ops
    names = ("vlrlhf_b200.ops", "vlrlhf_b200.engine", "vlrlhf_b200.engine_lora", "vlrlhf_b200.engine_qwen", "vlrlhf_b200.engine_xc2",
             "vlrlhf_b200.plugin", "vlrlhf_b200.plugin_lora")
    saved = {k: sys.modules.get(k) for k in names}
    saved_attr = {k: getattr(vlrlhf_b200, k.split(".")[1], None) for k in names}
    sys.modules["vlrlhf_b200.ops"] = mock_ops
    vlrlhf_b200.ops = mock_ops
    for k in names[1:]:
        sys.modules.pop(k, None)
        if hasattr(vlrlhf_b200, k.split(".")[1]):
            delattr(vlrlhf_b200, k.split(".")[1])
    plugin = importlib.import_module("vlrlhf_b200.plugin")
    pl = importlib.import_module("vlrlhf_b200.plugin_lora")
    from vlrlhf_b200 import config, host
    yield config, plugin, pl, host
    for k, v in saved.items():
        attr = k.split(".")[1]
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
        if saved_attr[k] is None:
            if hasattr(vlrlhf_b200, attr):
                delattr(vlrlhf_b200, attr)
        else:
            setattr(vlrlhf_b200, attr, saved_attr[k])


def test_lora_args_from_launcher_flags(lplug):
    config, plugin, pl, host = lplug
    argv = "--model_name_or_path ckpts/llava --use_lora True --lora_r 128 --lora_alpha 256 --lora_dropout 0.05".split()
    assert pl.lora_args_from_argv(argv) == {"lora_r": 128, "lora_alpha": 256.0}
    assert pl.lora_args_from_argv(["--use_lora=true"]) == {}
    assert pl.lora_args_from_argv(["--use_lora", "False", "--lora_r", "8"]) is None
    assert pl.lora_args_from_argv(["--bf16", "True"]) is None


@pytest.mark.parametrize("family", ["llava", "llava_next"])
def test_llava_lora_from_pretrained_adapters_merge_and_trainer_calls(lplug, tmp_path, family):
    transformers = pytest.importorskip("transformers")
    pytest.importorskip("safetensors")
    from oracle import make_fixtures as MF
    from vlrlhf_b200 import checkpoint
    config, plugin, pl, host = lplug
    lcfg = LR.TINY_LORA if family == "llava" else LR.TINY_NEXT_LORA
    w, lora = LR.make_weights(lcfg, 0)
    # a base checkpoint on disk in HF layout (written by transformers itself), holding the oracle's base weights
    hf = (transformers.LlavaForConditionalGeneration(MF.hf_config(lcfg)) if family == "llava" else
          transformers.LlavaNextForConditionalGeneration(MF.hf_config_next(lcfg))).to(torch.bfloat16)
    sd = hf.state_dict()
    with torch.no_grad():
        for k, v in w.items():
            sd[MF.hf_name(k)].copy_(v.reshape(sd[MF.hf_name(k)].shape))
    src = str(tmp_path / "base")
    hf.save_pretrained(src, safe_serialization=True)
    model = pl.B200LlavaLoRAForRL.from_pretrained(src, torch_dtype=torch.bfloat16, device="cpu", lora_r=lcfg.lora_r,
                                                  lora_alpha=lcfg.lora_alpha)
    eng = model.engine
    assert model.cfg.lora_r == lcfg.lora_r and model.cfg.family == lcfg.family and model.cfg.kv_heads == lcfg.kv_heads
    st = eng.hf_state("ref")
    for k, v in w.items():
        if k in st:
            assert torch.equal(st[k].float().reshape(v.shape), v), k
    trainable = {n for n, p in model._hf.items() if p.requires_grad}
    assert trainable == set(lora) and all(model._hf[n].grad is not None for n in trainable)
    assert sorted(model.default_lora_target) == sorted(["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"])
    # peft init (B = 0): policy == reference; then the synthetic adapters through the PEFT file format
    sizes = [(28, 28), (20, 50)] if family == "llava_next" else None
    batch = R.make_batch(lcfg, 2, 24, 8, 1, ddpo_like=True, image_sizes=sizes)
    trainer = SimpleNamespace(loss_type="sigmoid", is_encoder_decoder=False, label_pad_token_id=-100, padding_value=0)
    with torch.no_grad():
        pc, pr, _, _ = plugin.concatenated_forward(trainer, model, batch)
        rc, rr, _, _ = plugin.concatenated_forward(trainer, plugin.RefView(model), batch)
    assert torch.allclose(torch.cat([pc, pr]), torch.cat([rc, rr]), rtol=0, atol=1e-3)
    views = eng.lora_views(eng.policy)
    for k, v in lora.items():
        views[k].copy_(v.to(torch.bfloat16))
    files = model.save_pretrained(str(tmp_path / "adapter"))
    assert files == ["adapter_model.safetensors", "adapter_config.json"]
    acfg = json.loads((tmp_path / "adapter" / "adapter_config.json").read_text())
    assert acfg["r"] == lcfg.lora_r and acfg["lora_alpha"] == lcfg.lora_alpha and acfg["base_model_name_or_path"] == src
    from safetensors import safe_open
    with safe_open(str(tmp_path / "adapter" / "adapter_model.safetensors"), "pt") as f:
        assert set(f.keys()) == {f"base_model.model.{k}.weight" for k in lora}
    model.reset_adapters()
    assert float(views["language_model.model.layers.0.self_attn.k_proj.lora_B"].abs().max()) == 0.0
    model.load_adapter(str(tmp_path / "adapter"))
    for k, v in lora.items():
        assert torch.equal(views[k].float(), v), k
    # trainer-level calls with the adapters on / off against the oracle
    with torch.no_grad():
        pc, pr, _, _ = plugin.concatenated_forward(trainer, model, batch)
        rc, rr, _, _ = plugin.concatenated_forward(trainer, plugin.RefView(model), batch)
        w_pc, w_pr, *_ = LR.concatenated_forward(lcfg, w, lora, batch)
        w_rc, w_rr, *_ = LR.concatenated_forward(lcfg, w, None, batch)
    np.testing.assert_allclose(torch.cat([pc, pr]).numpy(), torch.cat([w_pc, w_pr]).numpy(), rtol=1e-3)
    np.testing.assert_allclose(torch.cat([rc, rr]).numpy(), torch.cat([w_rc, w_rr]).numpy(), rtol=1e-3)
    # autograd through the plugin: gradients land in the adapters' .grad views
    trainer.beta, trainer.label_smoothing, trainer.reference_free = 0.1, 0.0, False
    eng.grads.zero_()
    pc, pr, _, _ = plugin.concatenated_forward(trainer, model, batch)
    losses, _, _ = plugin.dpo_loss(trainer, pc, pr, rc, rr)
    losses.mean().backward()
    assert float(model._hf["language_model.model.layers.1.mlp.down_proj.lora_B"].grad.abs().max()) > 0
    # merge_peft_model.py: adapters folded into the base; the merged checkpoint reloads into HF and matches base + adapters
    merged = model.merged_state()
    k = "language_model.model.layers.0.self_attn.q_proj"
    want = (w[k + ".weight"] + lcfg.lora_scale * lora[k + ".lora_B"] @ lora[k + ".lora_A"]).to(torch.bfloat16)
    assert torch.equal(merged[k + ".weight"], want)
    out = str(tmp_path / "merged")
    model.save_merged(out)
    cls = transformers.LlavaForConditionalGeneration if family == "llava" else transformers.LlavaNextForConditionalGeneration
    back, info = cls.from_pretrained(out, torch_dtype=torch.bfloat16, output_loading_info=True)
    assert not info["missing_keys"] and not info["unexpected_keys"], info
    got = {checkpoint.legacy_name(n): v for n, v in back.state_dict().items()}
    assert torch.equal(got[k + ".weight"], want)
    # LoraConfig validation of the trainer hook
    ok = SimpleNamespace(r=lcfg.lora_r, lora_alpha=lcfg.lora_alpha, target_modules=list(model.default_lora_target))
    pl.check_peft_config(model, ok)
    for bad in (None, SimpleNamespace(r=4, lora_alpha=lcfg.lora_alpha, target_modules=ok.target_modules),
                SimpleNamespace(r=lcfg.lora_r, lora_alpha=lcfg.lora_alpha, target_modules=["q_proj"])):
        with pytest.raises(ValueError):
            pl.check_peft_config(model, bad)


def test_xc2_from_pretrained_adapters_and_trainer_calls(lplug, tmp_path):
    pytest.importorskip("safetensors")
    from safetensors.torch import save_file
    config, plugin, pl, host = lplug
    xcfg = X.TINY_XC2
    w, lora = X.make_weights(xcfg, 0)
    src = tmp_path / "xc2"
    src.mkdir()
    save_file({k: v.to(torch.bfloat16).contiguous() for k, v in w.items()}, str(src / "model.safetensors"))
    hf_cfg = dict(model_type="internlmxcomposer2", vocab_size=xcfg.vocab, hidden_size=xcfg.hidden, num_hidden_layers=xcfg.layers,
                  num_attention_heads=xcfg.heads, num_key_value_heads=xcfg.kv_heads, intermediate_size=xcfg.ff,
                  rms_norm_eps=xcfg.rms_eps, rope_theta=xcfg.rope_theta, bias=False, img_size=xcfg.image_size, max_length=4096,
                  image_token_index=xcfg.image_token_index, pad_token_id=xcfg.pad_token_id)
    (src / "config.json").write_text(json.dumps(hf_cfg))
    tiny = config.TINY_XC2
    got = pl.xc2_config_from_hf(hf_cfg)
    assert got.plora_r == 256 and got.v_hidden == 1024 and got.lora_r == 64   # the reference's hard-coded tower / ranks
    # the tiny checkpoint sizes its tower through the optional vision_config section
    hf_cfg.update(vision_config=dict(hidden_size=tiny.v_hidden, num_hidden_layers=tiny.v_layers, num_attention_heads=tiny.v_heads,
                                     intermediate_size=tiny.v_ff, patch_size=tiny.patch_size),
                  plora_r=tiny.plora_r, plora_alpha=tiny.plora_alpha)
    (src / "config.json").write_text(json.dumps(hf_cfg))
    model = pl.B200InternLMXC2ForRL.from_pretrained(str(src), torch_dtype=torch.bfloat16, device="cpu", lora_r=tiny.lora_r,
                                                    lora_alpha=tiny.lora_alpha)
    for f in ("hidden", "layers", "heads", "kv_heads", "ff", "vocab", "image_size", "image_token_index", "pad_token_id", "lora_r",
              "rope_theta", "rms_eps", "vision_feature_layer", "family", "v_hidden", "v_layers", "plora_r"):
        assert getattr(model.cfg, f) == getattr(tiny, f), f
    st = model.engine.hf_state("ref")
    for k, v in w.items():
        if not k.startswith("vit."):
            assert torch.equal(st[k].float().reshape(v.shape), v), k
    with pytest.raises(KeyError):   # an incomplete checkpoint is refused
        bad = tmp_path / "bad"
        bad.mkdir()
        save_file({k: v.to(torch.bfloat16).contiguous() for k, v in list(w.items())[:-3]}, str(bad / "model.safetensors"))
        (bad / "config.json").write_text(json.dumps(hf_cfg))
        pl.B200InternLMXC2ForRL.from_pretrained(str(bad), device="cpu", lora_r=tiny.lora_r, lora_alpha=tiny.lora_alpha)
    eng = model.engine
    trainable = {n for n, p in model._hf.items() if p.requires_grad}
    assert trainable == set(lora)
    assert sorted(model.default_lora_target) == sorted(X.LINEARS)
    batch = R.make_batch(xcfg, 2, 24, 8, 1, ddpo_like=True)
    trainer = SimpleNamespace(loss_type="kto_pair", is_encoder_decoder=False, label_pad_token_id=-100, padding_value=0)
    with torch.no_grad():
        pc, pr, _, _ = plugin.concatenated_forward(trainer, model, batch)
        rc, rr, _, _ = plugin.concatenated_forward(trainer, plugin.RefView(model), batch)
    assert torch.allclose(torch.cat([pc, pr]), torch.cat([rc, rr]), rtol=0, atol=1e-3)
    # adapters through the PEFT file format (reference row order on disk, engine row order in memory)
    for k, v in lora.items():
        model._write_adapter(k, v)
    model.save_pretrained(str(tmp_path / "adapter"))
    from safetensors import safe_open
    with safe_open(str(tmp_path / "adapter" / "adapter_model.safetensors"), "pt") as f:
        assert set(f.keys()) == {f"base_model.model.{k}.weight" for k in lora}
        k0 = "model.layers.1.attention.wqkv.lora_B"
        assert torch.equal(f.get_tensor(f"base_model.model.{k0}.weight").float(), lora[k0])
    model.reset_adapters()
    model.load_adapter(str(tmp_path / "adapter"))
    st = eng.hf_state("policy")
    for k, v in lora.items():
        assert torch.equal(st[k].float(), v), k
    with torch.no_grad():
        pc, pr, _, _ = plugin.concatenated_forward(trainer, model, batch)
        w_pc, w_pr, *_ = X.concatenated_forward(xcfg, w, lora, batch)
    np.testing.assert_allclose(torch.cat([pc, pr]).numpy(), torch.cat([w_pc, w_pr]).numpy(), rtol=1e-3)
    merged = model.merged_state()
    k = "model.layers.0.attention.wqkv"
    want = (w[k + ".weight"] + xcfg.lora_scale * lora[k + ".lora_B"] @ lora[k + ".lora_A"]).to(torch.bfloat16)
    assert torch.equal(merged[k + ".weight"], want)
