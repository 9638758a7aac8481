"""Drop-in model wrappers for the LoRA-trained families behind the reference's plugin API:

  * `B200LlavaLoRAForRL`   LLaVA-1.5 / LLaVA-Next with peft LoRA on the decoder linears -- what scripts/dpo_llava.sh,
                           dpo_llavanext.sh, kto_*.sh, ddpo_*.sh launch (`--use_lora True --lora_r 128 --lora_alpha 256
                           --lora_target_modules auto`, utils/auto_load.py:559-578) -- on engine_lora.LlavaLoRADPOEngine;
  * `B200InternLMXC2ForRL` InternLM-XComposer2-VL (models/InternLMXC2/__init__.py:106-299; scripts/dpo_internlmxc2vl7b.sh:
                           LoRA r 64 alpha 64 on attention.wqkv/wo, feed_forward.w1/w2/w3) on engine_xc2.XC2DPOEngine.

Both mirror the model-side contract of docs/CustomizedModel.md (default_lora_target, get_vision_tower,
freeze_vision_tower, prepare_default_generation_kwargs), `from_pretrained(dir, config=…, torch_dtype=…)` as
utils/auto_load.py:522-535 calls it, and `save_pretrained(dir)` = a PEFT-format adapter (`adapter_model.safetensors` +
`adapter_config.json`) so merge_peft_model.py:7-24 and the eval harness load the result (dpo.py:89-95).  The adapters are
the engine's tensors (not peft modules): the trainer subclasses built by `install_lora()` check the launcher's LoraConfig
against what was allocated and hand `peft_config=None` on, and the reference pass is the same engine with the adapters
off (`plugin.RefView`), which is what TRL's `null_ref_context()` does for a peft policy.
"""
from __future__ import annotations

import json
import math
import os
import sys
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .config import LLAVA_LORA_LINEARS, TrainConfig, XC2ModelConfig, with_lora
from .plugin import B200ModuleMixin


class _B200LoRAModel(B200ModuleMixin, nn.Module):
    """nn.Module whose parameters are views of an adapter-training engine's arenas: adapters trainable with `.grad` = the
    engine's flat gradient buffer, base and tower frozen (peft freezes every base parameter)."""

    def __init__(self, engine, cfg):
        super().__init__()
        self.engine, self.cfg = engine, cfg
        grads = self._storage(engine.g)
        self._register_engine_params({**self._base_storage(), **self._storage(engine.policy)}, grads, lambda n: n in grads)
        self.hf_config_dict: Optional[dict] = None
        self.base_model_name_or_path: Optional[str] = None
        self.config = None

    # storage-layout views (a family may store rows permuted; an elementwise optimizer does not care)
    def _storage(self, w) -> Dict[str, torch.Tensor]:
        eng = self.engine
        return eng._lora_storage(w) if hasattr(eng, "_lora_storage") else eng.lora_views(w)

    def _base_storage(self) -> Dict[str, torch.Tensor]:
        eng = self.engine
        return eng._base_storage() if hasattr(eng, "_base_storage") else eng.base_views()

    def _write_adapter(self, name: str, t: torch.Tensor):
        eng = self.engine
        dst = self._storage(eng.policy)
        if name not in dst:
            raise KeyError(f"adapter tensor {name} has no counterpart (targets: {self.default_lora_target})")
        if hasattr(eng, "_store"):
            eng._store(dst, name, t)
        else:
            dst[name].copy_(t.to(eng.device, torch.bfloat16).reshape(dst[name].shape))

    def forward(self, *a, **k):
        raise RuntimeError(f"{type(self).__name__} is driven through concatenated_forward / engine.train_step; generation "
                           "and evaluation forwards are outside the hot path this package replaces")

    # ---- adapters
    def reset_adapters(self, seed: int = 0):
        """peft's LoRA init: A ~ kaiming_uniform(a=sqrt(5)) = U(-1/sqrt(in), 1/sqrt(in)), B = 0 (policy == reference)."""
        eng = self.engine
        eng.wait_optimizer()
        gen = torch.Generator().manual_seed(seed)
        for name, t in self._storage(eng.policy).items():
            if name.endswith("lora_A"):
                bound = 1.0 / math.sqrt(t.shape[1])
                t.copy_(((torch.rand(t.shape, generator=gen) * 2 - 1) * bound).to(t.device, torch.bfloat16))
            else:
                t.zero_()   # (row order is irrelevant for zeros)
        eng.sync_master_from_params()

    def adapter_state(self) -> Dict[str, torch.Tensor]:
        """PEFT checkpoint names -> reference-layout adapter tensors."""
        eng = self.engine
        return {f"base_model.model.{k}.weight": v for k, v in eng.lora_views(eng.policy).items()}

    def save_pretrained(self, save_directory: str, **kwargs):
        """PEFT-format adapter checkpoint of the trained LoRA weights."""
        from safetensors.torch import save_file
        os.makedirs(save_directory, exist_ok=True)
        self.engine.wait_optimizer()
        state = {k: v.detach().to("cpu").contiguous() for k, v in self.adapter_state().items()}
        save_file(state, os.path.join(save_directory, "adapter_model.safetensors"), metadata={"format": "pt"})
        cfg = self.cfg
        with open(os.path.join(save_directory, "adapter_config.json"), "w") as f:
            json.dump({"peft_type": "LORA", "task_type": "CAUSAL_LM", "r": cfg.lora_r, "lora_alpha": cfg.lora_alpha,
                       "lora_dropout": 0.05, "bias": "none", "target_modules": self.default_lora_target,
                       "modules_to_save": None, "fan_in_fan_out": False, "inference_mode": True,
                       "base_model_name_or_path": self.base_model_name_or_path}, f, indent=2)
        return ["adapter_model.safetensors", "adapter_config.json"]

    def load_adapter(self, directory: str):
        from safetensors import safe_open
        with safe_open(os.path.join(directory, "adapter_model.safetensors"), framework="pt", device="cpu") as f:
            for k in f.keys():
                name = k[len("base_model.model."):] if k.startswith("base_model.model.") else k
                name = name[:-len(".weight")] if name.endswith(".weight") else name
                name = name.replace(".lora_A.default", ".lora_A").replace(".lora_B.default", ".lora_B")
                self._write_adapter(name, f.get_tensor(k))
        self.engine.sync_master_from_params()

    def merged_state(self) -> Dict[str, torch.Tensor]:
        """merge_peft_model.py:7-24 (`model.merge_and_unload()`): W' = W + (alpha / r) B A per adapted linear, bf16.
        Export utility (runs once after training), not part of the step."""
        eng = self.engine
        out = {k: v for k, v in eng.hf_state("ref").items()}
        lora = eng.lora_views(eng.policy)
        for k in [k for k in lora if k.endswith(".lora_A")]:
            mod = k[:-len(".lora_A")]
            A, B = lora[k].float(), lora[mod + ".lora_B"].float()
            out[mod + ".weight"] = (out[mod + ".weight"].float() + self.cfg.lora_scale * (B @ A)).to(torch.bfloat16)
        return out


# ------------------------------------------------------------------------------------------------
# LLaVA-1.5 / LLaVA-Next + LoRA
# ------------------------------------------------------------------------------------------------
class B200LlavaLoRAForRL(_B200LoRAModel):
    def __init__(self, cfg, train: Optional[TrainConfig] = None, device: str = "cuda", with_optimizer: bool = True):
        from .engine_lora import LlavaLoRADPOEngine
        super().__init__(LlavaLoRADPOEngine(cfg, train, device=device, with_optimizer=with_optimizer), cfg)

    @property
    def default_lora_target(self) -> List[str]:  # Llava/__init__.py:273-286, LlavaNext/__init__.py:347-360
        return [n.split(".")[1] for n in LLAVA_LORA_LINEARS]

    def get_vision_tower(self):
        return {k: v for k, v in self._hf.items() if k.startswith("vision_tower.")}

    def freeze_vision_tower(self):
        pass  # frozen by construction (peft freezes every base parameter; --freeze_vision_tower True)

    def prepare_default_generation_kwargs(self, generation_config):  # Llava/__init__.py:294-298
        generation_config.max_new_tokens = 1024
        generation_config.do_sample = False
        return dict(generation_config=generation_config)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, *args, config=None, torch_dtype=None,
                        train: Optional[TrainConfig] = None, device: str = "cuda", with_optimizer: bool = True,
                        lora_r: int = 128, lora_alpha: float = 256.0, lora_seed: int = 0, **kwargs):
        """A local HF LLaVA / LLaVA-Next checkpoint directory -> frozen base arena (+ tower); adapters start as peft
        starts them, or are read from `adapter_model.safetensors` when the directory holds one."""
        from . import checkpoint
        if torch_dtype not in (None, torch.bfloat16, "bfloat16", "auto"):
            raise ValueError(f"torch_dtype {torch_dtype}: the B200 path computes in bf16 only")
        path = pretrained_model_name_or_path
        if not os.path.isdir(path):
            raise FileNotFoundError(f"{path}: a local checkpoint directory is required (there is no hub access)")
        with open(os.path.join(path, "config.json")) as f:
            cfg_dict = json.load(f)
        cfg = with_lora(checkpoint.config_from_hf(config if config is not None else cfg_dict), int(lora_r), float(lora_alpha))
        model = cls(cfg, train, device=device, with_optimizer=with_optimizer)
        model.hf_config_dict, model.base_model_name_or_path, model.config = cfg_dict, path, config
        checkpoint.load_hf_checkpoint(model.engine, path)
        model.reset_adapters(lora_seed)
        if os.path.exists(os.path.join(path, "adapter_model.safetensors")):
            model.load_adapter(path)
        return model

    def save_merged(self, save_directory: str, max_shard_size: int = 5 << 30):
        """Full HF checkpoint with the adapters folded into the decoder weights (what merge_peft_model.py writes)."""
        from . import checkpoint
        eng = self.engine
        merged = self.merged_state()

        class _View:  # save_hf_checkpoint reads cfg, hf_state(which) and extra_state
            cfg, extra_state = eng.cfg, getattr(eng, "extra_state", {})

            @staticmethod
            def hf_state(which):
                return merged

        return checkpoint.save_hf_checkpoint(_View, save_directory, self.hf_config_dict or checkpoint.hf_config_dict(eng.cfg),
                                             max_shard_bytes=int(max_shard_size))


# ------------------------------------------------------------------------------------------------
# InternLM-XComposer2-VL
# ------------------------------------------------------------------------------------------------
def xc2_config_from_hf(hf_config, lora_r: int = 64, lora_alpha: float = 64.0) -> XC2ModelConfig:
    """InternLMXcomposer2Config (configuration_internlm_xcomposer2.py; object or parsed config.json) -> XC2ModelConfig.
    The tower is the constructor's hard-coded CLIP-L/14 resized to `img_size` (build_mlp.py:9-11,37-137), the partial-LoRA
    rank 256 / alpha 256 is hard-coded in modeling_internlm2.py:215-217,259-272."""
    g = (lambda k, d=None: hf_config.get(k, d)) if isinstance(hf_config, dict) else (lambda k, d=None: getattr(hf_config, k, d))
    hidden, heads = g("hidden_size", 4096), g("num_attention_heads", 32)
    if g("bias", False):
        raise ValueError("InternLM2 linears with bias are not supported (internlm-xcomposer2-vl-7b sets bias=false)")
    if g("rope_scaling") not in (None, {}) and (g("rope_scaling") or {}).get("type", (g("rope_scaling") or {}).get("rope_type")) \
            not in (None, "default"):
        raise ValueError("rope_scaling is not supported")
    image_token = g("image_token_index")
    if image_token is None:
        raise ValueError("config.image_token_index is required (models/InternLMXC2/__init__.py:37 reads it)")
    # non-standard, optional: a `vision_config` / `plora_r` section sizes a scaled-down tower (test checkpoints); the
    # reference model itself always builds CLIP-L/14 and rank-256 partial LoRA
    v = g("vision_config") or {}
    tower = dict(v_hidden=int(v.get("hidden_size", 1024)), v_layers=int(v.get("num_hidden_layers", 24)),
                 v_heads=int(v.get("num_attention_heads", 16)), v_ff=int(v.get("intermediate_size", 4096)),
                 patch_size=int(v.get("patch_size", 14)), plora_r=int(g("plora_r", 256)), plora_alpha=float(g("plora_alpha", 256.0)))
    return XC2ModelConfig(**tower, image_size=int(g("img_size", 490)), vision_feature_layer=-1, hidden=hidden,
                          layers=g("num_hidden_layers", 32), heads=heads, kv_heads=g("num_key_value_heads", heads) or heads,
                          ff=g("intermediate_size", 14336), vocab=g("vocab_size", 92544), rms_eps=g("rms_norm_eps", 1e-5),
                          rope_theta=float(g("rope_theta", 1e6)), image_token_index=int(image_token),
                          pad_token_id=int(g("pad_token_id", 2)), family="xc2", lora_r=int(lora_r), lora_alpha=float(lora_alpha),
                          max_positions=max(4096, min(int(g("max_position_embeddings", 4096)), 32768)))


class B200InternLMXC2ForRL(_B200LoRAModel):
    def __init__(self, cfg: XC2ModelConfig, train: Optional[TrainConfig] = None, device: str = "cuda",
                 with_optimizer: bool = True):
        from .engine_xc2 import XC2DPOEngine
        super().__init__(XC2DPOEngine(cfg, train, device=device, with_optimizer=with_optimizer), cfg)

    @property
    def default_lora_target(self) -> List[str]:  # InternLMXC2/__init__.py:249-251
        return ["attention.wqkv", "attention.wo", "feed_forward.w1", "feed_forward.w2", "feed_forward.w3"]

    def get_vision_tower(self):
        return self.engine.vparams

    def freeze_vision_tower(self):
        pass  # tower + vision_proj are frozen by construction (InternLMXC2/__init__.py:256-259)

    def prepare_default_generation_kwargs(self, generation_config):  # InternLMXC2/__init__.py:261-282 (stop words aside)
        generation_config.do_sample = False
        generation_config.eos_token_id = 2
        return dict(generation_config=generation_config)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, *args, config=None, torch_dtype=None,
                        train: Optional[TrainConfig] = None, device: str = "cuda", with_optimizer: bool = True,
                        lora_r: int = 64, lora_alpha: float = 64.0, lora_seed: int = 0, **kwargs):
        from . import checkpoint
        if torch_dtype not in (None, torch.bfloat16, "bfloat16", "auto"):
            raise ValueError(f"torch_dtype {torch_dtype}: the B200 path computes in bf16 only")
        path = pretrained_model_name_or_path
        if not os.path.isdir(path):
            raise FileNotFoundError(f"{path}: a local checkpoint directory is required (there is no hub access)")
        with open(os.path.join(path, "config.json")) as f:
            cfg_dict = json.load(f)
        model = cls(xc2_config_from_hf(config if config is not None else cfg_dict, lora_r, lora_alpha), train, device=device,
                    with_optimizer=with_optimizer)
        model.hf_config_dict, model.base_model_name_or_path, model.config = cfg_dict, path, config
        eng = model.engine
        need = set(eng._base_storage()) | set(eng._vision_names())
        seen = set()

        def stream():
            for name, t in checkpoint.iter_checkpoint(path):
                seen.add(name)
                yield name, t

        eng.load_state_dict_tensors(stream())
        missing = sorted(need - seen)
        if missing:
            raise KeyError(f"checkpoint lacks {len(missing)} tensors, e.g. {missing[:4]}")
        model.reset_adapters(lora_seed)
        if os.path.exists(os.path.join(path, "adapter_model.safetensors")):
            model.load_adapter(path)
        return model


# ------------------------------------------------------------------------------------------------
# registration
# ------------------------------------------------------------------------------------------------
def check_peft_config(model: _B200LoRAModel, peft_config) -> None:
    """The launcher hands the LoraConfig of utils/auto_load.py:559-571 to TRL; here the adapters live in the engine, so the
    config must describe exactly what was allocated."""
    if peft_config is None:
        raise ValueError(f"{type(model).__name__} trains LoRA adapters only: run with --use_lora True")
    r, alpha = getattr(peft_config, "r"), getattr(peft_config, "lora_alpha")
    targets = sorted(getattr(peft_config, "target_modules"))
    if r != model.cfg.lora_r or float(alpha) != float(model.cfg.lora_alpha):
        raise ValueError(f"LoraConfig(r={r}, lora_alpha={alpha}) != engine adapters (r={model.cfg.lora_r}, "
                         f"alpha={model.cfg.lora_alpha}); pass lora_r / lora_alpha to from_pretrained")
    if targets != sorted(model.default_lora_target):
        raise ValueError(f"lora_target_modules {targets}: this path adapts exactly {sorted(model.default_lora_target)}")


def lora_args_from_argv(argv: Optional[List[str]] = None) -> Optional[Dict[str, float]]:
    """`--use_lora True [--lora_r R] [--lora_alpha A]` as the reference's launch scripts spell them (HfArgumentParser
    booleans: True/true/1/yes) -> {"lora_r", "lora_alpha"} or None when LoRA is off (dpo.py:79 default)."""
    argv = list(sys.argv[1:] if argv is None else argv)

    def value(flag):
        for i, a in enumerate(argv):
            if a == flag and i + 1 < len(argv):
                return argv[i + 1]
            if a.startswith(flag + "="):
                return a.split("=", 1)[1]
        return None

    on = value("--use_lora")
    if on is None or on.lower() not in ("true", "1", "yes", "y", "t"):
        return None
    out = {}
    if value("--lora_r") is not None:
        out["lora_r"] = int(value("--lora_r"))
    if value("--lora_alpha") is not None:
        out["lora_alpha"] = float(value("--lora_alpha"))
    return out


def _trainer_class(base, plugin):
    return plugin.make_trainer_class(base, check_peft=check_peft_config, name="B200LoRADPOTrainer")


def install_lora(families=("Llava", "LlavaNext", "InternLMXC2"), lora_args: Optional[Dict[str, float]] = None):
    """`vlrlhf.models.<family>.core_mapper` -> (LoRA model wrapper, reference processor/collators, B200 DPO trainer), so
    `src/vlrlhf/dpo.py --use_lora True ...` (every scripts/*.sh) runs unmodified.  `lora_args` (default: read from the
    command line like the launcher will) become the defaults of `from_pretrained`, which the reference calls without
    any LoRA argument (utils/auto_load.py:522-535)."""
    import importlib
    import functools
    from . import plugin
    from vlrlhf.models.utils import ModelCoreMapper
    args = lora_args_from_argv() if lora_args is None else lora_args
    mappers = {}
    for fam in families:
        try:
            mod = importlib.import_module(f"vlrlhf.models.{fam}")
        except Exception:  # a family whose vendored code needs symbols the installed transformers lacks
            continue
        ref = mod.core_mapper
        cls = B200InternLMXC2ForRL if fam == "InternLMXC2" else B200LlavaLoRAForRL
        if args:
            bound = type(cls.__name__, (cls,), {})
            bound.from_pretrained = classmethod(functools.partial(cls.from_pretrained.__func__, **args))
            cls = bound
        mod.core_mapper = ModelCoreMapper(
            model=cls, processor=ref.processor, dpo_collator=ref.dpo_collator, dpo_trainer=_trainer_class(ref.dpo_trainer, plugin),
            reward_model=ref.reward_model, value_model=ref.value_model, reward_collator=ref.reward_collator,
            reward_trainer=ref.reward_trainer, sft_collator=ref.sft_collator, sft_trainer=ref.sft_trainer,
            ppo_collator=ref.ppo_collator, ppo_trainer=ref.ppo_trainer)
        mappers[fam] = mod.core_mapper
    return mappers
