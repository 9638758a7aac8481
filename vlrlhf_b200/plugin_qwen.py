"""Drop-in pieces for the Qwen-VL family (BASELINE.json configs[2]) behind the reference's plugin API.

`vlrlhf.models.QwenVL.core_mapper` (models/QwenVL/__init__.py:359-372) names a model class, a processor, collators and
`QwenVLDPOTrainer`; `install_qwen()` swaps in `B200QwenVLForRL` and a trainer that overrides the same three
`VLDPOTrainer` methods as the LLaVA plugin (plugin.py).  What the model wrapper mirrors:
  * `QwenVLForRL` (:24-46): default_lora_target, get_vision_tower, freeze_vision_tower, prepare_default_generation_kwargs;
  * `from_pretrained(dir, config=…, torch_dtype=…)` as utils/auto_load.py:522-535 calls it: Qwen-VL checkpoint directory
    (config.json of QWenConfig + safetensors with `transformer.*` / `lm_head.weight` names) streamed into the frozen base
    arena and the re-laid-out vision tower; adapters start as peft does (A ~ kaiming-uniform, B = 0);
  * `save_pretrained(dir)`: a PEFT-format adapter (`adapter_model.safetensors` + `adapter_config.json`) so
    merge_peft_model.py and the eval harness load the result (dpo.py:89-95);
  * the image path of `QWenModel.forward` (modeling_qwen.py:524-537): file names spelled in the token stream between <img>
    and </img> -> host decode -> `ClipPreprocessor(size=448, square=True)` on the GPU (visual.py:354-362, 417-427).
"""
from __future__ import annotations

import json
import math
import os
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .config import QWEN_LORA_TARGETS, QwenModelConfig, TrainConfig
from .engine_qwen import QwenVLDPOEngine
from .plugin import B200ModuleMixin


def qwen_config_from_hf(hf_config, lora_r: int = 64, lora_alpha: float = 16.0) -> QwenModelConfig:
    """QWenConfig (object or parsed config.json) -> QwenModelConfig."""
    g = (lambda k, d=None: hf_config.get(k, d)) if isinstance(hf_config, dict) else (lambda k, d=None: getattr(hf_config, k, d))
    v = g("visual")
    if v is None:
        raise ValueError("not a Qwen-VL config: no `visual` section")
    hidden, heads = g("hidden_size"), g("num_attention_heads")
    if g("kv_channels", hidden // heads) * heads != hidden:
        raise ValueError("kv_channels * num_attention_heads != hidden_size is not supported")
    n_queries = v.get("n_queries", 256)
    return QwenModelConfig(
        image_size=v["image_size"], patch_size=v["patch_size"], v_width=v["width"], v_layers=v["layers"], v_heads=v["heads"],
        v_mlp=int(v["width"] * v["mlp_ratio"]), n_queries=n_queries, hidden=hidden, layers=g("num_hidden_layers"), heads=heads,
        ff=g("intermediate_size") // 2, vocab=g("vocab_size"), rms_eps=g("layer_norm_epsilon", 1e-6),
        rope_theta=float(g("rotary_emb_base", 10000.0)), image_start_id=v["image_start_id"],
        pad_token_id=g("pad_token_id") if g("pad_token_id") is not None else 151643, lora_r=int(lora_r),
        lora_alpha=float(lora_alpha), max_positions=max(2048, int(g("seq_length", 2048))))


def image_paths_from_ids(input_ids: torch.Tensor, image_start_id: int) -> List[str]:
    """modeling_qwen.py:524-534: for every <img> … </img> span (row-major order) the UTF-8 path spelled by the tokens up
    to the first <imgpad> (image_start_id + 2)."""
    ids = input_ids.cpu()
    bos = torch.where(ids == image_start_id)
    eos = torch.where(ids == image_start_id + 1)
    if not bool((bos[0] == eos[0]).all()):
        raise ValueError("unbalanced <img> / </img> markers")
    out = []
    for i, a, b in zip(bos[0].tolist(), bos[1].tolist(), eos[1].tolist()):
        span = ids[i][a + 1:b - 1].tolist()
        span = span[: span.index(image_start_id + 2)]
        out.append(bytes(span).decode("utf-8"))
    return out


class B200QwenVLForRL(B200ModuleMixin, nn.Module):
    def __init__(self, cfg: QwenModelConfig, train: Optional[TrainConfig] = None, device: str = "cuda",
                 with_optimizer: bool = True):
        super().__init__()
        self.engine = QwenVLDPOEngine(cfg, train, device=device, with_optimizer=with_optimizer)
        self.cfg = cfg
        grads = self.engine.hf_state("grad")
        # trainable = the adapters; the base LM is frozen under LoRA (peft freezes every base parameter)
        self._register_engine_params(self.engine.hf_state("policy"), grads, lambda n: n in grads)
        self._preprocessor = None
        self.hf_config_dict: Optional[dict] = None
        self.base_model_name_or_path: Optional[str] = None

    # ---- the model-side contract of docs/CustomizedModel.md (models/QwenVL/__init__.py:24-46)
    @property
    def default_lora_target(self) -> List[str]:
        return ["c_attn", "attn.c_proj", "w1", "w2"]

    def get_vision_tower(self):
        return self.engine.vparams

    def freeze_vision_tower(self):
        pass  # frozen by construction on this path (--freeze_vision_tower True; peft freezes attn_pool as well)

    def prepare_default_generation_kwargs(self, generation_config):
        generation_config.stop_words_ids = [[151645], [151644]]
        generation_config.do_sample = False
        return dict(generation_config=generation_config)

    def forward(self, *a, **k):
        raise RuntimeError("B200QwenVLForRL is driven through concatenated_forward / engine.train_step; generation and "
                           "evaluation forwards are outside the hot path this package replaces")

    # ---- images named inside the token stream
    def pixel_values_for(self, input_ids: torch.Tensor, image_loader=None) -> torch.Tensor:
        from .collator import load_rgb
        from .preprocess import ClipPreprocessor
        if self._preprocessor is None:
            self._preprocessor = ClipPreprocessor(size=self.cfg.image_size, square=True, device=str(self.engine.device))
        loader = image_loader or load_rgb
        return self._preprocessor([loader(p) for p in image_paths_from_ids(input_ids, self.cfg.image_start_id)])

    # ---- checkpoints
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, *args, config=None, torch_dtype=None,
                        train: Optional[TrainConfig] = None, device: str = "cuda", with_optimizer: bool = True,
                        lora_r: int = 64, lora_alpha: float = 16.0, lora_seed: int = 0, **kwargs):
        from . import checkpoint
        if torch_dtype not in (None, torch.bfloat16, "bfloat16", "auto"):
            raise ValueError(f"torch_dtype {torch_dtype}: the B200 path computes in bf16 only")
        path = pretrained_model_name_or_path
        if not os.path.isdir(path):
            raise FileNotFoundError(f"{path}: a local checkpoint directory is required (there is no hub access)")
        with open(os.path.join(path, "config.json")) as f:
            cfg_dict = json.load(f)
        model = cls(qwen_config_from_hf(config if config is not None else cfg_dict, lora_r, lora_alpha), train, device=device,
                    with_optimizer=with_optimizer)
        model.hf_config_dict, model.base_model_name_or_path, model.config = cfg_dict, path, config
        eng = model.engine
        eng.load_state_dict_tensors(checkpoint.iter_checkpoint(path))
        model.reset_adapters(lora_seed)
        adapter = os.path.join(path, "adapter_model.safetensors")
        if os.path.exists(adapter):
            model.load_adapter(path)
        return model

    def reset_adapters(self, seed: int = 0):
        """peft's LoRA init: A ~ kaiming_uniform(a=sqrt(5)) = U(-1/sqrt(in), 1/sqrt(in)), B = 0 (policy == reference)."""
        eng = self.engine
        views = eng.lora_views(eng.policy)
        gen = torch.Generator().manual_seed(seed)
        for name, t in views.items():
            if name.endswith("lora_A"):
                bound = 1.0 / math.sqrt(t.shape[1])
                t.copy_(((torch.rand(t.shape, generator=gen) * 2 - 1) * bound).to(t.device, torch.bfloat16))
            else:
                t.zero_()
        eng.sync_master_from_params()

    def save_pretrained(self, save_directory: str, **kwargs):
        """PEFT-format adapter checkpoint of the trained LoRA weights."""
        from safetensors.torch import save_file
        os.makedirs(save_directory, exist_ok=True)
        eng = self.engine
        state = {f"base_model.model.{k}.weight": v.detach().to("cpu").contiguous()
                 for k, v in eng.lora_views(eng.policy).items()}
        eng.wait_optimizer()
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
        eng = self.engine
        views = eng.lora_views(eng.policy)
        with safe_open(os.path.join(directory, "adapter_model.safetensors"), framework="pt", device="cpu") as f:
            for k in f.keys():
                name = k[len("base_model.model."):] if k.startswith("base_model.model.") else k
                name = name[:-len(".weight")] if name.endswith(".weight") else name
                name = name.replace(".lora_A.default", ".lora_A").replace(".lora_B.default", ".lora_B")
                if name not in views:
                    raise KeyError(f"adapter tensor {k} has no counterpart (targets: {QWEN_LORA_TARGETS})")
                views[name].copy_(f.get_tensor(k).to(eng.device, torch.bfloat16))
        eng.sync_master_from_params()


def check_peft_config(model: B200QwenVLForRL, peft_config) -> None:
    """The trainer hands the LoraConfig of utils/auto_load.py:559-571 to TRL; here the adapters live in the engine, so the
    config must describe exactly what was allocated."""
    if peft_config is None:
        raise ValueError("the Qwen-VL B200 path trains LoRA adapters only: run with --use_lora True")
    r, alpha = getattr(peft_config, "r"), getattr(peft_config, "lora_alpha")
    targets = sorted(getattr(peft_config, "target_modules"))
    if r != model.cfg.lora_r or float(alpha) != float(model.cfg.lora_alpha):
        raise ValueError(f"LoraConfig(r={r}, lora_alpha={alpha}) != engine adapters (r={model.cfg.lora_r}, "
                         f"alpha={model.cfg.lora_alpha}); pass lora_r / lora_alpha to from_pretrained")
    if targets != sorted(model.default_lora_target):
        raise ValueError(f"lora_target_modules {targets}: this path adapts exactly {sorted(model.default_lora_target)}")


def install_qwen():
    """`vlrlhf.models.QwenVL.core_mapper` -> (B200QwenVLForRL, reference processor/collators, B200 DPO trainer)."""
    import importlib
    from . import plugin
    qwen = importlib.import_module("vlrlhf.models.QwenVL")
    from vlrlhf.models.utils import ModelCoreMapper
    ref = qwen.core_mapper

    # keeps QwenVLDPOTrainer.tokenize_row (:257-347); the adapters are the engine's, not peft modules
    QwenVLB200DPOTrainer = plugin.make_trainer_class(ref.dpo_trainer, check_peft=check_peft_config, name="QwenVLB200DPOTrainer")

    qwen.core_mapper = ModelCoreMapper(
        model=B200QwenVLForRL, processor=ref.processor, dpo_collator=ref.dpo_collator, dpo_trainer=QwenVLB200DPOTrainer,
        reward_model=ref.reward_model, value_model=ref.value_model, reward_collator=ref.reward_collator,
        reward_trainer=ref.reward_trainer, sft_collator=ref.sft_collator, sft_trainer=ref.sft_trainer,
        ppo_collator=ref.ppo_collator, ppo_trainer=ref.ppo_trainer)
    return qwen.core_mapper
