"""Import the UNMODIFIED reference (TideDra/VL-RLHF) from /root/reference/src.

Only usable in the build container (the GPU box has no /root/reference).  Used by
`oracle/make_fixtures.py` to mint golden vectors and by `tests/test_oracle_vs_reference.py`
(skipped when the reference tree is absent).

The reference imports trl / peft / accelerate / deepspeed at module import time
(base/trainer.py:3,16,21-25; utils/common.py:4-7); none is installed and there is no
network, so inert stub modules are registered first.  transformers here is 5.5.0 (the
reference pins 4.41.0): `LlavaShim` re-exposes the attributes the reference's
`LlavaForRL.forward` (models/Llava/__init__.py:111-271) reads, and re-adds the
`logits.float()` upcast that transformers-4.41 `LlamaForCausalLM.forward` performed.
"""
import os
import sys
import types

REFERENCE_SRC = "/root/reference/src"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "vlrlhf"))


_installed = False


def install():
    """Register stubs and put the reference on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not present at " + REFERENCE_SRC)
    # real libs FIRST: transformers' lazy is_*_available() probes choke on spec-less stubs
    import transformers  # noqa: F401
    import datasets  # noqa: F401
    from transformers import PreTrainedModel, TrainingArguments, Trainer  # noqa: F401
    from transformers.trainer_callback import TrainerCallback  # noqa: F401
    from transformers.trainer_utils import EvalPrediction, EvalLoopOutput  # noqa: F401

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Inert:
        def __init__(self, *a, **k):
            pass

    for mod in ("wandb", "loguru"):
        try:
            __import__(mod)
        except Exception:
            if mod == "loguru":
                import logging
                stub("loguru", logger=logging.getLogger("loguru-stub"))
            else:
                stub(mod)
    stub("trl", DPOTrainer=_Inert, PPOTrainer=_Inert, PPOConfig=_Inert, SFTTrainer=_Inert,
         RewardTrainer=_Inert, AutoModelForCausalLMWithValueHead=_Inert, RewardConfig=_Inert)
    stub("trl.trainer")
    stub("trl.trainer.reward_config", RewardConfig=_Inert)
    stub("peft", PeftConfig=_Inert, LoraConfig=_Inert, PeftModel=_Inert, get_peft_model=None,
         prepare_model_for_kbit_training=None)
    stub("accelerate")
    stub("accelerate.utils", gather_object=None, tqdm=None)
    stub("deepspeed", zero=None)
    stub("deepspeed.runtime")
    stub("deepspeed.runtime.zero")
    stub("deepspeed.runtime.zero.partition_parameters", ZeroParamStatus=None)
    transformers.__dict__["deepspeed"] = stub("transformers.deepspeed",
                                              is_deepspeed_zero3_enabled=lambda: False)
    sys.path.insert(0, REFERENCE_SRC)
    _installed = True


def reference_symbols():
    """-> (VLDPOTrainer, get_diff_ids, LlavaShim) from the reference tree."""
    install()
    from vlrlhf.base.trainer import VLDPOTrainer
    from vlrlhf.utils.diff_lib import get_diff_ids
    from oracle._llava_shim import LlavaShim
    return VLDPOTrainer, get_diff_ids, LlavaShim
