"""vlrlhf_b200: B200-native drop-in for the VL-RLHF DPO hot path
(`VLDPOTrainer.concatenated_forward -> get_batch_logps -> dpo_loss`, reference
src/vlrlhf/base/trainer.py:190-301 and the LLaVA forward it drives).

Layout: csrc/ (hand-written sm_100a CUDA + the C ABI declared in include/vlb200.h),
_lib.py (ctypes binding), ops.py (tensor-level wrappers), engine.py (the DPO step),
plugin.py (the reference's ModelCoreMapper / VLDPOTrainer interface).
There is no CPU or PyTorch fallback: if libvlb200.so is missing, importing `ops` raises.
"""
__version__ = "0.1.0"
