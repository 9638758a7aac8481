"""The B200-native DPO step for the LLaVA-1.5 family and LLaVA-Next (anyres crops + image_newline, GQA decoder).

Replaces, behind the reference's own interfaces (see plugin.py):
  * `LlavaForRL.forward` incl. `_merge_input_ids_with_image_features`   (models/Llava/__init__.py:36-271)
  * `LlavaNextForRL.forward` incl. its merge and `pack_image_features`   (models/LlavaNext/__init__.py:38-345)
  * `VLDPOTrainer.concatenated_forward / get_batch_logps / dpo_loss`      (base/trainer.py:148-301)
  * trl `DPOTrainer.get_batch_loss_metrics` (policy pass, no-grad reference pass, loss, reward stats)
  * autograd backward + AdamW + the data-parallel gradient all-reduce (accelerate/DeepSpeed in the reference)

Every FLOP/byte of device work is a libvlb200 kernel (ops.py -> C ABI).  torch provides device memory,
the CUDA stream and torch.distributed (NCCL) only.  Differences from the reference that do not change
results: the frozen vision tower runs once per pair (the reference runs it 4x on identical pixels), the
full-vocab logits are computed only for rows that can carry a label, the discarded CE loss
(Llava/__init__.py:245-257) is not computed, and there is no per-step empty_cache()/gc (trainer.py:303-308).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from .config import ModelConfig, TrainConfig, tensor_seed, weight_specs

ALIGN = 128  # elements (256 B): keeps every tensor TMA/16-byte aligned inside the flat arenas


class Arena:
    """Flat bf16 buffer with named row-major tensors carved out of it."""

    def __init__(self):
        self.shapes: Dict[str, Tuple[int, ...]] = {}
        self.offsets: Dict[str, int] = {}
        self.size = 0
        self.flat: Optional[torch.Tensor] = None

    def add(self, name: str, shape: Tuple[int, ...]):
        assert name not in self.shapes
        n = 1
        for s in shape:
            n *= s
        self.shapes[name] = shape
        self.offsets[name] = self.size
        self.size += (n + ALIGN - 1) // ALIGN * ALIGN

    def allocate(self, device, dtype=torch.bfloat16):
        self.flat = torch.zeros(self.size, dtype=dtype, device=device)
        return self

    def view_of(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        shape = self.shapes[name]
        n = 1
        for s in shape:
            n *= s
        off = self.offsets[name]
        return flat[off:off + n].view(shape)

    def __getitem__(self, name: str) -> torch.Tensor:
        return self.view_of(self.flat, name)


class Weights:
    """Name -> tensor view for one copy of the model (policy, reference) in engine layout."""

    def __init__(self, arena: Arena, flat: torch.Tensor):
        self.t = {n: arena.view_of(flat, n) for n in arena.shapes}

    def __getitem__(self, k):
        return self.t[k]


def _trainable_layout(cfg: ModelConfig) -> Arena:
    a = Arena()
    d = cfg.hidden
    a.add("proj.w1", (d, cfg.v_hidden)); a.add("proj.b1", (d,)); a.add("proj.w2", (d, d)); a.add("proj.b2", (d,))
    if cfg.family == "llava_next":
        a.add("image_newline", (d,))
    a.add("embed", (cfg.vocab, d))
    for i in range(cfg.layers):
        a.add(f"L{i}.ln1", (d,)); a.add(f"L{i}.wqkv", (cfg.qkv_dim, d)); a.add(f"L{i}.wo", (d, cfg.heads * cfg.head_dim))
        a.add(f"L{i}.ln2", (d,)); a.add(f"L{i}.wgu", (2 * cfg.ff, d)); a.add(f"L{i}.wd", (d, cfg.ff))
    a.add("norm", (d,)); a.add("lm_head", (cfg.vocab, d))
    return a


def _vision_layout(cfg: ModelConfig) -> Arena:
    a = Arena()
    dv = cfg.v_hidden
    a.add("v.cls", (dv,)); a.add("v.patch", (dv, cfg.patch_k_padded)); a.add("v.pos", (cfg.n_patches + 1, dv))
    a.add("v.pre.w", (dv,)); a.add("v.pre.b", (dv,))
    for i in range(cfg.v_used_layers):
        a.add(f"v{i}.ln1.w", (dv,)); a.add(f"v{i}.ln1.b", (dv,)); a.add(f"v{i}.wqkv", (3 * dv, dv)); a.add(f"v{i}.bqkv", (3 * dv,))
        a.add(f"v{i}.wo", (dv, dv)); a.add(f"v{i}.bo", (dv,)); a.add(f"v{i}.ln2.w", (dv,)); a.add(f"v{i}.ln2.b", (dv,))
        a.add(f"v{i}.w1", (cfg.v_ff, dv)); a.add(f"v{i}.b1", (cfg.v_ff,)); a.add(f"v{i}.w2", (dv, cfg.v_ff)); a.add(f"v{i}.b2", (dv,))
    return a


def hf_views(cfg: ModelConfig, w: Dict[str, torch.Tensor], vision: Optional[Dict[str, torch.Tensor]]) -> Dict[str, torch.Tensor]:
    """HF-4.41-named views into the engine's (fused) storage: q/k/v_proj and gate/up_proj are row slices."""
    out: Dict[str, torch.Tensor] = {}
    hd = cfg.heads * cfg.head_dim
    kvd = cfg.kv_heads * cfg.head_dim
    out["multi_modal_projector.linear_1.weight"] = w["proj.w1"]; out["multi_modal_projector.linear_1.bias"] = w["proj.b1"]
    out["multi_modal_projector.linear_2.weight"] = w["proj.w2"]; out["multi_modal_projector.linear_2.bias"] = w["proj.b2"]
    out["language_model.model.embed_tokens.weight"] = w["embed"]
    if cfg.family == "llava_next":
        out["image_newline"] = w["image_newline"]
    for i in range(cfg.layers):
        p = f"language_model.model.layers.{i}."
        qkv, gu = w[f"L{i}.wqkv"], w[f"L{i}.wgu"]
        out[p + "input_layernorm.weight"] = w[f"L{i}.ln1"]
        out[p + "self_attn.q_proj.weight"] = qkv[:hd]
        out[p + "self_attn.k_proj.weight"] = qkv[hd:hd + kvd]
        out[p + "self_attn.v_proj.weight"] = qkv[hd + kvd:]
        out[p + "self_attn.o_proj.weight"] = w[f"L{i}.wo"]
        out[p + "post_attention_layernorm.weight"] = w[f"L{i}.ln2"]
        out[p + "mlp.gate_proj.weight"] = gu[:cfg.ff]
        out[p + "mlp.up_proj.weight"] = gu[cfg.ff:]
        out[p + "mlp.down_proj.weight"] = w[f"L{i}.wd"]
    out["language_model.model.norm.weight"] = w["norm"]
    out["language_model.lm_head.weight"] = w["lm_head"]
    if vision is not None:
        v = vision
        vp = "vision_tower.vision_model."
        dv = cfg.v_hidden
        out[vp + "embeddings.class_embedding"] = v["v.cls"]
        out[vp + "embeddings.patch_embedding.weight"] = v["v.patch"][:, :cfg.patch_k]  # [dv, 3*p*p] strided view
        out[vp + "embeddings.position_embedding.weight"] = v["v.pos"]
        out[vp + "pre_layrnorm.weight"] = v["v.pre.w"]; out[vp + "pre_layrnorm.bias"] = v["v.pre.b"]
        for i in range(cfg.v_used_layers):
            p = f"{vp}encoder.layers.{i}."
            out[p + "layer_norm1.weight"] = v[f"v{i}.ln1.w"]; out[p + "layer_norm1.bias"] = v[f"v{i}.ln1.b"]
            out[p + "layer_norm2.weight"] = v[f"v{i}.ln2.w"]; out[p + "layer_norm2.bias"] = v[f"v{i}.ln2.b"]
            for j, pr in enumerate(("q_proj", "k_proj", "v_proj")):
                out[p + f"self_attn.{pr}.weight"] = v[f"v{i}.wqkv"][j * dv:(j + 1) * dv]
                out[p + f"self_attn.{pr}.bias"] = v[f"v{i}.bqkv"][j * dv:(j + 1) * dv]
            out[p + "self_attn.out_proj.weight"] = v[f"v{i}.wo"]; out[p + "self_attn.out_proj.bias"] = v[f"v{i}.bo"]
            out[p + "mlp.fc1.weight"] = v[f"v{i}.w1"]; out[p + "mlp.fc1.bias"] = v[f"v{i}.b1"]
            out[p + "mlp.fc2.weight"] = v[f"v{i}.w2"]; out[p + "mlp.fc2.bias"] = v[f"v{i}.b2"]
    return out


def attn_forward(m: "ops.MergeIndex", q, k, v, out, lse, H: int, KV: int, dh: int, scale: float):
    """Causal fused attention over the sequences of a merge index (padded, packed or shared-prefix rows)."""
    a = m.attn()
    return ops.attn_fwd_tc(q, k, v, out, lse, a["seqlens"], a["B"], a["S"], H, KV, dh, True, scale,
                           **{k_: a[k_] for k_ in ("row_starts", "total_rows", "ctx", "kids") if a.get(k_) is not None})


def attn_backward(m: "ops.MergeIndex", q, k, v, out, dout, lse, delta, dq, dk, dv, H: int, KV: int, dh: int, scale: float):
    a = m.attn()
    return ops.attn_bwd_tc(q, k, v, out, dout, lse, delta, dq, dk, dv, a["seqlens"], a["B"], a["S"], H, KV, dh, True, scale,
                           **{k_: a[k_] for k_ in ("row_starts", "total_rows", "ctx", "kids") if a.get(k_) is not None})


class StepOutput:
    __slots__ = ("loss", "losses", "chosen_rewards", "rejected_rewards", "stats", "policy_logps", "ref_logps", "grad_norm")


class LlavaDPOEngine:
    has_ref_copy = True       # a frozen reference copy of the trainable arena (full fine-tuning)
    needs_embed_grad = True   # fp32 scatter target for the embedding gradient

    def _make_layouts(self) -> Tuple[Arena, Arena]:
        """-> (trainable arena, frozen vision arena); subclasses add their own frozen arenas in _alloc_family()."""
        return _trainable_layout(self.cfg), _vision_layout(self.cfg)

    def _alloc_family(self):
        pass

    def __init__(self, cfg: ModelConfig, train: Optional[TrainConfig] = None, device: str = "cuda",
                 with_optimizer: bool = True, process_group=None):
        self.cfg, self.tc = cfg, train or TrainConfig()
        self.device = torch.device(device)
        self.pg = process_group
        self.layout, self.vlayout = self._make_layouts()
        n = self.layout.size
        import os as _os
        if train is None and _os.environ.get("VLB200_PACK_SEQUENCES", "0") == "1":
            self.tc.pack_sequences = True   # launcher-level switch: the reference's dpo.py builds the model without a TrainConfig
        if train is None and _os.environ.get("VLB200_SHARE_PREFIX", "0") == "1":
            self.tc.share_prefix = True
        # Optimizer sharding across the data-parallel ranks (ZeRO-1 style; the reference's default DeepSpeed config
        # shards optimizer state too, accelerate_config/zero2.yaml): gradients are reduce-SCATTERED, each rank runs
        # AdamW on its 1/world slice of the flat buffers (fp32 master + moments exist for that slice only) and the
        # updated bf16 parameters are all-gathered.  Same wire traffic as one all-reduce, AdamW time and optimizer
        # memory divided by world.  Results equal the replicated update (AdamW is elementwise).
        world = self.world_size()
        self.shard_optimizer = world > 1 and with_optimizer and _os.environ.get("VLB200_SHARD_OPTIMIZER", "1") != "0"
        gran = ALIGN * world
        n_flat = (n + gran - 1) // gran * gran if self.shard_optimizer else n
        self.n_flat = n_flat
        rank = torch.distributed.get_rank(self.pg) if world > 1 else 0
        self.shard_lo, self.shard_hi = (rank * (n_flat // world), (rank + 1) * (n_flat // world)) if self.shard_optimizer \
            else (0, n_flat)
        self.params = torch.zeros(n_flat, dtype=torch.bfloat16, device=self.device)  # policy (projector + LLM)
        self.ref_params = torch.zeros(n if self.has_ref_copy else 0, dtype=torch.bfloat16, device=self.device)  # frozen reference copy
        self.grads = torch.zeros(n_flat, dtype=torch.bfloat16, device=self.device)
        self.vparams = torch.zeros(self.vlayout.size, dtype=torch.bfloat16, device=self.device)  # frozen vision tower
        self.policy = Weights(self.layout, self.params)
        self.ref = Weights(self.layout, self.ref_params) if self.has_ref_copy else None
        self.g = Weights(self.layout, self.grads)
        self.vis = Weights(self.vlayout, self.vparams)
        self.with_optimizer = with_optimizer
        if with_optimizer:
            ns = self.shard_hi - self.shard_lo
            self.master = torch.zeros(ns, dtype=torch.float32, device=self.device)
            self.exp_avg = torch.zeros(ns, dtype=torch.float32, device=self.device)
            self.exp_avg_sq = torch.zeros(ns, dtype=torch.float32, device=self.device)
        self.dembed_f32 = torch.zeros(cfg.vocab, cfg.hidden, dtype=torch.float32, device=self.device) \
            if self.needs_embed_grad else None
        self._alloc_family()
        self.sumsq_ws = torch.zeros(1024, dtype=torch.float32, device=self.device)
        self.grad_sumsq = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.opt_step = 0
        # SwiGLU backward inside the down-projection dgrad GEMM's epilogue (vlb200_gemm_swiglu_bwd_bf16: dact never reaches HBM,
        # two elementwise launches per layer fewer, bit-identical results).  Measured on a B200 (profiles/r2d_*): 551 vs 543
        # ms/step -- the epilogue's row-per-thread gate|up reads and writes (32 rows per warp request) cost more than the two
        # coalesced elementwise passes save, so it stays opt-in until the epilogue stages through shared memory.
        self.fuse_swiglu_bwd = _os.environ.get("VLB200_FUSE_SWIGLU_BWD", "0") == "1"
        self.force_logit_means = False   # plugin: also produce TRL's logits/* means on no-grad passes (evaluation)
        self._micro_step = 0       # train_step calls so far (gradient_accumulation_steps micro-batches per optimizer step)
        self.last_lr = 0.0
        # VLB200_OVERLAP_ALLREDUCE=1 launches per-layer gradient buckets on the NCCL stream while backward continues.
        # Measured on 2xB200 (profiles/r1_bench_7b_2gpu_overlap.md): 697.4 ms/step overlapped vs 689.9 ms with ONE
        # all-reduce after backward -- the NCCL kernels take SMs from the persistent GEMMs' static tile schedule and
        # cost more than the ~28 ms they hide, so the single all-reduce is the default.
        self.overlap_allreduce = _os.environ.get("VLB200_OVERLAP_ALLREDUCE", "0") == "1"
        self._pending = []
        # Deferred optimizer: the gradient reduction + grad-norm + AdamW (+ parameter all-gather) of step t run on a
        # side stream and overlap the frozen-reference forward of step t+1, which does not read the policy weights.
        # AdamW/sumsq are HBM-bound kernels without shared memory, so their CTAs co-reside with the persistent GEMM
        # CTAs instead of displacing them.  The policy forward waits on `opt_done` (device-side event wait).
        self.async_optimizer = self.device.type == "cuda" and _os.environ.get("VLB200_ASYNC_OPTIMIZER", "1") != "0"
        self._opt_pending = False
        if self.device.type == "cuda":
            self.opt_stream = torch.cuda.Stream(device=self.device)
            self.opt_done = torch.cuda.Event()
            self.norm_done = torch.cuda.Event()
        self._anyres = None  # host.AnyresPlan of the batch in flight (LLaVA-Next only)
        self._bufs: Dict[str, torch.Tensor] = {}
        self._stores: Dict[str, torch.Tensor] = {}   # flat storage behind the (possibly smaller) views in _bufs
        self._pad_rows = 0                            # packed rows: padded row count (capacity) and row count of the batch
        self._cur_rows = 0
        self._build_rope_tables()

    # ------------------------------------------------------------------ weights
    def _build_rope_tables(self, n_positions: Optional[int] = None):
        # identical arithmetic to LlamaRotaryEmbedding (modeling_llama.py:138-168): fp32 inv_freq, fp32 angles
        cfg = self.cfg
        dh = cfg.head_dim
        inv_freq = 1.0 / (cfg.rope_theta ** (torch.arange(0, dh, 2, dtype=torch.int64).float() / dh))
        self.rope_len = int(n_positions or cfg.max_positions)
        t = torch.arange(self.rope_len, dtype=torch.float32)
        freqs = t[:, None] * inv_freq[None, :]
        self.rope_cos = freqs.cos().contiguous().to(self.device)
        self.rope_sin = freqs.sin().contiguous().to(self.device)

    def ensure_rope_len(self, n_positions: int):
        """HF's rotary embedding grows its cos/sin cache on demand (modeling_llama.py:150-168); the tables here are rebuilt
        when a merged sequence is longer than they are (e.g. --max_length 4096 + 575 image rows), never read out of bounds."""
        if n_positions > self.rope_len:
            self._build_rope_tables(max(int(n_positions), 2 * self.rope_len))

    def wait_optimizer(self):
        """Order the current stream after the deferred optimizer step (no host block).  Called before anything that
        reads the policy weights or writes the gradient buffer; call it yourself before touching `params`,
        `master`, `grads` … directly."""
        if self._opt_pending:
            torch.cuda.current_stream(self.device).wait_event(self.opt_done)
            self._opt_pending = False

    def hf_state(self, which: str = "policy") -> Dict[str, torch.Tensor]:
        self.wait_optimizer()
        w = {"policy": self.policy, "ref": self.ref, "grad": self.g}[which]
        return hf_views(self.cfg, w.t, self.vis.t if which != "grad" else None)

    def init_synthetic(self, seed: int, ref_alpha: float = 0.05):
        """Seeded random-init weights of the architecture (bit-identical to oracle.restate.make_policy_and_ref)."""
        self.wait_optimizer()
        cfg = self.cfg
        pol, ref = self.hf_state("policy"), self.hf_state("ref")
        tmp_cache: Dict[int, torch.Tensor] = {}

        def tmp(n):
            if n not in tmp_cache:
                tmp_cache[n] = torch.empty(n, dtype=torch.bfloat16, device=self.device)
            return tmp_cache[n]

        for name, shape, scale, shift in weight_specs(cfg):
            if name not in pol:
                continue  # vision layers above vision_feature_layer are never used by the path
            dst = pol[name]
            n = dst.numel()
            if dst.is_contiguous():
                ops.init_uniform_(dst.view(-1), tensor_seed(name, seed), scale, shift)
            else:  # padded patch-embedding weight
                t = tmp(n)
                ops.init_uniform_(t, tensor_seed(name, seed), scale, shift)
                dst.copy_(t.view(dst.shape))
            if name.startswith("vision_tower."):
                continue
            other = tmp(n)
            ops.init_uniform_(other, tensor_seed(name, seed + 1), scale, shift)
            ops.perturb_(ref[name].view(-1), dst.view(-1), other, ref_alpha, 1.0 if name.endswith("norm.weight") else 0.0)
        if self.with_optimizer:
            ops.cast_bf16_to_f32(self.params[self.shard_lo:self.shard_hi], self.master)
        if self.device.type == "cuda":
            torch.cuda.synchronize()

    def sync_master_from_params(self):
        """fp32 master copy of this rank's optimizer slice <- the bf16 policy parameters (after loading weights)."""
        self.wait_optimizer()
        if self.with_optimizer:
            ops.cast_bf16_to_f32(self.params[self.shard_lo:self.shard_hi], self.master)

    def load_hf_state_dict(self, sd: Dict[str, torch.Tensor], which: str = "policy"):
        dst = self.hf_state(which)  # (waits for a deferred optimizer step)
        for k, v in sd.items():
            if k in dst:
                dst[k].copy_(v.to(device=self.device, dtype=torch.bfloat16).view(dst[k].shape))
        if which == "policy" and self.with_optimizer:
            ops.cast_bf16_to_f32(self.params[self.shard_lo:self.shard_hi], self.master)

    # ------------------------------------------------------------------ buffers
    def buf(self, name: str, shape, dtype=torch.bfloat16) -> torch.Tensor:
        """Named workspace: allocated once, reused every step.  Packed rows (TrainConfig.pack_sequences): the row count changes
        with every batch, so a workspace whose leading dimension is the batch's row count reserves the padded row count and
        later requests are served as views of that storage (it only ever grows)."""
        t = self._bufs.get(name)
        shape = tuple(int(s) for s in shape)
        if t is not None and t.dtype == dtype and tuple(t.shape) == shape:
            return t
        if not self._pad_rows:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._bufs[name] = t
            return t
        n = 1
        for dim in shape:
            n *= dim
        cap = n // shape[0] * self._pad_rows if shape and shape[0] == self._cur_rows and self._cur_rows < self._pad_rows else n
        store = self._stores.get(name)
        if store is None or store.dtype != dtype or store.numel() < n:
            store = torch.empty(max(n, cap), dtype=dtype, device=self.device)
            self._stores[name] = store
        t = store[:n].view(shape)
        self._bufs[name] = t
        return t

    # ------------------------------------------------------------------ vision tower (frozen, once per pair)
    def vision_features(self, pixels: torch.Tensor) -> torch.Tensor:
        cfg, v = self.cfg, self.vis
        Bv = pixels.shape[0]
        P, Sv, dv = cfg.n_patches, cfg.n_patches + 1, cfg.v_hidden
        K, Kp = cfg.patch_k, cfg.patch_k_padded
        patches = self.buf("v.patches", (Bv * P, Kp))
        ops.clip_im2col(pixels, cfg.patch_size, patches)
        x = self.buf("v.x", (Bv * Sv, dv))
        wpatch = v["v.patch"][:, :K]
        pos = v["v.pos"]
        for b in range(Bv):  # conv-as-GEMM; epilogue adds the position embedding (K1)
            ops.gemm(patches[b * P:(b + 1) * P, :K], wpatch, out=x[b * Sv + 1:(b + 1) * Sv], residual=pos[1:])
        ops.clip_cls_rows_(x, v["v.cls"], pos[0], Bv, Sv)
        h = self.buf("v.h", (Bv * Sv, dv))
        ops.layernorm_fwd(x, v["v.pre.w"], v["v.pre.b"], cfg.v_eps, out=h)
        xb = x                                                       # bf16 scratch (embeddings no longer needed)
        x = self.buf("v.x32", (Bv * Sv, dv), torch.float32)          # fp32 residual stream, like the decoder
        ops.cast_bf16_to_f32(h.view(-1), x.view(-1))
        qkv = self.buf("v.qkv", (Bv * Sv, 3 * dv))
        att = self.buf("v.att", (Bv * Sv, dv))
        f = self.buf("v.f", (Bv * Sv, cfg.v_ff))
        scale = cfg.v_head_dim ** -0.5
        for i in range(cfg.v_used_layers):
            ops.layernorm_fwd(x, v[f"v{i}.ln1.w"], v[f"v{i}.ln1.b"], cfg.v_eps, out=h)
            ops.gemm(h, v[f"v{i}.wqkv"], out=qkv, bias=v[f"v{i}.bqkv"])
            ops.attn_fwd_tc(qkv[:, :dv], qkv[:, dv:2 * dv], qkv[:, 2 * dv:], att, None, None, Bv, Sv, cfg.v_heads, cfg.v_heads,
                         cfg.v_head_dim, False, scale)
            ops.gemm(att, v[f"v{i}.wo"], out=x, bias=v[f"v{i}.bo"], residual=x)
            ops.layernorm_fwd(x, v[f"v{i}.ln2.w"], v[f"v{i}.ln2.b"], cfg.v_eps, out=h)
            ops.gemm(h, v[f"v{i}.w1"], out=f, bias=v[f"v{i}.b1"], act=ops.ACT_QUICK_GELU)
            ops.gemm(f, v[f"v{i}.w2"], out=x, bias=v[f"v{i}.b2"], residual=x)
        feats = self.buf("v.feats", (Bv * P, dv))  # drop CLS (Llava/__init__.py:182-183)
        ops.cast_f32_to_bf16(x.view(-1), xb.view(-1))
        ops.copy_rows(xb, Sv * dv, dv, 1, feats, P * dv, dv, Bv, P, dv)
        return feats

    # ------------------------------------------------------------------ one decoder layer
    def _layer_bufs(self, pre: str, sfx: str, m: "ops.MergeIndex") -> Dict[str, torch.Tensor]:
        cfg = self.cfg
        T = m.T
        H, dh = cfg.heads, cfg.head_dim
        return dict(rstd1=self.buf(f"{pre}.rstd1{sfx}", (T,), torch.float32),
                    rstd2=self.buf(f"{pre}.rstd2{sfx}", (T,), torch.float32),
                    qkv=self.buf(f"{pre}.qkv{sfx}", (T, cfg.qkv_dim)), att=self.buf(f"{pre}.att{sfx}", (T, H * dh)),
                    lse=self.buf(f"{pre}.lse{sfx}", (m.n_attn_seq, H, m.S), torch.float32),
                    xmid=self.buf(f"{pre}.xmid{sfx}", (T, cfg.hidden), torch.float32),
                    gu=self.buf(f"{pre}.gu{sfx}", (T, 2 * cfg.ff)))

    def _layer_fwd(self, w: Weights, i: int, x: torch.Tensor, b: Dict[str, torch.Tensor], m: "ops.MergeIndex",
                   xn: Optional[torch.Tensor], keep_gu: bool = True):
        """Decoder layer i on the fp32 residual stream x -> xn (K9-K14).  xn=None stops after the gate|up GEMM: the
        recompute of a checkpointed layer needs the saved-for-backward tensors, not the layer output.  keep_gu=False
        (reference pass, checkpointed forward): the gate|up projections are consumed inside the GEMM epilogue and never
        reach HBM."""
        cfg = self.cfg
        d, T = cfg.hidden, m.T
        H, KV, dh = cfg.heads, cfg.kv_heads, cfg.head_dim
        hd, kvd = H * dh, KV * dh
        h = self.buf("s.h", (T, d))
        qkv, att, xmid, gu = b["qkv"], b["att"], b["xmid"], b["gu"]
        ops.rmsnorm_fwd(x, w[f"L{i}.ln1"], cfg.rms_eps, out=h, rstd=b["rstd1"])
        ops.gemm(h, w[f"L{i}.wqkv"], out=qkv)
        ops.rope_(qkv, m.pos, self.rope_cos, self.rope_sin, H + KV, dh)
        attn_forward(m, qkv[:, :hd], qkv[:, hd:hd + kvd], qkv[:, hd + kvd:], att, b["lse"], H, KV, dh, 1.0 / math.sqrt(dh))
        ops.gemm(att, w[f"L{i}.wo"], out=xmid, residual=x)
        ops.rmsnorm_fwd(xmid, w[f"L{i}.ln2"], cfg.rms_eps, out=h, rstd=b["rstd2"])
        if xn is None:
            ops.gemm(h, w[f"L{i}.wgu"], out=gu)
        else:
            act = self.buf("s.act", (T, cfg.ff))
            ops.gemm_swiglu(h, w[f"L{i}.wgu"], gu, act, write_gu=keep_gu)   # SwiGLU in the gate|up GEMM's epilogue
            ops.gemm(act, w[f"L{i}.wd"], out=xn, residual=xmid)

    # ------------------------------------------------------------------ forward of one model copy
    def _merged_embeddings(self, w: Weights, m: "ops.MergeIndex", feats: torch.Tensor, save_projector: bool,
                           x_name: str) -> torch.Tensor:
        """Projector (K6) -> [LLaVA-Next: pack_image_features] -> token embedding + merge (K7, K8): the fp32 input of
        decoder layer 0, [T, d].  save_projector keeps the pre-GELU activations for the projector's backward."""
        cfg = self.cfg
        d, T = cfg.hidden, m.T
        nimg = feats.shape[0]
        # projector (K6)
        if save_projector:
            z = self.buf("p.z", (nimg, d)); ph = self.buf("p.h", (nimg, d))
            ops.gemm(feats, w["proj.w1"], out=z, bias=w["proj.b1"])
            ops.gelu_fwd(z, ph)
        else:
            ph = self.buf("p.h_ref", (nimg, d))
            ops.gemm(feats, w["proj.w1"], out=ph, bias=w["proj.b1"], act=ops.ACT_GELU_ERF)
        plan = self._anyres
        img = self.buf("p.img", (nimg + (1 if plan is not None else 0), d))
        ops.gemm(ph, w["proj.w2"], out=img[:nimg], bias=w["proj.b2"])
        if plan is not None:
            # LLaVA-Next "spatial_unpad" packing (LlavaNext/__init__.py:240-249): one row gather over the crop
            # features; image_newline sits in the extra last row of `img` and is read once per stitched map row
            ops.copy_rows(w["image_newline"], d, d, 0, img[nimg:], d, d, 1, 1, d)
            packed = self.buf("p.packed", (plan.total_feats, d))
            ops.gather_rows(img, plan.pack_index, packed)
            img = packed
        # embed + merge (K7, K8)
        # the residual stream is kept in fp32 (bf16 would add a 2^-9 relative rounding per layer that compounds
        # through 32 layers); every GEMM operand (normed activations, q/k/v, attention out, SwiGLU out) is bf16
        x = self.buf(x_name, (T, d), torch.float32)
        ops.llava_merge_embed(m, w["embed"], img, x)
        return x

    def _forward(self, w: Weights, m: "ops.MergeIndex", feats: torch.Tensor, tag: str, save: bool,
                 ddpo_weight: Optional[torch.Tensor]):
        cfg = self.cfg
        d, T = cfg.hidden, m.T
        L = cfg.layers
        x = self._merged_embeddings(w, m, feats, save, "x.0" if save else "s.x0")
        ckpt = save and self.tc.activation_checkpointing
        for i in range(L):
            # saved for backward: everything (pre "a", one set per layer) or, with activation checkpointing, only the
            # fp32 layer input x.{i} (the rest is recomputed by _backward into the shared scratch set "s")
            keep = save and not ckpt
            b = self._layer_bufs("a" if keep else "s", f".{i}" if keep else "", m)
            xn = self.buf(f"x.{i + 1}" if save else ("s.x1" if i % 2 == 0 else "s.x0"), (T, d), torch.float32)
            self._layer_fwd(w, i, x, b, m, xn, keep_gu=keep)
            x = xn
        return self._head_forward(x, w["norm"], w["lm_head"], m, feats, save, ddpo_weight)

    def _head_forward(self, x: torch.Tensor, norm_w: torch.Tensor, lm_w: torch.Tensor, m: "ops.MergeIndex", feats, save: bool,
                      ddpo_weight: Optional[torch.Tensor]):
        """Final RMSNorm -> lm_head on the rows that can carry a label (K15) -> fused log-prob gather (K16)."""
        cfg = self.cfg
        d, T = cfg.hidden, m.T
        h = self.buf("s.h", (T, d))
        rstd_f = self.buf("a.rstd_f" if save else "s.rstd_f", (T,), torch.float32)
        ops.rmsnorm_fwd(x, norm_w, cfg.rms_eps, out=h, rstd=rstd_f)
        R = m.n_seq * (m.L - 1)
        hsel = self.buf("a.hsel" if save else "s.hsel", (R, d))
        ops.gather_rows(h, m.row_of_text, hsel)
        logits = self.buf("a.logits" if save else "s.logits", (R, cfg.vocab), torch.float32)
        ops.gemm(hsel, lm_w, out=logits)
        logps, per_tok, lse_v = ops.logps_fwd(logits, m.target, m.n_seq, weight=ddpo_weight)
        if save:
            self._saved = dict(m=m, feats=feats, x_last=x, lse_v=lse_v, ddpo_weight=ddpo_weight)
        if save or self.force_logit_means:
            # TRL's `logits/chosen|rejected` = mean of the full [B,S,V] logits = dot(colsum(h), colsum(W_lm)) / (B*S*V)  (K19)
            # (packed rows: the mean runs over the attended positions only -- the reference also averages the logits of its
            # padding positions, which a packed batch never computes; the one metric that differs)
            (c0, c1), (r0, r1) = m.chosen_rows, m.rejected_rows   # shared-prefix rows: the prefixes belong to both halves
            cs = self.buf("m.colsum", (3, d), torch.float32)
            ops.colsum_f32(h[c0:c1], cs[0]); ops.colsum_f32(h[r0:r1], cs[1]); ops.colsum_f32(lm_w, cs[2])
            self.logit_means = self.buf("m.logit_means", (2,), torch.float32)
            inv_c, inv_r = 1.0 / (float(max(c1 - c0, 1)) * cfg.vocab), 1.0 / (float(max(r1 - r0, 1)) * cfg.vocab)
            ops.dot_f32(cs[0], cs[2], inv_c, self.logit_means[0:1]); ops.dot_f32(cs[1], cs[2], inv_r, self.logit_means[1:2])
        return logps

    def _head_backward(self, grad_logps: torch.Tensor, norm_w: torch.Tensor, lm_w: torch.Tensor, g_norm: torch.Tensor,
                       g_lm: Optional[torch.Tensor], acc: bool = False) -> torch.Tensor:
        """d(log-probs) -> gradient of the last decoder layer's output (bf16 [T, d]); g_lm=None: lm_head is frozen."""
        cfg, sv = self.cfg, self._saved
        m = sv["m"]
        d, T = cfg.hidden, m.T
        R = m.n_seq * (m.L - 1)
        logits, hsel = self._bufs["a.logits"], self._bufs["a.hsel"]
        dlogits = self.buf("b.dlogits", (R, cfg.vocab))
        ops.logps_bwd(logits, m.target, m.n_seq, sv["lse_v"], grad_logps, weight=sv["ddpo_weight"], out=dlogits)
        if g_lm is not None:
            ops.gemm(dlogits, hsel, a_kmajor=False, b_kmajor=False, out=g_lm, accumulate=acc)   # dW = dlogits^T hsel
        dhsel = self.buf("b.dhsel", (R, d))
        ops.gemm(dlogits, lm_w, b_kmajor=False, out=dhsel)                                 # dh = dlogits W
        dxf = self.buf("b.dxf", (T, d))
        ops.zero_(dxf)
        if m.shared:   # a shared prefix row feeds the heads of BOTH sequences of its pair: chosen half written, rejected half added
            hr = R // 2
            ops.scatter_rows(dhsel[:hr], m.row_of_text[:hr], dxf)
            ops.scatter_add_rows(dhsel[hr:], m.row_of_text[hr:], dxf)
        else:
            ops.scatter_rows(dhsel, m.row_of_text, dxf)
        dx = self.buf("b.dx0", (T, d))
        ops.rmsnorm_bwd(dxf, sv["x_last"], norm_w, self._bufs["a.rstd_f"], g_norm, out=dx, dw_accumulate=acc)
        return dx

    # ------------------------------------------------------------------ backward of the policy copy
    def _backward(self, grad_logps: torch.Tensor, accumulate: bool = False):
        """accumulate: add this micro-batch's gradients to the gradient arena (gradient_accumulation_steps > 1) instead
        of overwriting it -- every weight-gradient GEMM / reduction takes its `accumulate` epilogue."""
        acc = bool(accumulate)
        self.wait_optimizer()  # the gradient buffer is about to be written
        cfg, w, g = self.cfg, self.policy, self.g
        sv = self._saved
        m: ops.MergeIndex = sv["m"]
        d, T = cfg.hidden, m.T
        H, KV, dh = cfg.heads, cfg.kv_heads, cfg.head_dim
        hd, kvd = H * dh, KV * dh
        dx = self._head_backward(grad_logps, w["norm"], w["lm_head"], g["norm"], g["lm_head"], acc)
        dxf = self._bufs["b.dxf"]
        dx2 = self.buf("b.dx1", (T, d))
        self._reduce_bucket(self.layout.offsets["norm"], self.layout.size)          # norm + lm_head gradients are final
        h = self.buf("s.h", (T, d))
        act = self.buf("s.act", (T, cfg.ff))
        dact = None if self.fuse_swiglu_bwd else self.buf("b.dact", (T, cfg.ff))
        dnorm = dxf  # reuse: [T, d] scratch for the gradients of the normed activations
        dqkv = self.buf("b.dqkv", (T, cfg.qkv_dim))
        datt = self.buf("b.datt", (T, hd))
        delta = self.buf("b.delta", (m.n_attn_seq, H, m.S), torch.float32)
        scale = 1.0 / math.sqrt(dh)
        for i in reversed(range(cfg.layers)):
            x_in = self._bufs[f"x.{i}"]
            if self.tc.activation_checkpointing:  # recompute the layer's saved tensors from its fp32 input
                sb = self._layer_bufs("s", "", m)
                self._layer_fwd(w, i, x_in, sb, m, None)
            else:
                sb = self._layer_bufs("a", f".{i}", m)
            xmid, gu, qkv, att = (sb[k] for k in ("xmid", "gu", "qkv", "att"))
            rstd1, rstd2, lse = (sb[k] for k in ("rstd1", "rstd2", "lse"))
            # ---- MLP
            ops.rmsnorm_fwd(xmid, w[f"L{i}.ln2"], cfg.rms_eps, out=h)                         # recompute h2
            if self.fuse_swiglu_bwd:   # dact = dx Wd never reaches HBM: SwiGLU backward (+ the act recompute) in its epilogue
                ops.gemm_swiglu_bwd(dx, w[f"L{i}.wd"], gu, act)                               # gu <- dgu (in place), act recomputed
                ops.gemm(dx, act, a_kmajor=False, b_kmajor=False, out=g[f"L{i}.wd"], accumulate=acc)          # dWd = dx^T act
            else:
                ops.swiglu_fwd(gu, act)                                                       # recompute act
                ops.gemm(dx, act, a_kmajor=False, b_kmajor=False, out=g[f"L{i}.wd"], accumulate=acc)          # dWd = dx^T act
                ops.gemm(dx, w[f"L{i}.wd"], b_kmajor=False, out=dact)                         # dact = dx Wd
                ops.swiglu_bwd(gu, dact, out=gu)                                              # dgu (in place)
            ops.gemm(gu, h, a_kmajor=False, b_kmajor=False, out=g[f"L{i}.wgu"], accumulate=acc)               # dWgu = dgu^T h2
            ops.gemm(gu, w[f"L{i}.wgu"], b_kmajor=False, out=dnorm)                           # dh2 = dgu Wgu
            ops.rmsnorm_bwd(dnorm, xmid, w[f"L{i}.ln2"], rstd2, g[f"L{i}.ln2"], dres=dx, out=dx2,
                            dw_accumulate=acc)                                                # dxmid
            # ---- attention
            ops.gemm(dx2, att, a_kmajor=False, b_kmajor=False, out=g[f"L{i}.wo"], accumulate=acc)             # dWo = dxmid^T att
            ops.gemm(dx2, w[f"L{i}.wo"], b_kmajor=False, out=datt)                            # datt = dxmid Wo
            attn_backward(m, qkv[:, :hd], qkv[:, hd:hd + kvd], qkv[:, hd + kvd:], att, datt, lse, delta,
                          dqkv[:, :hd], dqkv[:, hd:hd + kvd], dqkv[:, hd + kvd:], H, KV, dh, scale)
            ops.rope_(dqkv, m.pos, self.rope_cos, self.rope_sin, H + KV, dh, inverse=True)
            ops.rmsnorm_fwd(x_in, w[f"L{i}.ln1"], cfg.rms_eps, out=h)                         # recompute h1
            ops.gemm(dqkv, h, a_kmajor=False, b_kmajor=False, out=g[f"L{i}.wqkv"], accumulate=acc)            # dWqkv = dqkv^T h1
            ops.gemm(dqkv, w[f"L{i}.wqkv"], b_kmajor=False, out=dnorm)                        # dh1 = dqkv Wqkv
            ops.rmsnorm_bwd(dnorm, x_in, w[f"L{i}.ln1"], rstd1, g[f"L{i}.ln1"], dres=dx2, out=dx,
                            dw_accumulate=acc)
            self._reduce_bucket(self.layout.offsets[f"L{i}.ln1"],
                                self.layout.offsets[f"L{i + 1}.ln1"] if i + 1 < cfg.layers else self.layout.offsets["norm"])
        # ---- embedding / merge / projector
        feats = sv["feats"]
        nimg = feats.shape[0]
        dimg = self.buf("b.dimg", (nimg, d))
        if acc:   # the fp32 scatter target starts from the accumulated bf16 embedding gradient
            ops.cast_bf16_to_f32(g["embed"].view(-1), self.dembed_f32.view(-1))
        else:
            ops.zero_(self.dembed_f32)
        plan = self._anyres
        if plan is None:
            ops.llava_merge_bwd(m, dx, self.dembed_f32, dimg)
        else:
            dpacked = self.buf("b.dpacked", (plan.total_feats, d))
            ops.llavanext_merge_bwd(m, dx, self.dembed_f32, dpacked)
            ops.zero_(dimg)                                    # crop rows cut away by the unpadding get no gradient
            ops.scatter_rows(dpacked, plan.scatter_index, dimg)
            dnl = self.buf("b.dnewline", (plan.newline_rows.numel(), d))
            ops.gather_rows(dpacked, plan.newline_rows, dnl)
            ops.colsum(dnl, g["image_newline"], accumulate=acc)
        ops.cast_f32_to_bf16(self.dembed_f32.view(-1), g["embed"].view(-1))
        ph, z = self._bufs["p.h"], self._bufs["p.z"]
        ops.gemm(dimg, ph, a_kmajor=False, b_kmajor=False, out=g["proj.w2"], accumulate=acc)
        ops.colsum(dimg, g["proj.b2"], accumulate=acc)
        dph = self.buf("b.dph", (nimg, d))
        ops.gemm(dimg, w["proj.w2"], b_kmajor=False, out=dph)
        ops.gelu_bwd(z, dph, out=dph)
        ops.gemm(dph, feats, a_kmajor=False, b_kmajor=False, out=g["proj.w1"], accumulate=acc)
        ops.colsum(dph, g["proj.b1"], accumulate=acc)
        self._reduce_bucket(0, self.layout.offsets["L0.ln1"])                         # projector + embedding

    # ------------------------------------------------------------------ optimizer + data parallel
    def _dist_on(self) -> bool:
        return torch.distributed.is_available() and torch.distributed.is_initialized() and self.world_size() > 1

    def _reduce_bucket(self, lo: int, hi: int):
        """Bucketed all-reduce overlapped with backward: the gradients in [lo, hi) of the flat buffer are final, so
        their sum-reduction starts now on the NCCL stream (ordered after the kernels already enqueued)."""
        if self.overlap_allreduce and self._dist_on() and hi > lo:
            self._pending.append(torch.distributed.all_reduce(self.grads[lo:hi], op=torch.distributed.ReduceOp.SUM,
                                                              group=self.pg, async_op=True))

    def allreduce_grads(self):
        """Gradient all-reduce over the flat bf16 buffer (sum; the 1/world scale is folded into AdamW): either the
        per-layer buckets launched during backward are awaited, or one all-reduce covers the whole buffer."""
        if not self._dist_on():
            return
        if self._pending:
            for h in self._pending:
                h.wait()
            self._pending = []
        elif self.shard_optimizer:
            # in place: rank r keeps the sum of slice r (only that slice is read by the sharded AdamW)
            torch.distributed.reduce_scatter_tensor(self.grads[self.shard_lo:self.shard_hi], self.grads,
                                                    op=torch.distributed.ReduceOp.SUM, group=self.pg)
        else:
            torch.distributed.all_reduce(self.grads, op=torch.distributed.ReduceOp.SUM, group=self.pg)

    def world_size(self) -> int:
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            return torch.distributed.get_world_size(self.pg)
        return 1

    def optimizer_step(self, lr: Optional[float] = None):
        """One clipped AdamW step over this rank's slice of the flat buffers.  `lr`: the learning rate of this step when an
        outer scheduler owns it (plugin.B200FlatAdamW under the HF Trainer); default TrainConfig.lr_at(step index)."""
        tc = self.tc
        if lr is None:
            lr = tc.lr_at(self.opt_step)
        self.last_lr = float(lr)
        self.opt_step += 1
        lo, hi = self.shard_lo, self.shard_hi
        g, p = self.grads[lo:hi], self.params[lo:hi]
        ops.sumsq(g, self.grad_sumsq, self.sumsq_ws)
        if self.shard_optimizer:  # global gradient norm = sum of the slices' squared norms (one fp32 scalar)
            torch.distributed.all_reduce(self.grad_sumsq, op=torch.distributed.ReduceOp.SUM, group=self.pg)
        if self.device.type == "cuda":
            self.norm_done.record()  # the logged grad_norm is final here; AdamW only reads it
        ops.adamw_(p, g, self.master, self.exp_avg, self.exp_avg_sq, float(lr), tc.adam_beta1,
                   tc.adam_beta2, tc.adam_eps, tc.weight_decay, self.opt_step, grad_scale=1.0 / self.world_size(),
                   grad_sumsq=self.grad_sumsq, max_grad_norm=tc.max_grad_norm)
        if self.shard_optimizer:
            torch.distributed.all_gather_into_tensor(self.params, p, group=self.pg)

    def reduce_and_step(self, lr: Optional[float] = None):
        """Gradient reduction over the data-parallel ranks + AdamW (+ parameter all-gather when sharded); deferred to the
        side stream when `async_optimizer` (the next step's reference pass overlaps it).  -> the device scalar that will hold
        the squared global gradient norm (None without an optimizer)."""
        if self.async_optimizer and self.with_optimizer:
            main = torch.cuda.current_stream(self.device)
            self.opt_stream.wait_stream(main)            # gradients are final
            with torch.cuda.stream(self.opt_stream):
                self.allreduce_grads()
                self.optimizer_step(lr)
                self.opt_done.record()
            self._opt_pending = True
            return self.grad_sumsq
        self.allreduce_grads()
        if self.with_optimizer:
            self.optimizer_step(lr)
            return self.grad_sumsq
        return None

    # ------------------------------------------------------------------ the step
    def prepare_inputs(self, input_ids: torch.Tensor, attention_mask: torch.Tensor, labels: torch.Tensor,
                       pixel_values: torch.Tensor, ddpo_weight: Optional[torch.Tensor] = None,
                       image_sizes: Optional[torch.Tensor] = None, imgs_per_seq: int = 1):
        """Host -> device staging of one concatenated batch (the H2D copies of the step).  Inputs may be CPU
        (pinned or not) or CUDA tensors.  pixel_values may be [B,...] or the reference's duplicated [2B,...]; with
        `imgs_per_seq` = k > 1 (LLaVA-1.5: every sequence holds k <image> placeholders, Llava/__init__.py:44-47) it is
        [B*k,...] pair-major (or the duplicated [2*B*k,...]).
        LLaVA-Next: pixel_values is the processor's [B, max_crops, 3, H, W] (or the flat [sum crops, 3, H, W]) and
        `image_sizes` [B, 2] is required; the returned tuple then carries a 6th element, the anyres plan."""
        dev = self.device
        cfg = self.cfg
        n_seq = input_ids.shape[0]
        plan = None
        from . import host
        host.validate_token_batch(input_ids, labels, cfg.vocab, cfg.image_token_index,
                                  imgs_per_seq if cfg.family != "llava_next" else 1, self.tc.label_pad_token_id)
        if cfg.family == "llava_next":
            if image_sizes is None:
                raise ValueError("LLaVA-Next needs image_sizes (LlavaNext/__init__.py:211-218)")
            if image_sizes.shape[0] == n_seq:
                image_sizes = image_sizes[: n_seq // 2]
            plan = host.anyres_pack_index(image_sizes.cpu(), cfg.image_grid_pinpoints, cfg.image_size, cfg.patch_size)
            plan.merged_len = host.next_merged_len(input_ids, attention_mask, plan.feature_lens, cfg.image_token_index)
            if pixel_values.dim() == 5:  # stacked crops: keep each image's real crops (:220-225)
                if pixel_values.shape[0] == n_seq:
                    pixel_values = pixel_values[: n_seq // 2]
                if all(c == pixel_values.shape[1] for c in plan.crops):
                    pixel_values = pixel_values.reshape(-1, *pixel_values.shape[2:])
                else:
                    pixel_values = torch.cat([pv[:c] for pv, c in zip(pixel_values, plan.crops)], dim=0)
            elif pixel_values.dim() != 4:
                raise ValueError(f"pixel_values of shape {pixel_values.shape}, expect to be of 4 or 5 dimensions")
            elif pixel_values.shape[0] == 2 * sum(plan.crops):
                pixel_values = pixel_values[: sum(plan.crops)]
            if pixel_values.shape[0] != sum(plan.crops):
                raise ValueError(f"{pixel_values.shape[0]} crops given, image_sizes imply {sum(plan.crops)}")
            plan.to(dev)
        else:
            if pixel_values.shape[0] == n_seq * imgs_per_seq:  # concatenated_inputs duplicated the images ([v, v], trainer.py:135-145)
                pixel_values = pixel_values[: (n_seq // 2) * imgs_per_seq]
            if imgs_per_seq != 1 and pixel_values.shape[0] != (n_seq // 2) * imgs_per_seq:
                raise ValueError(f"{pixel_values.shape[0]} images for {n_seq // 2} pairs of {imgs_per_seq} images each")
        ids = input_ids.to(dev, non_blocking=True).contiguous()
        am = attention_mask.to(dev, non_blocking=True).contiguous()
        lb = labels.to(dev, non_blocking=True).contiguous()
        px = pixel_values.to(dev, non_blocking=True).contiguous()
        if px.dtype not in (torch.float32, torch.bfloat16):
            px = px.float()
        wt = ddpo_weight.to(dev, non_blocking=True).reshape(-1).contiguous() if ddpo_weight is not None else None
        return (ids, am, lb, px, wt) if plan is None else (ids, am, lb, px, wt, plan)

    def forward_logps(self, ids, am, lb, px, ddpo_weight=None, anyres=None, which: str = "policy", save: bool = False,
                      feats: Optional[torch.Tensor] = None, m: Optional["ops.MergeIndex"] = None, seq_lens=None,
                      imgs_per_seq: int = 1, prefix_rows=None):
        """`seq_lens` (host ints, merged length of every sequence; host.merged_seq_lens) is only read with
        TrainConfig.pack_sequences; without it the lengths are read back from the device (one synchronisation)."""
        cfg = self.cfg
        if (cfg.family == "llava_next") != (anyres is not None):
            raise ValueError("the anyres plan from prepare_inputs is required for (and only for) LLaVA-Next")
        self._anyres = anyres
        if m is None:
            if anyres is not None:
                if imgs_per_seq != 1:
                    raise ValueError("LLaVA-Next: one image per sequence (several anyres images per sequence are not implemented)")
                m = ops.llavanext_merge_index(ids, am, lb, anyres.feat_off, anyres.total_feats, anyres.merged_len,
                                              len(anyres.crops), imgs_per_seq, cfg.image_token_index, cfg.ignore_index)
            else:
                if px.shape[0] % imgs_per_seq:
                    raise ValueError(f"{px.shape[0]} images do not split into groups of {imgs_per_seq}")
                m = ops.llava_merge_index(ids, am, lb, cfg.n_patches, px.shape[0] // imgs_per_seq, imgs_per_seq,
                                          cfg.image_token_index, cfg.pad_token_id, cfg.ignore_index)
            if self.tc.share_prefix:     # one copy of every pair's common prefix (implies packed rows)
                if anyres is not None:
                    raise ValueError("share_prefix: LLaVA-Next (variable packed feature lengths) is not supported yet")
                if seq_lens is None or prefix_rows is None:
                    raise ValueError("share_prefix needs the host-side row plan: pass **engine.host_row_plan(ids, am) "
                                     "(train_step and plugin.concatenated_forward do)")
                ops.share_prefix_rows(m, seq_lens, prefix_rows)
                self._pad_rows, self._cur_rows = m.n_seq * m.S, m.T
            elif self.tc.pack_sequences:   # drop the padding rows: every kernel below runs over sum(len) rows
                ops.pack_merge_rows(m, seq_lens if seq_lens is not None else m.seqlens.cpu().tolist())
                self._pad_rows, self._cur_rows = m.n_seq * m.S, m.T
        self.ensure_rope_len(m.S)
        if feats is None:
            feats = self.vision_features(px)
        if which == "policy":
            self.wait_optimizer()
        w = self.policy if which == "policy" else self.ref
        return self._forward(w, m, feats, which, save, ddpo_weight), m, feats

    def step(self, ids, am, lb, px, ddpo_weight=None, anyres=None, train: bool = True,
             ref_logps: Optional[torch.Tensor] = None, seq_lens=None, imgs_per_seq: int = 1,
             accumulate: bool = False, sync: bool = True, loss_scale: float = 1.0, prefix_rows=None) -> StepOutput:
        """One DPO step on device-resident inputs: policy fwd, reference fwd (no grad), loss, and when `train`
        backward + gradient all-reduce + AdamW.  `ref_logps` ([2B] fp32, chosen then rejected) replaces the reference
        pass: TRL's precompute_ref_log_probs branch of get_batch_loss_metrics (plumbed at base/trainer.py:61,96 and
        the collator's `_logps` keys, base/collator.py:62-64).
        Gradient accumulation: `accumulate` adds this micro-batch's gradients (scaled by `loss_scale` = 1/k) to the arena,
        `sync=False` stops before the reduction + optimizer (HF Trainer's no_sync micro-steps)."""
        tc = self.tc
        if ref_logps is not None:
            pol, m, feats = self.forward_logps(ids, am, lb, px, ddpo_weight, anyres, "policy", save=train, seq_lens=seq_lens,
                                               **({"imgs_per_seq": imgs_per_seq} if imgs_per_seq != 1 else {}),
                                               **({"prefix_rows": prefix_rows} if prefix_rows is not None else {}))
            ref = ref_logps.to(self.device, torch.float32).reshape(-1).contiguous()
            if ref.numel() != pol.numel():
                raise ValueError(f"ref_logps holds {ref.numel()} values, the batch has {pol.numel()} sequences")
        else:
            # reference pass first: it reads none of the policy weights, so the previous step's deferred optimizer
            # (side stream) overlaps it; the policy pass below waits for `opt_done`
            ref, m, feats = self.forward_logps(ids, am, lb, px, ddpo_weight, anyres, "ref", save=False, seq_lens=seq_lens,
                                               **({"imgs_per_seq": imgs_per_seq} if imgs_per_seq != 1 else {}),
                                               **({"prefix_rows": prefix_rows} if prefix_rows is not None else {}))
            pol, _, _ = self.forward_logps(ids, am, lb, px, ddpo_weight, anyres, "policy", save=train, feats=feats, m=m)
        losses, cr, rr, stats, grad = ops.dpo_loss(pol, ref, tc.beta, tc.label_smoothing, tc.loss_type, tc.reference_free,
                                                   float(loss_scale), want_grad=train)
        out = StepOutput()
        out.losses, out.chosen_rewards, out.rejected_rewards, out.stats = losses, cr, rr, stats
        out.policy_logps, out.ref_logps = pol, ref
        out.loss = stats[0:1]
        out.grad_norm = None
        if train:
            self._backward(grad, accumulate=accumulate)
            if sync:
                out.grad_norm = self.reduce_and_step()
        return out

    def train_step(self, batch: Dict, train: bool = True) -> Dict[str, float]:
        """Public end-to-end call: one collated host batch (the dict VLDPODataCollatorWithPadding emits) in,
        the TRL metric dict out.  Includes the H2D staging of the inputs and the D2H read of the results."""
        from . import host
        tc = self.tc
        cb = host.concatenated_inputs(batch, False, tc.label_pad_token_id, tc.padding_value)
        ids, am, lb = host.right_pad_valid_tokens(cb["concatenated_input_ids"], cb["concatenated_attention_mask"],
                                                  cb["concatenated_labels"], tc.padding_value, tc.label_pad_token_id,
                                                  tc.loss_type)
        px = batch["img_input_dict"]["pixel_values"]  # one copy per pair: the [v, v] duplicate is never shipped
        sizes = batch["img_input_dict"].get("image_sizes")
        wt = None
        if tc.loss_type == "ddpo":
            wt = self.ddpo_weights(ids, am, lb, sizes)
        ref_logps = None
        if "reference_chosen_logps" in batch and "reference_rejected_logps" in batch:  # precompute_ref_log_probs
            ref_logps = torch.cat([torch.as_tensor(batch["reference_chosen_logps"], dtype=torch.float32).reshape(-1),
                                   torch.as_tensor(batch["reference_rejected_logps"], dtype=torch.float32).reshape(-1)])
        k = self.images_per_sequence(batch)
        kw = {"imgs_per_seq": k} if k != 1 else {}
        plan = self.host_row_plan(ids, am, sizes, **kw)
        ga = max(1, int(tc.gradient_accumulation_steps)) if train else 1
        micro = self._micro_step % ga
        sync = micro == ga - 1
        if train:
            self._micro_step += 1
        out = self.step(*self.prepare_inputs(ids, am, lb, px, wt, sizes, **kw), train=train, ref_logps=ref_logps,
                        accumulate=micro > 0, sync=sync, loss_scale=1.0 / ga, **plan, **kw)
        n = out.policy_logps.numel() // 2
        if self._opt_pending and out.grad_norm is not None:  # grad_norm comes from the side stream; AdamW itself keeps running behind this read
            torch.cuda.current_stream(self.device).wait_event(self.norm_done)
        packed = torch.cat([out.stats, out.policy_logps[:n].mean()[None], out.policy_logps[n:].mean()[None],
                            (out.grad_norm if out.grad_norm is not None else out.stats[:1] * 0),
                            (self.logit_means if train else out.stats[:2] * 0)]).cpu()  # the D2H read
        world = self.world_size()
        return {"loss": float(packed[0]), "rewards/accuracies": float(packed[1]), "rewards/chosen": float(packed[2]),
                "rewards/rejected": float(packed[3]), "rewards/margins": float(packed[4]),
                "logps/chosen": float(packed[6]), "logps/rejected": float(packed[7]),
                "logits/chosen": float(packed[9]), "logits/rejected": float(packed[10]),
                "grad_norm": float(packed[8]) ** 0.5 / world}

    def images_per_sequence(self, batch: Dict) -> int:
        """k for a collated batch whose img_input_dict.pixel_values holds k images per pair, pair-major ([B*k, 3, H, W]; the DPO
        collator emits k = 1, models/Llava/__init__.py:435-443).  LLaVA-1.5 family only: the other families' pixel layouts carry
        their own structure (anyres crops, file names in the token stream)."""
        px = batch["img_input_dict"]["pixel_values"]
        n_pairs = batch["chosen_input_ids"].shape[0]
        if self.cfg.family != "llava" or px.dim() != 4 or px.shape[0] == n_pairs:
            return 1
        if px.shape[0] % n_pairs:
            raise ValueError(f"{px.shape[0]} images do not split over {n_pairs} pairs")
        return px.shape[0] // n_pairs

    def compute_reference_log_probs(self, batch: Dict) -> Tuple[torch.Tensor, torch.Tensor]:
        """trl 0.8.1 DPOTrainer.compute_reference_log_probs (the producer half of precompute_ref_log_probs, base/trainer.py:61,96):
        one no-grad reference pass over a collated batch -> (reference_chosen_logps, reference_rejected_logps) on the host,
        ready to be stored with the dataset and fed back through the batch keys of the same names (base/collator.py:62-64),
        which makes `train_step` skip its reference pass."""
        from . import host
        tc = self.tc
        cb = host.concatenated_inputs(batch, False, tc.label_pad_token_id, tc.padding_value)
        ids, am, lb = host.right_pad_valid_tokens(cb["concatenated_input_ids"], cb["concatenated_attention_mask"],
                                                  cb["concatenated_labels"], tc.padding_value, tc.label_pad_token_id,
                                                  tc.loss_type)
        sizes = batch["img_input_dict"].get("image_sizes")
        wt = self.ddpo_weights(ids, am, lb, sizes) if tc.loss_type == "ddpo" else None
        k = self.images_per_sequence(batch)
        kw = {"imgs_per_seq": k} if k != 1 else {}
        plan = self.host_row_plan(ids, am, sizes, **kw)
        inputs = self.prepare_inputs(ids, am, lb, batch["img_input_dict"]["pixel_values"], wt, sizes, **kw)
        logps, _, _ = self.forward_logps(*inputs, which="ref", save=False, **plan, **kw)
        logps = logps.float().cpu()
        n = logps.numel() // 2
        return logps[:n], logps[n:]

    def host_row_plan(self, ids, am, image_sizes=None, imgs_per_seq: int = 1) -> Dict:
        """Keyword arguments for step / forward_logps that describe the row layout of one concatenated HOST batch, so the
        step needs no device read-back: {} (padded rows), {seq_lens} (TrainConfig.pack_sequences) or {seq_lens, prefix_rows}
        (TrainConfig.share_prefix: merged rows the chosen and rejected sequence of every pair have in common)."""
        tc = self.tc
        kw = {"imgs_per_seq": imgs_per_seq} if imgs_per_seq != 1 else {}
        if not (tc.pack_sequences or tc.share_prefix):
            return {}
        plan = {"seq_lens": self.host_seq_lens(ids, am, image_sizes, **kw)}
        if tc.share_prefix:
            from . import host
            fam = self.cfg.family
            if fam in ("llava", "xc2"):     # one <image> token stands for n_patches merged rows
                plan["prefix_rows"] = host.shared_prefix_rows(ids, am, self.cfg.image_token_index, self.cfg.n_patches)
            elif fam == "qwen_vl":          # the 256 image rows are placeholder tokens of the text itself (S == L)
                plan["prefix_rows"] = host.shared_prefix_rows(ids, am, -1, 1)
            else:
                raise ValueError(f"share_prefix is not implemented for {fam!r} (LLaVA-Next: variable packed feature lengths)")
        return plan

    def host_seq_lens(self, ids, am, image_sizes=None, imgs_per_seq: int = 1) -> List[int]:
        """Merged length of every sequence of one concatenated host batch (packed steps: the rows that survive)."""
        from . import host
        cfg = self.cfg
        if cfg.family != "llava_next":
            return host.merged_seq_lens(ids, am, cfg.image_token_index, cfg.n_patches * imgs_per_seq)
        n_seq = ids.shape[0]
        if image_sizes.shape[0] == n_seq:
            image_sizes = image_sizes[: n_seq // 2]
        plan = host.anyres_pack_index(image_sizes.cpu(), cfg.image_grid_pinpoints, cfg.image_size, cfg.patch_size)
        return host.merged_seq_lens(ids, am, cfg.image_token_index,
                                    [plan.feature_lens[b % len(plan.feature_lens)] for b in range(n_seq)])

    def ddpo_weights(self, ids, am, lb, image_sizes=None) -> torch.Tensor:
        """Host-side DDPO row weights of one concatenated batch (trainer.py:169-184 over the merged label layout)."""
        from . import host
        cfg, tc = self.cfg, self.tc
        if cfg.family != "llava_next":
            return host.ddpo_row_weights_native(ids, lb, cfg.image_token_index, cfg.n_patches, tc.label_pad_token_id)
        n_seq = ids.shape[0]
        if image_sizes.shape[0] == n_seq:
            image_sizes = image_sizes[: n_seq // 2]
        plan = host.anyres_pack_index(image_sizes.cpu(), cfg.image_grid_pinpoints, cfg.image_size, cfg.patch_size)
        per_seq = [plan.feature_lens[b % len(plan.feature_lens)] for b in range(n_seq)]
        S = host.next_merged_len(ids, am, plan.feature_lens, cfg.image_token_index)
        return host.ddpo_row_weights_native(ids, lb, cfg.image_token_index, per_seq, tc.label_pad_token_id,
                                            attention_mask=am, merged_len=S)

    def check_merge_status(self, m: "ops.MergeIndex"):
        """Synchronising validity check mirroring the reference's ValueError (Llava/__init__.py:90-94)."""
        st = int(m.status.item())
        if st == 1 or st == 2:
            raise ValueError("The input provided to the model are wrong. The number of image tokens does not match the "
                             "number of images given to the model (every sequence must hold the same number of <image> "
                             "tokens in this build).")
        if st == 3:
            raise ValueError("attention_mask must be a right-padded prefix mask (left padding is not supported yet)")
        if st == 4:
            raise ValueError("an <image> placeholder lies outside the attended prefix of its sequence")
