"""The B200-native DPO step for Qwen-VL with LoRA (BASELINE.json configs[2]).

Replaces, behind the same engine interface as engine.LlavaDPOEngine:
  * `QWenLMHeadModel.forward` / `QWenModel.forward`      (models/QwenVL/modeling_qwen.py:509-699, 800-855)
  * `VisionTransformer.forward` + `Resampler.forward`     (models/QwenVL/visual.py:393-415, 140-152)
  * peft's LoRA linear on c_attn / attn.c_proj / w1 / w2  (scripts/dpo_qwenvl.sh; reference pass = adapters disabled)
  * the trainer-side functions shared with LLaVA (concatenated_forward, get_batch_logps, dpo_loss, TRL metrics).

What is different from the LLaVA engine:
  * ONE frozen bf16 copy of the base model serves both passes (policy = base + adapters, reference = base); only the
    adapters are trainable, so the flat trainable/gradient/optimizer arenas hold 112 M parameters instead of 7 B, the
    backward runs no base weight-gradient GEMM (2/3 of the full-FT backward FLOPs) and nothing flows into embeddings or
    the vision tower (frozen: `--freeze_vision_tower True`; peft freezes the resampler too).
  * LoRA linear: ts = bf16(s * x A^T) -> y = x W^T + b + ts B^T in ONE launch (the adapter term is a second operand pair
    contracted into the base GEMM's accumulator, `vlb200_gemm_bf16_ex`: no [T, out] intermediate in HBM, one rounding);
    backward: dB = dy^T ts, dt = bf16(s * dy B), dA = dt^T x, dx = dy W + dt A (again one launch, two operand pairs).
  * ViT-bigG: per-head interleaved q|k|v projection re-laid out at load time as [q heads | k heads | v heads] with each
    104-wide head zero-padded to 128 (q.k and the context are unchanged), so the tcgen05 attention kernel runs as is;
    positional tables are interpolated once on the host (visual.py:24-45 is a weight transform for a fixed image size).
  * Resampler: its 256 queries do not depend on the image -> q is a constant table; the cross-attention runs on the
    self-attention kernel with the query block padded to the key length (rows >= n_queries are discarded).
  * the image features overwrite the placeholder positions of the text (S == L): `vlb200_qwen_merge_index`.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from . import ops
from .config import QwenModelConfig, qwen_lora_specs, qwen_weight_specs, tensor_seed
from .engine import Arena, LlavaDPOEngine, Weights, attn_backward, attn_forward


def _lora_layout(cfg: QwenModelConfig) -> Arena:
    a = Arena()
    d, r = cfg.hidden, cfg.lora_r
    for i in range(cfg.layers):
        a.add(f"L{i}.qkv.A", (r, d)); a.add(f"L{i}.qkv.B", (3 * d, r))
        a.add(f"L{i}.o.A", (r, d)); a.add(f"L{i}.o.B", (d, r))
        # w2 (the silu branch) first: it pairs with the first half of the fused gate|up weight
        a.add(f"L{i}.gu.A", (2 * r, d)); a.add(f"L{i}.w2.B", (cfg.ff, r)); a.add(f"L{i}.w1.B", (cfg.ff, r))
    return a


def _base_layout(cfg: QwenModelConfig) -> Arena:
    a = Arena()
    d = cfg.hidden
    a.add("embed", (cfg.vocab, d))
    for i in range(cfg.layers):
        a.add(f"L{i}.ln1", (d,)); a.add(f"L{i}.wqkv", (3 * d, d)); a.add(f"L{i}.bqkv", (3 * d,)); a.add(f"L{i}.wo", (d, d))
        a.add(f"L{i}.ln2", (d,)); a.add(f"L{i}.wgu", (2 * cfg.ff, d)); a.add(f"L{i}.wd", (d, cfg.ff))
    a.add("norm", (d,)); a.add("lm_head", (cfg.vocab, d))
    return a


def _visual_layout(cfg: QwenModelConfig) -> Arena:
    a = Arena()
    w, d, H, hp = cfg.v_width, cfg.hidden, cfg.v_heads, cfg.v_head_pad
    P, Q = cfg.n_patches, cfg.n_queries
    a.add("v.patch", (w, cfg.patch_k_padded)); a.add("v.pos", (P, w)); a.add("v.pre.w", (w,)); a.add("v.pre.b", (w,))
    for i in range(cfg.v_layers):
        a.add(f"v{i}.ln1.w", (w,)); a.add(f"v{i}.ln1.b", (w,)); a.add(f"v{i}.wqkv", (3 * H * hp, w)); a.add(f"v{i}.bqkv", (3 * H * hp,))
        a.add(f"v{i}.wo", (w, H * hp)); a.add(f"v{i}.bo", (w,)); a.add(f"v{i}.ln2.w", (w,)); a.add(f"v{i}.ln2.b", (w,))
        a.add(f"v{i}.w1", (cfg.v_mlp, w)); a.add(f"v{i}.b1", (cfg.v_mlp,)); a.add(f"v{i}.w2", (w, cfg.v_mlp)); a.add(f"v{i}.b2", (w,))
    a.add("r.wkv", (d, w)); a.add("r.lnkv.w", (d,)); a.add("r.lnkv.b", (d,))
    a.add("r.wk", (d, d)); a.add("r.bk", (d,)); a.add("r.wv", (d, d)); a.add("r.bv", (d,))
    a.add("r.kpos", (P, d))      # (interpolated sincos positions of the keys) @ Wk^T: enters k through the residual slot
    a.add("r.q", (Q, d))         # (ln_q(query) + sincos) @ Wq^T + bq: image independent
    a.add("r.wo", (d, d)); a.add("r.bo", (d,)); a.add("r.post.w", (d,)); a.add("r.post.b", (d,)); a.add("r.proj", (d, d))
    return a


class QwenVLDPOEngine(LlavaDPOEngine):
    has_ref_copy = False
    needs_embed_grad = False

    def _make_layouts(self):
        return _lora_layout(self.cfg), _visual_layout(self.cfg)

    def _alloc_family(self):
        cfg = self.cfg
        if cfg.n_queries > cfg.n_patches:
            raise ValueError("the resampler's query block is padded to the key length: n_queries must be <= n_patches")
        self.blayout = _base_layout(cfg)
        self.bparams = torch.zeros(self.blayout.size, dtype=torch.bfloat16, device=self.device)  # frozen base LM
        self.base = Weights(self.blayout, self.bparams)
        self.extra_state: Dict[str, torch.Tensor] = {}
        self._dw_scratch = torch.zeros(cfg.hidden, dtype=torch.bfloat16, device=self.device)  # frozen norm weights' "gradient"

    # ------------------------------------------------------------------ names
    def lora_views(self, w: Weights) -> Dict[str, torch.Tensor]:
        """peft-style names (`transformer.h.{i}.<module>.lora_A|lora_B`) -> views of the LoRA arena."""
        cfg, r = self.cfg, self.cfg.lora_r
        out: Dict[str, torch.Tensor] = {}
        for i in range(cfg.layers):
            p = f"transformer.h.{i}."
            out[p + "attn.c_attn.lora_A"] = w[f"L{i}.qkv.A"]; out[p + "attn.c_attn.lora_B"] = w[f"L{i}.qkv.B"]
            out[p + "attn.c_proj.lora_A"] = w[f"L{i}.o.A"]; out[p + "attn.c_proj.lora_B"] = w[f"L{i}.o.B"]
            out[p + "mlp.w2.lora_A"] = w[f"L{i}.gu.A"][:r]; out[p + "mlp.w1.lora_A"] = w[f"L{i}.gu.A"][r:]
            out[p + "mlp.w2.lora_B"] = w[f"L{i}.w2.B"]; out[p + "mlp.w1.lora_B"] = w[f"L{i}.w1.B"]
        return out

    def base_views(self) -> Dict[str, torch.Tensor]:
        cfg, b = self.cfg, self.base
        out = {"transformer.wte.weight": b["embed"], "transformer.ln_f.weight": b["norm"], "lm_head.weight": b["lm_head"]}
        for i in range(cfg.layers):
            p = f"transformer.h.{i}."
            out[p + "ln_1.weight"] = b[f"L{i}.ln1"]; out[p + "attn.c_attn.weight"] = b[f"L{i}.wqkv"]
            out[p + "attn.c_attn.bias"] = b[f"L{i}.bqkv"]; out[p + "attn.c_proj.weight"] = b[f"L{i}.wo"]
            out[p + "ln_2.weight"] = b[f"L{i}.ln2"]; out[p + "mlp.w2.weight"] = b[f"L{i}.wgu"][:cfg.ff]
            out[p + "mlp.w1.weight"] = b[f"L{i}.wgu"][cfg.ff:]; out[p + "mlp.c_proj.weight"] = b[f"L{i}.wd"]
        return out

    def hf_state(self, which: str = "policy") -> Dict[str, torch.Tensor]:
        """policy: base + adapters; ref: base (adapters disabled); grad: adapter gradients.  The vision tower is stored
        re-laid-out, its original tensors travel in `extra_state` when loaded from a checkpoint."""
        self.wait_optimizer()
        if which == "grad":
            return self.lora_views(self.g)
        out = dict(self.base_views())
        if which == "policy":
            out.update(self.lora_views(self.policy))
        return out

    # ------------------------------------------------------------------ weights
    def set_visual_tensor(self, name: str, t: torch.Tensor):
        """One tensor of `transformer.visual.*` (reference layout, any float dtype/device) -> the engine's layout."""
        cfg, v = self.cfg, self.vis
        w, d, H, hn, hp = cfg.v_width, cfg.hidden, cfg.v_heads, cfg.v_head_dim, cfg.v_head_pad
        k = name[len("transformer.visual."):]
        t = t.to(self.device, torch.float32)
        bf = lambda x: x.to(torch.bfloat16)  # noqa: E731
        if k == "conv1.weight":
            v["v.patch"][:, :cfg.patch_k].copy_(bf(t.reshape(w, -1)))
        elif k == "positional_embedding":
            from .host import interpolate_pos_table
            v["v.pos"].copy_(bf(interpolate_pos_table(t.to(torch.bfloat16).float().cpu(), cfg.n_patches).to(self.device)))
        elif k in ("ln_pre.weight", "ln_pre.bias"):
            v["v.pre.w" if k.endswith("weight") else "v.pre.b"].copy_(bf(t))
        elif k.startswith("transformer.resblocks."):
            i, rest = k[len("transformer.resblocks."):].split(".", 1)
            if rest in ("attn.in_proj.weight", "attn.in_proj.bias"):
                # rows are (head, [q|k|v], hn) -> (part, head, hp) with hp - hn zero rows per head
                src = t.reshape(H, 3, hn, -1)
                dst = torch.zeros(3, H, hp, src.shape[-1], device=self.device)
                dst[:, :, :hn] = src.permute(1, 0, 2, 3)
                tgt = v[f"v{i}.wqkv"] if rest.endswith("weight") else v[f"v{i}.bqkv"]
                tgt.copy_(bf(dst.reshape(tgt.shape)))
            elif rest == "attn.out_proj.weight":
                dst = torch.zeros(w, H, hp, device=self.device)
                dst[:, :, :hn] = t.reshape(w, H, hn)
                v[f"v{i}.wo"].copy_(bf(dst.reshape(w, H * hp)))
            else:
                key = {"attn.out_proj.bias": "bo", "ln_1.weight": "ln1.w", "ln_1.bias": "ln1.b", "ln_2.weight": "ln2.w",
                       "ln_2.bias": "ln2.b", "mlp.c_fc.weight": "w1", "mlp.c_fc.bias": "b1", "mlp.c_proj.weight": "w2",
                       "mlp.c_proj.bias": "b2"}[rest]
                v[f"v{i}.{key}"].copy_(bf(t))
        elif k.startswith("attn_pool."):
            self._pool_raw[k[len("attn_pool."):]] = t.to(torch.bfloat16).float()  # bf16-representable, as the model holds them
            simple = {"kv_proj.weight": "r.wkv", "ln_kv.weight": "r.lnkv.w", "ln_kv.bias": "r.lnkv.b",
                      "attn.out_proj.weight": "r.wo", "attn.out_proj.bias": "r.bo"}
            if k[len("attn_pool."):] in simple:
                v[simple[k[len("attn_pool."):]]].copy_(bf(t))
        elif k in ("ln_post.weight", "ln_post.bias"):
            v["r.post.w" if k.endswith("weight") else "r.post.b"].copy_(bf(t))
        elif k == "proj":
            v["r.proj"].copy_(bf(t))
        elif k == "attn_pool.pos_embed":
            pass  # deterministic sincos table (visual.py:112-114), rebuilt in finalize_visual
        else:
            raise KeyError(name)

    def finalize_visual(self):
        """Constant tables of the resampler: q = (ln_q(query) + sincos) Wq^T + bq and kpos @ Wk^T (frozen weights, fixed
        image size), computed once on the host in fp32 from the bf16-representable weights."""
        from .host import interpolate_pos_table, sincos_2d
        cfg, v, raw = self.cfg, self.vis, self._pool_raw
        d = cfg.hidden
        need = ("query", "ln_q.weight", "ln_q.bias", "attn.in_proj_weight", "attn.in_proj_bias")
        if any(k not in raw for k in need):
            raise KeyError(f"visual.attn_pool tensors missing: {[k for k in need if k not in raw]}")
        Wi, bi = raw["attn.in_proj_weight"].cpu(), raw["attn.in_proj_bias"].cpu()
        qpos = sincos_2d(d, int(math.sqrt(cfg.n_queries)))
        kpos = interpolate_pos_table(qpos, cfg.n_patches)
        qin = torch.nn.functional.layer_norm(raw["query"].cpu(), (d,), raw["ln_q.weight"].cpu(), raw["ln_q.bias"].cpu(),
                                             cfg.v_eps) + qpos
        dev = self.device
        v["r.q"].copy_((qin @ Wi[:d].T + bi[:d]).to(dev, torch.bfloat16))
        v["r.kpos"].copy_((kpos @ Wi[d:2 * d].T).to(dev, torch.bfloat16))
        v["r.wk"].copy_(Wi[d:2 * d].to(dev, torch.bfloat16)); v["r.bk"].copy_(bi[d:2 * d].to(dev, torch.bfloat16))
        v["r.wv"].copy_(Wi[2 * d:].to(dev, torch.bfloat16)); v["r.bv"].copy_(bi[2 * d:].to(dev, torch.bfloat16))

    def init_synthetic(self, seed: int, ref_alpha: float = 0.0):
        """Seeded random-init base + adapters (bit-identical to oracle.qwen_restate.make_weights)."""
        self.wait_optimizer()
        self._pool_raw: Dict[str, torch.Tensor] = {}
        base, lora = self.base_views(), self.lora_views(self.policy)

        def draw(name, shape, scale, shift):
            n = 1
            for x in shape:
                n *= x
            t = torch.empty(n, dtype=torch.bfloat16, device=self.device)
            ops.init_uniform_(t, tensor_seed(name, seed), scale, shift)
            return t.view(shape)

        for name, shape, scale, shift in qwen_weight_specs(self.cfg):
            if name.startswith("transformer.visual."):
                self.set_visual_tensor(name, draw(name, shape, scale, shift))
            else:
                base[name].copy_(draw(name, shape, scale, shift))
        self.finalize_visual()
        for name, shape, scale, shift in qwen_lora_specs(self.cfg):
            lora[name].copy_(draw(name, shape, scale, shift))
        self.sync_master_from_params()
        if self.device.type == "cuda":
            torch.cuda.synchronize()

    def load_state_dict_tensors(self, tensors):
        """Iterable of (reference state-dict name, tensor): base LM, vision tower and (optionally) adapters."""
        self.wait_optimizer()
        self._pool_raw = {}
        base, lora = self.base_views(), self.lora_views(self.policy)
        for name, t in tensors:
            if name.startswith("transformer.visual."):
                self.set_visual_tensor(name, t)
                self.extra_state[name] = t.detach().to("cpu")
            elif name in base:
                base[name].copy_(t.to(self.device, torch.bfloat16).reshape(base[name].shape))
            elif name in lora:
                lora[name].copy_(t.to(self.device, torch.bfloat16).reshape(lora[name].shape))
            else:
                self.extra_state[name] = t.detach().to("cpu")
        self.finalize_visual()
        self.sync_master_from_params()

    # ------------------------------------------------------------------ vision tower + resampler (frozen, once per pair)
    def vision_features(self, pixels: torch.Tensor) -> torch.Tensor:
        cfg, v = self.cfg, self.vis
        Bv = pixels.shape[0]
        P, Q, w, d = cfg.n_patches, cfg.n_queries, cfg.v_width, cfg.hidden
        H, hp = cfg.v_heads, cfg.v_head_pad
        K, Kp = cfg.patch_k, cfg.patch_k_padded
        patches = self.buf("v.patches", (Bv * P, Kp))
        ops.clip_im2col(pixels, cfg.patch_size, patches)
        xb = self.buf("v.x", (Bv * P, w))
        for b in range(Bv):  # conv-as-GEMM; the epilogue adds the (interpolated) position table
            ops.gemm(patches[b * P:(b + 1) * P, :K], v["v.patch"][:, :K], out=xb[b * P:(b + 1) * P], residual=v["v.pos"])
        h = self.buf("v.h", (Bv * P, w))
        ops.layernorm_fwd(xb, v["v.pre.w"], v["v.pre.b"], cfg.v_eps, out=h)
        x = self.buf("v.x32", (Bv * P, w), torch.float32)
        ops.cast_bf16_to_f32(h.view(-1), x.view(-1))
        qkv = self.buf("v.qkv", (Bv * P, 3 * H * hp))
        att = self.buf("v.att", (Bv * P, H * hp))
        f = self.buf("v.f", (Bv * P, cfg.v_mlp))
        scale = cfg.v_head_dim ** -0.5
        for i in range(cfg.v_layers):
            ops.layernorm_fwd(x, v[f"v{i}.ln1.w"], v[f"v{i}.ln1.b"], cfg.v_eps, out=h)
            ops.gemm(h, v[f"v{i}.wqkv"], out=qkv, bias=v[f"v{i}.bqkv"])
            ops.attn_fwd_tc(qkv[:, :H * hp], qkv[:, H * hp:2 * H * hp], qkv[:, 2 * H * hp:], att, None, None, Bv, P, H, H, hp,
                            False, scale)
            ops.gemm(att, v[f"v{i}.wo"], out=x, bias=v[f"v{i}.bo"], residual=x)
            ops.layernorm_fwd(x, v[f"v{i}.ln2.w"], v[f"v{i}.ln2.b"], cfg.v_eps, out=h)
            ops.gemm(h, v[f"v{i}.w1"], out=f, bias=v[f"v{i}.b1"], act=ops.ACT_GELU_ERF)
            ops.gemm(f, v[f"v{i}.w2"], out=x, bias=v[f"v{i}.b2"], residual=x)
        ops.cast_f32_to_bf16(x.view(-1), xb.view(-1))
        # resampler: k = (ln_kv(kv_proj x) + pos) Wk^T + bk, v = ln_kv(kv_proj x) Wv^T + bv, q = constant table
        kvp = self.buf("r.kvp", (Bv * P, d))
        ops.gemm(xb, v["r.wkv"], out=kvp)
        kvn = self.buf("r.kvn", (Bv * P, d))
        ops.layernorm_fwd(kvp, v["r.lnkv.w"], v["r.lnkv.b"], cfg.v_eps, out=kvn)
        kk = self.buf("r.k", (Bv * P, d)); vv = self.buf("r.v", (Bv * P, d))
        for b in range(Bv):
            ops.gemm(kvn[b * P:(b + 1) * P], v["r.wk"], out=kk[b * P:(b + 1) * P], bias=v["r.bk"], residual=v["r.kpos"])
        ops.gemm(kvn, v["r.wv"], out=vv, bias=v["r.bv"])
        qpad = self.buf("r.qpad", (Bv * P, d))
        if getattr(self, "_qpad_for", None) != (Bv, qpad.data_ptr()):
            ops.zero_(qpad)
            ops.copy_rows(v["r.q"], 0, d, 0, qpad, P * d, d, Bv, Q, d)  # the same Q rows at the top of every image block
            self._qpad_for = (Bv, qpad.data_ptr())
        ro = self.buf("r.att", (Bv * P, d))
        ops.attn_fwd_tc(qpad, kk, vv, ro, None, None, Bv, P, cfg.r_heads, cfg.r_heads, 128, False, 1.0 / math.sqrt(128.0))
        o = self.buf("r.o", (Bv * Q, d))
        ops.copy_rows(ro, P * d, d, 0, o, Q * d, d, Bv, Q, d)
        o2 = self.buf("r.o2", (Bv * Q, d))
        ops.gemm(o, v["r.wo"], out=o2, bias=v["r.bo"])
        ops.layernorm_fwd(o2, v["r.post.w"], v["r.post.b"], cfg.v_eps, out=o)
        feats = self.buf("v.feats", (Bv * Q, d))
        ops.gemm(o, v["r.proj"], b_kmajor=False, out=feats)
        return feats

    # ------------------------------------------------------------------ decoder layer (base weights b, adapters l or None)
    def _layer_bufs(self, pre: str, sfx: str, m):
        b = super()._layer_bufs(pre, sfx, m)
        T, r = m.T, self.cfg.lora_r
        if pre == "a":
            b.update(ts_qkv=self.buf(f"a.ts_qkv{sfx}", (T, r)), ts_o=self.buf(f"a.ts_o{sfx}", (T, r)),
                     ts_gu=self.buf(f"a.ts_gu{sfx}", (T, 2 * r)))
        return b

    def _layer_fwd(self, w, i: int, x, b, m, xn, lora: Optional[Weights] = None):
        cfg, base = self.cfg, self.base
        d, T, ff, r = cfg.hidden, m.T, cfg.ff, cfg.lora_r
        H, dh = cfg.heads, cfg.head_dim
        h = self.buf("s.h", (T, d))
        qkv, att, xmid, gu = b["qkv"], b["att"], b["xmid"], b["gu"]
        ops.rmsnorm_fwd(x, base[f"L{i}.ln1"], cfg.rms_eps, out=h, rstd=b["rstd1"])
        def lora_t(xin, A, key, cols, scratch):
            """ts = bf16(s * xin A^T): kept in the saved set when the backward will need it, else in scratch"""
            ts = b[key] if key in b else self.buf(scratch, (T, cols))
            ops.gemm(xin, A, out=ts, alpha=cfg.lora_scale)
            return ts

        # every adapted linear is ONE launch: y = x W^T + b + ts B^T (second operand pair of the GEMM)
        if lora is None:
            ops.gemm(h, base[f"L{i}.wqkv"], out=qkv, bias=base[f"L{i}.bqkv"])
        else:
            ts = lora_t(h, lora[f"L{i}.qkv.A"], "ts_qkv", r, "l.ts")
            ops.gemm(h, base[f"L{i}.wqkv"], a2=ts, b2=lora[f"L{i}.qkv.B"], out=qkv, bias=base[f"L{i}.bqkv"])
        ops.rope_(qkv, m.pos, self.rope_cos, self.rope_sin, 2 * H, dh)
        attn_forward(m, qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], att, b["lse"], H, H, dh, 1.0 / math.sqrt(dh))
        if lora is None:
            ops.gemm(att, base[f"L{i}.wo"], out=xmid, residual=x)
        else:
            ts = lora_t(att, lora[f"L{i}.o.A"], "ts_o", r, "l.ts")
            ops.gemm(att, base[f"L{i}.wo"], a2=ts, b2=lora[f"L{i}.o.B"], out=xmid, residual=x)
        ops.rmsnorm_fwd(xmid, base[f"L{i}.ln2"], cfg.rms_eps, out=h, rstd=b["rstd2"])
        wgu = base[f"L{i}.wgu"]
        act = self.buf("s.act", (T, ff))
        if lora is None:   # reference pass: SwiGLU in the GEMM epilogue, gate|up never reaches HBM
            ops.gemm_swiglu(h, wgu, gu, act, write_gu=False)
        else:
            ts = lora_t(h, lora[f"L{i}.gu.A"], "ts_gu", 2 * r, "l.ts2")
            ops.gemm(h, wgu[:ff], a2=ts[:, :r], b2=lora[f"L{i}.w2.B"], out=gu[:, :ff])
            ops.gemm(h, wgu[ff:], a2=ts[:, r:], b2=lora[f"L{i}.w1.B"], out=gu[:, ff:])
        if xn is not None:
            if lora is not None:
                ops.swiglu_fwd(gu, act)
            ops.gemm(act, base[f"L{i}.wd"], out=xn, residual=xmid)

    # ------------------------------------------------------------------ forward of one pass
    def _forward(self, w, m, feats, tag: str, save: bool, ddpo_weight):
        cfg, base = self.cfg, self.base
        d, T = cfg.hidden, m.T
        lora = self.policy if tag == "policy" else None
        x = self.buf("x.0" if save else "s.x0", (T, d), torch.float32)
        ops.llava_merge_embed(m, base["embed"], feats, x)   # text rows: wte; placeholder rows: image features (:614-621)
        ckpt = save and self.tc.activation_checkpointing
        for i in range(cfg.layers):
            keep = save and not ckpt
            b = self._layer_bufs("a" if keep else "s", f".{i}" if keep else "", m)
            xn = self.buf(f"x.{i + 1}" if save else ("s.x1" if i % 2 == 0 else "s.x0"), (T, d), torch.float32)
            self._layer_fwd(None, i, x, b, m, xn, lora)
            x = xn
        return self._head_forward(x, base["norm"], base["lm_head"], m, feats, save, ddpo_weight)

    # ------------------------------------------------------------------ backward: adapter gradients only
    def _backward(self, grad_logps: torch.Tensor, accumulate: bool = False):
        """accumulate: add this micro-batch's gradients to the gradient arena (gradient_accumulation_steps > 1) instead
        of overwriting it -- every weight-gradient GEMM / reduction takes its `accumulate` epilogue."""
        acc = bool(accumulate)
        self.wait_optimizer()
        cfg, base, lora, g = self.cfg, self.base, self.policy, self.g
        sv = self._saved
        m = sv["m"]
        d, T, ff, r = cfg.hidden, m.T, cfg.ff, cfg.lora_r
        H, dh = cfg.heads, cfg.head_dim
        s = cfg.lora_scale
        dx = self._head_backward(grad_logps, base["norm"], base["lm_head"], self._dw_scratch, None)
        dxf = self._bufs["b.dxf"]
        dx2 = self.buf("b.dx1", (T, d))
        h = self.buf("s.h", (T, d))
        act = self.buf("s.act", (T, ff))
        dact = self.buf("b.dact", (T, ff))
        dnorm = dxf
        dqkv = self.buf("b.dqkv", (T, 3 * d))
        datt = self.buf("b.datt", (T, d))
        delta = self.buf("b.delta", (m.n_attn_seq, H, m.S), torch.float32)
        dt = self.buf("b.dt", (T, 2 * r))   # dt = bf16(s * dy B) of the two MLP adapters, side by side
        dr = self.buf("b.dr", (T, r))       # same for the r-wide adapters (c_attn, attn.c_proj)
        scale = 1.0 / math.sqrt(dh)

        for i in reversed(range(cfg.layers)):
            x_in = self._bufs[f"x.{i}"]
            if self.tc.activation_checkpointing:
                sb = self._layer_bufs("a", ".ckpt", m)   # one recompute set incl. the LoRA intermediates
                self._layer_fwd(None, i, x_in, sb, m, None, lora)
            else:
                sb = self._layer_bufs("a", f".{i}", m)
            xmid, gu, qkv, att = (sb[k] for k in ("xmid", "gu", "qkv", "att"))
            rstd1, rstd2, lse = (sb[k] for k in ("rstd1", "rstd2", "lse"))
            # ---- MLP (no LoRA on mlp.c_proj, base frozen: dgrad only)
            ops.rmsnorm_fwd(xmid, base[f"L{i}.ln2"], cfg.rms_eps, out=h)                      # recompute h2
            ops.gemm(dx, base[f"L{i}.wd"], b_kmajor=False, out=dact)                          # dact = dx Wd
            ops.swiglu_bwd(gu, dact, out=gu)                                                  # dgu (in place)
            tsg = sb["ts_gu"]
            ops.gemm(gu[:, :ff], tsg[:, :r], a_kmajor=False, b_kmajor=False, out=g[f"L{i}.w2.B"], accumulate=acc)   # dB2 = dgate^T ts2
            ops.gemm(gu[:, ff:], tsg[:, r:], a_kmajor=False, b_kmajor=False, out=g[f"L{i}.w1.B"], accumulate=acc)   # dB1 = dup^T ts1
            ops.gemm(gu[:, :ff], lora[f"L{i}.w2.B"], b_kmajor=False, out=dt[:, :r], alpha=s)   # dt2 = s dgate B2
            ops.gemm(gu[:, ff:], lora[f"L{i}.w1.B"], b_kmajor=False, out=dt[:, r:], alpha=s)   # dt1 = s dup B1
            ops.gemm(dt, h, a_kmajor=False, b_kmajor=False, out=g[f"L{i}.gu.A"], accumulate=acc)              # dA = dt^T h2  [2r, d]
            ops.gemm(gu, base[f"L{i}.wgu"], b_kmajor=False, a2=dt, b2=lora[f"L{i}.gu.A"], out=dnorm)   # dh2 = dgu Wgu + dt A
            ops.rmsnorm_bwd(dnorm, xmid, base[f"L{i}.ln2"], rstd2, self._dw_scratch, dres=dx, out=dx2)   # dxmid
            # ---- attention output projection (LoRA on attn.c_proj)
            ops.gemm(dx2, sb["ts_o"], a_kmajor=False, b_kmajor=False, out=g[f"L{i}.o.B"], accumulate=acc)     # dBo = dxmid^T ts_o
            ops.gemm(dx2, lora[f"L{i}.o.B"], b_kmajor=False, out=dr, alpha=s)                 # dt = s dxmid Bo
            ops.gemm(dr, att, a_kmajor=False, b_kmajor=False, out=g[f"L{i}.o.A"], accumulate=acc)            # dAo = dt^T att
            ops.gemm(dx2, base[f"L{i}.wo"], b_kmajor=False, a2=dr, b2=lora[f"L{i}.o.A"], out=datt)   # datt = dxmid Wo + dt Ao
            attn_backward(m, qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], att, datt, lse, delta, dqkv[:, :d], dqkv[:, d:2 * d],
                            dqkv[:, 2 * d:], H, H, dh, scale)
            ops.rope_(dqkv, m.pos, self.rope_cos, self.rope_sin, 2 * H, dh, inverse=True)
            # ---- fused qkv projection (LoRA on attn.c_attn; the bias is frozen)
            ops.rmsnorm_fwd(x_in, base[f"L{i}.ln1"], cfg.rms_eps, out=h)                      # recompute h1
            ops.gemm(dqkv, sb["ts_qkv"], a_kmajor=False, b_kmajor=False, out=g[f"L{i}.qkv.B"], accumulate=acc)
            ops.gemm(dqkv, lora[f"L{i}.qkv.B"], b_kmajor=False, out=dr, alpha=s)
            ops.gemm(dr, h, a_kmajor=False, b_kmajor=False, out=g[f"L{i}.qkv.A"], accumulate=acc)
            ops.gemm(dqkv, base[f"L{i}.wqkv"], b_kmajor=False, a2=dr, b2=lora[f"L{i}.qkv.A"], out=dnorm)   # dh1 = dqkv Wqkv + dt A
            ops.rmsnorm_bwd(dnorm, x_in, base[f"L{i}.ln1"], rstd1, self._dw_scratch, dres=dx2, out=dx)
            self._reduce_bucket(self.layout.offsets[f"L{i}.qkv.A"],
                                self.layout.offsets[f"L{i + 1}.qkv.A"] if i + 1 < cfg.layers else self.layout.size)

    # ------------------------------------------------------------------ inputs
    def prepare_inputs(self, input_ids, attention_mask, labels, pixel_values, ddpo_weight=None, image_sizes=None):
        dev = self.device
        n_seq = input_ids.shape[0]
        from . import host
        host.validate_token_batch(input_ids, labels, self.cfg.vocab, None, None, self.tc.label_pad_token_id)
        if pixel_values.shape[0] == n_seq:  # concatenated_inputs duplicated the images ([v, v], trainer.py:135-145)
            pixel_values = pixel_values[: n_seq // 2]
        ids = input_ids.to(dev, non_blocking=True).contiguous()
        am = attention_mask.to(dev, non_blocking=True).contiguous()
        lb = labels.to(dev, non_blocking=True).contiguous()
        px = pixel_values.to(dev, non_blocking=True).contiguous()
        if px.dtype not in (torch.float32, torch.bfloat16):
            px = px.float()
        wt = ddpo_weight.to(dev, non_blocking=True).reshape(-1).contiguous() if ddpo_weight is not None else None
        return ids, am, lb, px, wt

    def forward_logps(self, ids, am, lb, px, ddpo_weight=None, anyres=None, which: str = "policy", save: bool = False,
                      feats=None, m=None, seq_lens=None, prefix_rows=None):
        cfg = self.cfg
        self._anyres = None
        if m is None:
            m = ops.qwen_merge_index(ids, am, lb, cfg.n_queries, px.shape[0], 1, cfg.image_start_id, cfg.ignore_index)
            if self.tc.share_prefix:     # one copy of every pair's common prefix (engine.py; implies packed rows)
                if seq_lens is None or prefix_rows is None:
                    raise ValueError("share_prefix needs the host-side row plan: pass **engine.host_row_plan(ids, am)")
                ops.share_prefix_rows(m, seq_lens, prefix_rows)
                self._pad_rows, self._cur_rows = m.n_seq * m.S, m.T
            elif self.tc.pack_sequences:   # S == L here: the surviving rows are the attended tokens
                ops.pack_merge_rows(m, seq_lens if seq_lens is not None else m.seqlens.cpu().tolist())
                self._pad_rows, self._cur_rows = m.n_seq * m.S, m.T
        self.ensure_rope_len(m.S)
        if feats is None:
            feats = self.vision_features(px)
        if which == "policy":
            self.wait_optimizer()
        return self._forward(None, m, feats, which, save, ddpo_weight), m, feats

    def host_seq_lens(self, ids, am, image_sizes=None):
        """Packed steps: no token is expanded (S == L), a sequence keeps its attended tokens."""
        from . import host
        return host.merged_seq_lens(ids, am, -1, 0)

    def ddpo_weights(self, ids, am, lb, image_sizes=None) -> torch.Tensor:
        """DDPO row weights: no token is expanded (S == L), padding stays in the label sequence (as ignore labels)."""
        from . import host
        return host.ddpo_row_weights_native(ids, lb, -1, 0, self.tc.label_pad_token_id)

    def check_merge_status(self, m):
        st = int(m.status.item())
        if st == 2:
            raise ValueError("malformed image span: every sequence needs exactly one <img> ... </img> block holding "
                             f"{self.cfg.n_queries} placeholder tokens (modeling_qwen.py:524-528)")
        if st == 3:
            raise ValueError("attention_mask must be a right-padded prefix mask")
