"""The B200-native DPO / KTO-pair step for InternLM-XComposer2-VL with LoRA (BASELINE.json configs[4]).

Replaces, behind the same engine interface as engine.LlavaDPOEngine / engine_qwen.QwenVLDPOEngine:
  * `InternLMXC2ForRL.forward` + its merge                          (models/InternLMXC2/__init__.py:28-233)
  * `CLIPVisionTower` at 490 px + the Linear-GELU-Linear projector   (models/InternLMXC2/build_mlp.py:6-137)
  * `InternLM2Attention/MLP/DecoderLayer` with `PLoRA` linears       (modeling_internlm2.py:206-385,509-570; build_mlp.py:194-203)
  * peft LoRA r 64 on attention.wqkv/wo, feed_forward.w1/w2/w3        (scripts/dpo_internlmxc2vl7b.sh; reference = adapters off).

Built on the Qwen-VL engine's scheme (one frozen bf16 base for both passes, adapter-only backward); what is new here:
  * **partial LoRA (PLoRA)**: every linear adds `Plora_B(Plora_A(x)) * alpha/r` on the IMAGE rows only.  The image rows of
    all sequences are gathered once per linear input (`vlb200_gather_rows`), pushed through two dense GEMMs (9 800 x 256
    x 4096 at config 5) and added back with `vlb200_scatter_add_rows`; the backward mirrors it for the input gradient
    (the PLoRA weights are frozen: peft freezes every base parameter).
  * trainable LoRA also on the down projection (w2): its term accumulates into the fp32 residual stream like attention.wo.
  * InternLM2's fused wqkv is laid out per KV group as [q heads of the group | k | v]; it is re-laid out at load time to
    [all q | all k | all v] (rows of wqkv.weight, Plora_B and lora_B alike) so the attention kernels read plain column
    slices; `hf_state` undoes the permutation.
  * rotary positions are arange(S) for every row (modeling_internlm2.py:186-203 ignores position_ids).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from . import ops
from .config import XC2ModelConfig, tensor_seed, xc2_lora_specs, xc2_weight_specs
from .engine import Arena, LlavaDPOEngine, Weights, _vision_layout, attn_backward, attn_forward
from .engine_qwen import QwenVLDPOEngine


def _lora_layout(cfg: XC2ModelConfig) -> Arena:
    a = Arena()
    d, r, ff, hd = cfg.hidden, cfg.lora_r, cfg.ff, cfg.heads * cfg.head_dim
    for i in range(cfg.layers):
        a.add(f"L{i}.qkv.A", (r, d)); a.add(f"L{i}.qkv.B", (cfg.qkv_dim, r))
        a.add(f"L{i}.o.A", (r, hd)); a.add(f"L{i}.o.B", (d, r))
        a.add(f"L{i}.gu.A", (2 * r, d)); a.add(f"L{i}.w1.B", (ff, r)); a.add(f"L{i}.w3.B", (ff, r))
        a.add(f"L{i}.d.A", (r, ff)); a.add(f"L{i}.d.B", (d, r))
    return a


def _base_layout(cfg: XC2ModelConfig) -> Arena:
    a = Arena()
    d, pr, ff, hd = cfg.hidden, cfg.plora_r, cfg.ff, cfg.heads * cfg.head_dim
    a.add("proj.w1", (d, cfg.v_hidden)); a.add("proj.b1", (d,)); a.add("proj.w2", (d, d)); a.add("proj.b2", (d,))
    a.add("embed", (cfg.vocab, d))
    for i in range(cfg.layers):
        a.add(f"L{i}.ln1", (d,)); a.add(f"L{i}.wqkv", (cfg.qkv_dim, d)); a.add(f"L{i}.wo", (d, hd))
        a.add(f"L{i}.ln2", (d,)); a.add(f"L{i}.wgu", (2 * ff, d)); a.add(f"L{i}.wd", (d, ff))
        a.add(f"L{i}.p.qkv.A", (pr, d)); a.add(f"L{i}.p.qkv.B", (cfg.qkv_dim, pr))
        a.add(f"L{i}.p.o.A", (pr, hd)); a.add(f"L{i}.p.o.B", (d, pr))
        a.add(f"L{i}.p.gu.A", (2 * pr, d)); a.add(f"L{i}.p.w1.B", (ff, pr)); a.add(f"L{i}.p.w3.B", (ff, pr))
        a.add(f"L{i}.p.d.A", (pr, ff)); a.add(f"L{i}.p.d.B", (d, pr))
    a.add("norm", (d,)); a.add("lm_head", (cfg.vocab, d))
    return a


def qkv_permutation(cfg: XC2ModelConfig) -> torch.Tensor:
    """perm with new_rows = ref_rows[perm]: reference rows are (kv group, [q x n_rep | k | v], head_dim)."""
    H, KV, dh = cfg.heads, cfg.kv_heads, cfg.head_dim
    n_rep = H // KV
    idx = torch.arange((H + 2 * KV) * dh).view(KV, n_rep + 2, dh)
    return torch.cat([idx[:, :n_rep].reshape(-1), idx[:, n_rep].reshape(-1), idx[:, n_rep + 1].reshape(-1)])


class XC2DPOEngine(QwenVLDPOEngine):
    has_ref_copy = False
    needs_embed_grad = False

    def _make_layouts(self):
        return _lora_layout(self.cfg), _vision_layout(self.cfg)

    def _alloc_family(self):
        cfg = self.cfg
        self.blayout = _base_layout(cfg)
        self.bparams = torch.zeros(self.blayout.size, dtype=torch.bfloat16, device=self.device)
        self.base = Weights(self.blayout, self.bparams)
        self.extra_state: Dict[str, torch.Tensor] = {}
        self._dw_scratch = torch.zeros(cfg.hidden, dtype=torch.bfloat16, device=self.device)
        self._perm = qkv_permutation(cfg).to(self.device)
        self._inv_perm = torch.argsort(self._perm)
        # Combined second operands [lora_B | Plora_B] ([out, r + pr]) of the linears whose output is bf16 (wqkv, w1, w3): the
        # trainable LoRA term and the frozen partial-LoRA term then enter the base GEMM's accumulator TOGETHER (one rounding of
        # the output; the r1 path scatter-added the partial-LoRA term into the rounded result: 1.2e-3 on the 7B-shape log-probs).
        # Copies: the Plora columns are refreshed when the base weights change, the lora columns before every policy pass.
        ca = Arena()
        for i in range(cfg.layers):
            ca.add(f"L{i}.qkv.Bc", (cfg.qkv_dim, cfg.lora_r + cfg.plora_r))
            ca.add(f"L{i}.w1.Bc", (cfg.ff, cfg.lora_r + cfg.plora_r)); ca.add(f"L{i}.w3.Bc", (cfg.ff, cfg.lora_r + cfg.plora_r))
        self.clayout = ca
        self.cparams = torch.zeros(ca.size, dtype=torch.bfloat16, device=self.device)
        self.comb = Weights(ca, self.cparams)
        self._comb_base_stale = True

    # ------------------------------------------------------------------ names (storage views; wqkv-row tensors are permuted)
    def _lora_storage(self, w: Weights) -> Dict[str, torch.Tensor]:
        cfg, r = self.cfg, self.cfg.lora_r
        out: Dict[str, torch.Tensor] = {}
        for i in range(cfg.layers):
            p = f"model.layers.{i}."
            out[p + "attention.wqkv.lora_A"] = w[f"L{i}.qkv.A"]; out[p + "attention.wqkv.lora_B"] = w[f"L{i}.qkv.B"]
            out[p + "attention.wo.lora_A"] = w[f"L{i}.o.A"]; out[p + "attention.wo.lora_B"] = w[f"L{i}.o.B"]
            out[p + "feed_forward.w1.lora_A"] = w[f"L{i}.gu.A"][:r]; out[p + "feed_forward.w3.lora_A"] = w[f"L{i}.gu.A"][r:]
            out[p + "feed_forward.w1.lora_B"] = w[f"L{i}.w1.B"]; out[p + "feed_forward.w3.lora_B"] = w[f"L{i}.w3.B"]
            out[p + "feed_forward.w2.lora_A"] = w[f"L{i}.d.A"]; out[p + "feed_forward.w2.lora_B"] = w[f"L{i}.d.B"]
        return out

    def _base_storage(self) -> Dict[str, torch.Tensor]:
        cfg, b, pr = self.cfg, self.base, self.cfg.plora_r
        out = {"model.tok_embeddings.weight": b["embed"], "model.norm.weight": b["norm"], "output.weight": b["lm_head"],
               "vision_proj.0.weight": b["proj.w1"], "vision_proj.0.bias": b["proj.b1"], "vision_proj.2.weight": b["proj.w2"],
               "vision_proj.2.bias": b["proj.b2"]}
        for i in range(cfg.layers):
            p = f"model.layers.{i}."
            out[p + "attention_norm.weight"] = b[f"L{i}.ln1"]; out[p + "ffn_norm.weight"] = b[f"L{i}.ln2"]
            out[p + "attention.wqkv.weight"] = b[f"L{i}.wqkv"]; out[p + "attention.wo.weight"] = b[f"L{i}.wo"]
            out[p + "feed_forward.w1.weight"] = b[f"L{i}.wgu"][:cfg.ff]; out[p + "feed_forward.w3.weight"] = b[f"L{i}.wgu"][cfg.ff:]
            out[p + "feed_forward.w2.weight"] = b[f"L{i}.wd"]
            out[p + "attention.wqkv.Plora_A.weight"] = b[f"L{i}.p.qkv.A"]; out[p + "attention.wqkv.Plora_B.weight"] = b[f"L{i}.p.qkv.B"]
            out[p + "attention.wo.Plora_A.weight"] = b[f"L{i}.p.o.A"]; out[p + "attention.wo.Plora_B.weight"] = b[f"L{i}.p.o.B"]
            out[p + "feed_forward.w1.Plora_A.weight"] = b[f"L{i}.p.gu.A"][:pr]
            out[p + "feed_forward.w3.Plora_A.weight"] = b[f"L{i}.p.gu.A"][pr:]
            out[p + "feed_forward.w1.Plora_B.weight"] = b[f"L{i}.p.w1.B"]; out[p + "feed_forward.w3.Plora_B.weight"] = b[f"L{i}.p.w3.B"]
            out[p + "feed_forward.w2.Plora_A.weight"] = b[f"L{i}.p.d.A"]; out[p + "feed_forward.w2.Plora_B.weight"] = b[f"L{i}.p.d.B"]
        return out

    @staticmethod
    def _qkv_rows(name: str) -> bool:
        return name.endswith(("attention.wqkv.weight", "attention.wqkv.Plora_B.weight", "attention.wqkv.lora_B"))

    def lora_views(self, w: Weights) -> Dict[str, torch.Tensor]:
        """reference-layout tensors: views of the arena, except wqkv's lora_B which is un-permuted (a copy)."""
        return {k: (v.index_select(0, self._inv_perm) if self._qkv_rows(k) else v) for k, v in self._lora_storage(w).items()}

    def base_views(self) -> Dict[str, torch.Tensor]:
        return {k: (v.index_select(0, self._inv_perm) if self._qkv_rows(k) else v) for k, v in self._base_storage().items()}

    def _store(self, dst: Dict[str, torch.Tensor], name: str, t: torch.Tensor):
        self._comb_base_stale = True   # (a base tensor may have changed: the combined operands are rebuilt at the next pass)
        t = t.to(self.device, torch.bfloat16).reshape(dst[name].shape)
        dst[name].copy_(t.index_select(0, self._perm) if self._qkv_rows(name) else t)

    def init_synthetic(self, seed: int, ref_alpha: float = 0.0):
        """Seeded random-init base (incl. the frozen PLoRA adapters) + trainable adapters, bit-identical to
        oracle.xc2_restate.make_weights."""
        self.wait_optimizer()
        self._comb_base_stale = True
        cfg = self.cfg
        base, lora = self._base_storage(), self._lora_storage(self.policy)

        def draw(name, shape, scale, shift):
            n = 1
            for x in shape:
                n *= x
            t = torch.empty(n, dtype=torch.bfloat16, device=self.device)
            ops.init_uniform_(t, tensor_seed(name, seed), scale, shift)
            return t.view(shape)

        vnames = self._vision_names()
        for name, shape, scale, shift in xc2_weight_specs(cfg):
            if name.startswith("vit."):
                if name in vnames:
                    vnames[name].copy_(draw(name, shape, scale, shift).reshape(vnames[name].shape))
            else:
                self._store(base, name, draw(name, shape, scale, shift))
        for name, shape, scale, shift in xc2_lora_specs(cfg):
            self._store(lora, name, draw(name, shape, scale, shift))
        self.sync_master_from_params()
        if self.device.type == "cuda":
            torch.cuda.synchronize()

    def _vision_names(self) -> Dict[str, torch.Tensor]:
        """`vit.vision_tower.vision_model.*` -> views of the CLIP arena (fused q|k|v rows, padded patch kernel)."""
        cfg, v = self.cfg, self.vis
        out: Dict[str, torch.Tensor] = {}
        vp, dv = "vit.vision_tower.vision_model.", cfg.v_hidden
        out[vp + "embeddings.class_embedding"] = v["v.cls"]
        out[vp + "embeddings.patch_embedding.weight"] = v["v.patch"][:, :cfg.patch_k]
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

    def load_state_dict_tensors(self, tensors):
        self.wait_optimizer()
        base, lora, vis = self._base_storage(), self._lora_storage(self.policy), self._vision_names()
        for name, t in tensors:
            if name in vis:
                vis[name].copy_(t.to(self.device, torch.bfloat16).reshape(vis[name].shape))
            elif name in base:
                self._store(base, name, t)
            elif name in lora:
                self._store(lora, name, t)
            else:
                self.extra_state[name] = t.detach().to("cpu")
        self.sync_master_from_params()

    # ------------------------------------------------------------------ frozen tower (CLIP path of the LLaVA engine)
    def vision_features(self, pixels: torch.Tensor) -> torch.Tensor:
        return LlavaDPOEngine.vision_features(self, pixels)

    # ------------------------------------------------------------------ combined [lora_B | Plora_B] operands
    def _refresh_comb(self, lora: Optional[Weights]):
        """Copy the frozen Plora_B columns (when the base changed) and the current lora_B columns (policy pass) into the
        combined second operands."""
        cfg = self.cfg
        r, pr = cfg.lora_r, cfg.plora_r
        k2 = r + pr
        names = (("qkv", "qkv.B", "p.qkv.B"), ("w1", "w1.B", "p.w1.B"), ("w3", "w3.B", "p.w3.B"))
        for i in range(cfg.layers):
            for cn, ln, pn in names:
                dst = self.comb[f"L{i}.{cn}.Bc"]
                if self._comb_base_stale:
                    src = self.base[f"L{i}.{pn}"]
                    ops.copy_rows(src, 0, pr, 0, dst[:, r:], 0, k2, 1, src.shape[0], pr)
                if lora is not None:
                    src = lora[f"L{i}.{ln}"]
                    ops.copy_rows(src, 0, r, 0, dst, 0, k2, 1, src.shape[0], r)
        self._comb_base_stale = False

    def _adapter_operand(self, xin: torch.Tensor, A_p: torch.Tensor, ts: Optional[torch.Tensor], name: str, n_out: int):
        """-> list of n_out activation operands [T, r + pr] = [ts_k | scale_p * (x A_p_k^T) on the image rows, 0 elsewhere]
        (ts: the trainable LoRA activations [T, n_out * r] of the policy pass, or None)."""
        cfg = self.cfg
        r, pr, T = cfg.lora_r, cfg.plora_r, xin.shape[0]
        rows = self._img_rows
        n = rows.numel()
        xc = self.buf(f"p.xc.{xin.shape[1]}", (n, xin.shape[1]))
        ops.gather_rows(xin, rows, xc)
        tp = self.buf(f"p.t.{A_p.shape[0]}", (n, A_p.shape[0]))
        ops.gemm(xc, A_p, out=tp, alpha=cfg.plora_scale)
        outs = []
        for k in range(n_out):
            ta = self.buf(f"l.ts_all.{name}.{k}", (T, r + pr))
            ops.zero_(ta)                                   # text rows of the partial-LoRA columns (and the lora columns when off)
            if ts is not None:
                ops.copy_rows(ts[:, k * r:(k + 1) * r], 0, ts.stride(0), 0, ta, 0, r + pr, 1, T, r)
            ops.scatter_rows(tp[:, k * pr:(k + 1) * pr], rows, ta[:, r:])
            outs.append(ta)
        return outs

    # ------------------------------------------------------------------ one linear: base + trainable LoRA + frozen PLoRA
    def _plora_fwd(self, xin: torch.Tensor, A: torch.Tensor, outs, m, tag: str):
        """outs: list of (B [out, pr], dst [T, out] bf16|fp32, column range of the gathered t)."""
        cfg = self.cfg
        rows = self._img_rows
        n, pr = rows.numel(), A.shape[0]
        xc = self.buf(f"p.xc.{xin.shape[1]}", (n, xin.shape[1]))
        ops.gather_rows(xin, rows, xc)
        tp = self.buf(f"p.t.{pr}", (n, pr))
        ops.gemm(xc, A, out=tp)
        for B, dst, cols in outs:
            up = self.buf(f"p.up.{B.shape[0]}", (n, B.shape[0]))
            ops.gemm(tp[:, cols], B, out=up)
            ops.scatter_add_rows(up, rows, dst, cfg.plora_scale)

    def _plora_bwd(self, dys, A: torch.Tensor, dxin: torch.Tensor):
        """dxin[img rows] += scale * ([dy_k[img] B_k]_k) A ; dys: list of (dy [T, out], B [out, pr], column range)."""
        cfg = self.cfg
        rows = self._img_rows
        n, pr = rows.numel(), A.shape[0]
        dtp = self.buf(f"p.dt.{pr}", (n, pr))
        for dy, B, cols in dys:
            dyc = self.buf(f"p.dyc.{dy.shape[1]}", (n, dy.shape[1]))
            ops.gather_rows(dy, rows, dyc)
            ops.gemm(dyc, B, b_kmajor=False, out=dtp[:, cols])
        dxc = self.buf(f"p.dxc.{A.shape[1]}", (n, A.shape[1]))
        ops.gemm(dtp, A, b_kmajor=False, out=dxc)
        ops.scatter_add_rows(dxc, rows, dxin, cfg.plora_scale)

    def _layer_bufs(self, pre: str, sfx: str, m):
        b = LlavaDPOEngine._layer_bufs(self, pre, sfx, m)
        T, r = m.T, self.cfg.lora_r
        if pre == "a":
            b.update(ts_qkv=self.buf(f"a.ts_qkv{sfx}", (T, r)), ts_o=self.buf(f"a.ts_o{sfx}", (T, r)),
                     ts_gu=self.buf(f"a.ts_gu{sfx}", (T, 2 * r)), ts_d=self.buf(f"a.ts_d{sfx}", (T, r)))
        return b

    def _layer_fwd(self, w, i: int, x, b, m, xn, lora: Optional[Weights] = None):
        cfg, base = self.cfg, self.base
        d, T, ff, r, pr = cfg.hidden, m.T, cfg.ff, cfg.lora_r, cfg.plora_r
        H, KV, dh = cfg.heads, cfg.kv_heads, cfg.head_dim
        hd, kvd = H * dh, KV * dh
        h = self.buf("s.h", (T, d))
        qkv, att, xmid, gu = b["qkv"], b["att"], b["xmid"], b["gu"]

        def lora_t(xin, A, key, cols, scratch):
            ts = b[key] if key in b else self.buf(scratch, (T, cols))
            ops.gemm(xin, A, out=ts, alpha=cfg.lora_scale)
            return ts

        ops.rmsnorm_fwd(x, base[f"L{i}.ln1"], cfg.rms_eps, out=h, rstd=b["rstd1"])
        # y = x W^T + [ts | tp] [lora_B | Plora_B]^T in ONE launch (one rounding of the bf16 output); reference pass: the
        # partial-LoRA columns only
        ts = lora_t(h, lora[f"L{i}.qkv.A"], "ts_qkv", r, "l.ts") if lora is not None else None
        (ta,) = self._adapter_operand(h, base[f"L{i}.p.qkv.A"], ts, "qkv", 1)
        bc = self.comb[f"L{i}.qkv.Bc"]
        if lora is None:
            ops.gemm(h, base[f"L{i}.wqkv"], a2=ta[:, r:], b2=bc[:, r:], out=qkv)
        else:
            ops.gemm(h, base[f"L{i}.wqkv"], a2=ta, b2=bc, out=qkv)
        ops.rope_(qkv, m.pos, self.rope_cos, self.rope_sin, H + KV, dh)
        attn_forward(m, qkv[:, :hd], qkv[:, hd:hd + kvd], qkv[:, hd + kvd:], att, b["lse"], H, KV, dh, 1.0 / math.sqrt(dh))
        if lora is None:
            ops.gemm(att, base[f"L{i}.wo"], out=xmid, residual=x)
        else:
            ts = lora_t(att, lora[f"L{i}.o.A"], "ts_o", r, "l.ts")
            ops.gemm(att, base[f"L{i}.wo"], a2=ts, b2=lora[f"L{i}.o.B"], out=xmid, residual=x)
        self._plora_fwd(att, base[f"L{i}.p.o.A"], [(base[f"L{i}.p.o.B"], xmid, slice(0, pr))], m, "o")
        ops.rmsnorm_fwd(xmid, base[f"L{i}.ln2"], cfg.rms_eps, out=h, rstd=b["rstd2"])
        ts = lora_t(h, lora[f"L{i}.gu.A"], "ts_gu", 2 * r, "l.ts2") if lora is not None else None
        ta1, ta3 = self._adapter_operand(h, base[f"L{i}.p.gu.A"], ts, "gu", 2)
        wgu = base[f"L{i}.wgu"]
        c0 = r if lora is None else 0
        ops.gemm(h, wgu[:ff], a2=ta1[:, c0:], b2=self.comb[f"L{i}.w1.Bc"][:, c0:], out=gu[:, :ff])
        ops.gemm(h, wgu[ff:], a2=ta3[:, c0:], b2=self.comb[f"L{i}.w3.Bc"][:, c0:], out=gu[:, ff:])
        if xn is not None or (lora is not None and "ts_d" in b):
            act = self.buf("s.act", (T, ff))
            ops.swiglu_fwd(gu, act)
            if lora is not None:
                ts = lora_t(act, lora[f"L{i}.d.A"], "ts_d", r, "l.ts")
            if xn is not None:
                if lora is None:
                    ops.gemm(act, base[f"L{i}.wd"], out=xn, residual=xmid)
                else:
                    ops.gemm(act, base[f"L{i}.wd"], a2=ts, b2=lora[f"L{i}.d.B"], out=xn, residual=xmid)
                self._plora_fwd(act, base[f"L{i}.p.d.A"], [(base[f"L{i}.p.d.B"], xn, slice(0, pr))], m, "d")

    # ------------------------------------------------------------------ forward of one pass
    def _forward(self, w, m, feats, tag: str, save: bool, ddpo_weight):
        cfg, base = self.cfg, self.base
        d, T = cfg.hidden, m.T
        lora = self.policy if tag == "policy" else None
        self._refresh_comb(lora)
        nimg = feats.shape[0]
        ph = self.buf("p.h_ref", (nimg, d))                               # frozen projector (build_mlp.py:14-28)
        ops.gemm(feats, base["proj.w1"], out=ph, bias=base["proj.b1"], act=ops.ACT_GELU_ERF)
        img = self.buf("p.img", (nimg, d))
        ops.gemm(ph, base["proj.w2"], out=img, bias=base["proj.b2"])
        x = self.buf("x.0" if save else "s.x0", (T, d), torch.float32)
        ops.llava_merge_embed(m, base["embed"], img, x)
        ckpt = save and self.tc.activation_checkpointing
        for i in range(cfg.layers):
            keep = save and not ckpt
            b = self._layer_bufs("a" if keep else "s", f".{i}" if keep else "", m)
            xn = self.buf(f"x.{i + 1}" if save else ("s.x1" if i % 2 == 0 else "s.x0"), (T, d), torch.float32)
            self._layer_fwd(None, i, x, b, m, xn, lora)
            x = xn
        return self._head_forward(x, base["norm"], base["lm_head"], m, feats, save, ddpo_weight)

    # ------------------------------------------------------------------ backward: trainable-adapter gradients only
    def _backward(self, grad_logps: torch.Tensor, accumulate: bool = False):
        """accumulate: add this micro-batch's gradients to the gradient arena (gradient_accumulation_steps > 1) instead
        of overwriting it -- every weight-gradient GEMM / reduction takes its `accumulate` epilogue."""
        acc = bool(accumulate)
        self.wait_optimizer()
        cfg, base, lora, g = self.cfg, self.base, self.policy, self.g
        sv = self._saved
        m = sv["m"]
        d, T, ff, r, pr = cfg.hidden, m.T, cfg.ff, cfg.lora_r, cfg.plora_r
        H, KV, dh = cfg.heads, cfg.kv_heads, cfg.head_dim
        hd, kvd = H * dh, KV * dh
        s = cfg.lora_scale
        dx = self._head_backward(grad_logps, base["norm"], base["lm_head"], self._dw_scratch, None)
        dxf = self._bufs["b.dxf"]
        dx2 = self.buf("b.dx1", (T, d))
        h = self.buf("s.h", (T, d))
        act = self.buf("s.act", (T, ff))
        dact = self.buf("b.dact", (T, ff))
        dnorm = dxf
        dqkv = self.buf("b.dqkv", (T, cfg.qkv_dim))
        datt = self.buf("b.datt", (T, hd))
        delta = self.buf("b.delta", (m.n_attn_seq, H, m.S), torch.float32)
        dt = self.buf("b.dt", (T, 2 * r))   # dt = bf16(s * dy B): gate | up side by side; dr: the r-wide adapters
        dr = self.buf("b.dr", (T, r))
        scale = 1.0 / math.sqrt(dh)
        for i in reversed(range(cfg.layers)):
            x_in = self._bufs[f"x.{i}"]
            if self.tc.activation_checkpointing:
                sb = self._layer_bufs("a", ".ckpt", m)
                self._layer_fwd(None, i, x_in, sb, m, None, lora)   # recomputes up to gu + act + ts_d
            else:
                sb = self._layer_bufs("a", f".{i}", m)
                ops.swiglu_fwd(sb["gu"], act)                        # recompute act
            xmid, gu, qkv, att = (sb[k] for k in ("xmid", "gu", "qkv", "att"))
            rstd1, rstd2, lse = (sb[k] for k in ("rstd1", "rstd2", "lse"))
            # ---- down projection (LoRA + PLoRA on feed_forward.w2)
            ops.gemm(dx, sb["ts_d"], a_kmajor=False, b_kmajor=False, out=g[f"L{i}.d.B"], accumulate=acc)      # dBd = dx^T ts_d
            ops.gemm(dx, lora[f"L{i}.d.B"], b_kmajor=False, out=dr, alpha=s)
            ops.gemm(dr, act, a_kmajor=False, b_kmajor=False, out=g[f"L{i}.d.A"], accumulate=acc)             # dAd = dt^T act
            ops.gemm(dx, base[f"L{i}.wd"], b_kmajor=False, a2=dr, b2=lora[f"L{i}.d.A"], out=dact)   # dact = dx Wd + dt Ad
            self._plora_bwd([(dx, base[f"L{i}.p.d.B"], slice(0, pr))], base[f"L{i}.p.d.A"], dact)
            # ---- gate | up
            ops.rmsnorm_fwd(xmid, base[f"L{i}.ln2"], cfg.rms_eps, out=h)                      # recompute h2
            ops.swiglu_bwd(gu, dact, out=gu)                                                  # dgu (in place)
            tsg = sb["ts_gu"]
            ops.gemm(gu[:, :ff], tsg[:, :r], a_kmajor=False, b_kmajor=False, out=g[f"L{i}.w1.B"], accumulate=acc)
            ops.gemm(gu[:, ff:], tsg[:, r:], a_kmajor=False, b_kmajor=False, out=g[f"L{i}.w3.B"], accumulate=acc)
            ops.gemm(gu[:, :ff], lora[f"L{i}.w1.B"], b_kmajor=False, out=dt[:, :r], alpha=s)
            ops.gemm(gu[:, ff:], lora[f"L{i}.w3.B"], b_kmajor=False, out=dt[:, r:], alpha=s)
            ops.gemm(dt, h, a_kmajor=False, b_kmajor=False, out=g[f"L{i}.gu.A"], accumulate=acc)
            ops.gemm(gu, base[f"L{i}.wgu"], b_kmajor=False, a2=dt, b2=lora[f"L{i}.gu.A"], out=dnorm)   # dh2 = dgu Wgu + dt A
            self._plora_bwd([(gu[:, :ff], base[f"L{i}.p.w1.B"], slice(0, pr)), (gu[:, ff:], base[f"L{i}.p.w3.B"], slice(pr, 2 * pr))],
                            base[f"L{i}.p.gu.A"], dnorm)
            ops.rmsnorm_bwd(dnorm, xmid, base[f"L{i}.ln2"], rstd2, self._dw_scratch, dres=dx, out=dx2)   # dxmid
            # ---- attention output projection
            ops.gemm(dx2, sb["ts_o"], a_kmajor=False, b_kmajor=False, out=g[f"L{i}.o.B"], accumulate=acc)
            ops.gemm(dx2, lora[f"L{i}.o.B"], b_kmajor=False, out=dr, alpha=s)
            ops.gemm(dr, att, a_kmajor=False, b_kmajor=False, out=g[f"L{i}.o.A"], accumulate=acc)
            ops.gemm(dx2, base[f"L{i}.wo"], b_kmajor=False, a2=dr, b2=lora[f"L{i}.o.A"], out=datt)
            self._plora_bwd([(dx2, base[f"L{i}.p.o.B"], slice(0, pr))], base[f"L{i}.p.o.A"], datt)
            attn_backward(m, qkv[:, :hd], qkv[:, hd:hd + kvd], qkv[:, hd + kvd:], att, datt, lse, delta, dqkv[:, :hd],
                            dqkv[:, hd:hd + kvd], dqkv[:, hd + kvd:], H, KV, dh, scale)
            ops.rope_(dqkv, m.pos, self.rope_cos, self.rope_sin, H + KV, dh, inverse=True)
            # ---- fused qkv projection
            ops.rmsnorm_fwd(x_in, base[f"L{i}.ln1"], cfg.rms_eps, out=h)                      # recompute h1
            ops.gemm(dqkv, sb["ts_qkv"], a_kmajor=False, b_kmajor=False, out=g[f"L{i}.qkv.B"], accumulate=acc)
            ops.gemm(dqkv, lora[f"L{i}.qkv.B"], b_kmajor=False, out=dr, alpha=s)
            ops.gemm(dr, h, a_kmajor=False, b_kmajor=False, out=g[f"L{i}.qkv.A"], accumulate=acc)
            ops.gemm(dqkv, base[f"L{i}.wqkv"], b_kmajor=False, a2=dr, b2=lora[f"L{i}.qkv.A"], out=dnorm)
            self._plora_bwd([(dqkv, base[f"L{i}.p.qkv.B"], slice(0, pr))], base[f"L{i}.p.qkv.A"], dnorm)
            ops.rmsnorm_bwd(dnorm, x_in, base[f"L{i}.ln1"], rstd1, self._dw_scratch, dres=dx2, out=dx)
            self._reduce_bucket(self.layout.offsets[f"L{i}.qkv.A"],
                                self.layout.offsets[f"L{i + 1}.qkv.A"] if i + 1 < cfg.layers else self.layout.size)

    # ------------------------------------------------------------------ inputs
    def prepare_inputs(self, input_ids, attention_mask, labels, pixel_values, ddpo_weight=None, image_sizes=None):
        from . import host
        host.validate_token_batch(input_ids, None, self.cfg.vocab, self.cfg.image_token_index, 1)   # one image per sequence
        return QwenVLDPOEngine.prepare_inputs(self, input_ids, attention_mask, labels, pixel_values, ddpo_weight)

    def forward_logps(self, ids, am, lb, px, ddpo_weight=None, anyres=None, which: str = "policy", save: bool = False,
                      feats=None, m=None, seq_lens=None, prefix_rows=None):
        cfg = self.cfg
        self._anyres = None
        if m is None:
            m = ops.llava_merge_index(ids, am, lb, cfg.n_patches, px.shape[0], 1, cfg.image_token_index, cfg.pad_token_id,
                                      cfg.ignore_index)
            # rotary positions are arange(S) for every row, padded or not (modeling_internlm2.py:186-203)
            m.pos = torch.arange(m.S, dtype=torch.int32, device=ids.device).repeat(m.n_seq)
            if self.tc.share_prefix:     # one copy of every pair's common prefix (engine.py; the rotary positions are packed along)
                if seq_lens is None or prefix_rows is None:
                    raise ValueError("share_prefix needs the host-side row plan: pass **engine.host_row_plan(ids, am)")
                ops.share_prefix_rows(m, seq_lens, prefix_rows)
                self._pad_rows, self._cur_rows = m.n_seq * m.S, m.T
            elif self.tc.pack_sequences:   # drop the padding rows (the rotary positions above are packed along)
                ops.pack_merge_rows(m, seq_lens if seq_lens is not None else m.seqlens.cpu().tolist())
                self._pad_rows, self._cur_rows = m.n_seq * m.S, m.T
        # flat merged rows of every image position (the PLoRA row mask im_mask, __init__.py:87-104); packed: already absolute
        if getattr(self, "_img_rows_of", None) is not m:
            self._img_rows = m.img_pos if m.packed else (
                m.img_pos.view(m.n_seq, -1) +
                (torch.arange(m.n_seq, dtype=torch.int32, device=m.img_pos.device) * m.S)[:, None]).reshape(-1).contiguous()
            if m.shared and prefix_rows is not None and all(int(p) > 0 for p in prefix_rows):
                # every pair shares its image rows: the rejected half of the list is all -1 (skipped rows) -- drop it
                self._img_rows = m.img_pos[: m.img_pos.numel() // 2]
            self._img_rows_of = m
        self.ensure_rope_len(m.S)
        if feats is None:
            feats = self.vision_features(px)
        if which == "policy":
            self.wait_optimizer()
        return self._forward(None, m, feats, which, save, ddpo_weight), m, feats

    def host_seq_lens(self, ids, am, image_sizes=None):
        return LlavaDPOEngine.host_seq_lens(self, ids, am, image_sizes)   # one <image> token -> n_patches rows, as LLaVA-1.5

    def ddpo_weights(self, ids, am, lb, image_sizes=None) -> torch.Tensor:
        from . import host
        return host.ddpo_row_weights_native(ids, lb, self.cfg.image_token_index, self.cfg.n_patches, self.tc.label_pad_token_id)

    def check_merge_status(self, m):
        return LlavaDPOEngine.check_merge_status(self, m)
