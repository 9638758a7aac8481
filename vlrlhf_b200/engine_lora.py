"""The B200-native DPO step for LLaVA-1.5 / LLaVA-Next trained with peft LoRA on the decoder linears -- the
configuration every reference launch script runs (scripts/dpo_llava.sh:24-30, dpo_llavanext.sh:24-30, kto_llava.sh,
ddpo_llava.sh: `--use_lora True --lora_r 128 --lora_alpha 256 --lora_target_modules auto`; utils/auto_load.py:559-578;
`default_lora_target`, models/Llava/__init__.py:273-286, models/LlavaNext/__init__.py:347-360).

Same interface as engine.LlavaDPOEngine (prepare_inputs / forward_logps / step / train_step), same vision tower, projector,
anyres packing, merge, head and loss kernels.  What changes with LoRA:
  * ONE frozen bf16 copy of projector + LLM serves both passes: policy = base + adapters, reference = base with the
    adapters disabled (TRL's `null_ref_context()` for a peft model without ref_model).  HBM holds 13.5 GB of weights
    instead of 27 GB and no fp32 master/moments for them; the trainable / gradient / optimizer arenas hold only the
    adapters (320 M parameters at r = 128 on the 7B decoder) and so does the data-parallel gradient reduction.
  * backward runs no base weight-gradient GEMM (a third of the full fine-tuning step's FLOPs) and stops at decoder layer
    0: embeddings, projector, image_newline and the tower are frozen by peft.
  * LoRA linear (peft lora.Linear.forward, dropout off under TRL's disable_dropout): t = x A^T (fp32 accumulator) ->
    ts = bf16(s t) -> y = x W^T + ts B^T in ONE launch (the adapter term is a second operand pair contracted into the
    base GEMM's accumulator, `vlb200_gemm_bf16_ex`: no [T, out]-sized intermediate in HBM, one rounding); q/k/v_proj and
    gate/up_proj share their input, so their A matrices are stacked ([3r, d], [2r, d]: one GEMM each).  Backward:
    dB = dy^T ts, dt = bf16(s dy B), dA = dt^T x, dx = dy W + dt A (again one launch with two operand pairs).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from . import ops
from .config import LlavaLoRAModelConfig, llava_lora_specs, tensor_seed, weight_specs
from .engine import Arena, LlavaDPOEngine, Weights, _trainable_layout, _vision_layout, attn_backward, attn_forward, hf_views


def _lora_layout(cfg: LlavaLoRAModelConfig) -> Arena:
    a = Arena()
    d, r, ff = cfg.hidden, cfg.lora_r, cfg.ff
    hd, kvd = cfg.heads * cfg.head_dim, cfg.kv_heads * cfg.head_dim
    for i in range(cfg.layers):
        a.add(f"L{i}.qkv.A", (3 * r, d)); a.add(f"L{i}.q.B", (hd, r)); a.add(f"L{i}.k.B", (kvd, r)); a.add(f"L{i}.v.B", (kvd, r))
        a.add(f"L{i}.o.A", (r, hd)); a.add(f"L{i}.o.B", (d, r))
        a.add(f"L{i}.gu.A", (2 * r, d)); a.add(f"L{i}.g.B", (ff, r)); a.add(f"L{i}.u.B", (ff, r))
        a.add(f"L{i}.d.A", (r, ff)); a.add(f"L{i}.d.B", (d, r))
    return a


class LlavaLoRADPOEngine(LlavaDPOEngine):
    has_ref_copy = False
    needs_embed_grad = False

    def _make_layouts(self):
        if getattr(self.cfg, "lora_r", 0) <= 0:
            raise ValueError("LlavaLoRADPOEngine needs a LlavaLoRAModelConfig (config.with_lora(cfg, r, alpha))")
        return _lora_layout(self.cfg), _vision_layout(self.cfg)

    def _alloc_family(self):
        cfg = self.cfg
        self.blayout = _trainable_layout(cfg)   # the full-FT engine's trainable arena is this engine's frozen base
        self.bparams = torch.zeros(self.blayout.size, dtype=torch.bfloat16, device=self.device)
        self.base = Weights(self.blayout, self.bparams)
        self.extra_state: Dict[str, torch.Tensor] = {}
        self._dw_scratch = torch.zeros(cfg.hidden, dtype=torch.bfloat16, device=self.device)  # frozen norm weights' "gradient"

    # ------------------------------------------------------------------ names
    def lora_views(self, w: Weights) -> Dict[str, torch.Tensor]:
        """`language_model.model.layers.{i}.<linear>.lora_A|lora_B` -> views of the adapter arena."""
        cfg, r = self.cfg, self.cfg.lora_r
        out: Dict[str, torch.Tensor] = {}
        for i in range(cfg.layers):
            p = f"language_model.model.layers.{i}."
            qa, ga = w[f"L{i}.qkv.A"], w[f"L{i}.gu.A"]
            for j, n in enumerate(("q", "k", "v")):
                out[p + f"self_attn.{n}_proj.lora_A"] = qa[j * r:(j + 1) * r]
                out[p + f"self_attn.{n}_proj.lora_B"] = w[f"L{i}.{n}.B"]
            out[p + "self_attn.o_proj.lora_A"] = w[f"L{i}.o.A"]; out[p + "self_attn.o_proj.lora_B"] = w[f"L{i}.o.B"]
            out[p + "mlp.gate_proj.lora_A"] = ga[:r]; out[p + "mlp.gate_proj.lora_B"] = w[f"L{i}.g.B"]
            out[p + "mlp.up_proj.lora_A"] = ga[r:]; out[p + "mlp.up_proj.lora_B"] = w[f"L{i}.u.B"]
            out[p + "mlp.down_proj.lora_A"] = w[f"L{i}.d.A"]; out[p + "mlp.down_proj.lora_B"] = w[f"L{i}.d.B"]
        return out

    def base_views(self) -> Dict[str, torch.Tensor]:
        return hf_views(self.cfg, self.base.t, self.vis.t)

    def hf_state(self, which: str = "policy") -> Dict[str, torch.Tensor]:
        """policy: base + adapters; ref: base (adapters disabled); grad: adapter gradients."""
        self.wait_optimizer()
        if which == "grad":
            return self.lora_views(self.g)
        out = dict(self.base_views())
        if which == "policy":
            out.update(self.lora_views(self.policy))
        return out

    # ------------------------------------------------------------------ weights
    def init_synthetic(self, seed: int, ref_alpha: float = 0.0):
        """Seeded random-init base + adapters (bit-identical to oracle.lora_restate.make_weights)."""
        self.wait_optimizer()
        base, lora = self.base_views(), self.lora_views(self.policy)

        def draw(name, shape, scale, shift):
            n = 1
            for x in shape:
                n *= x
            t = torch.empty(n, dtype=torch.bfloat16, device=self.device)
            ops.init_uniform_(t, tensor_seed(name, seed), scale, shift)
            return t

        for name, shape, scale, shift in weight_specs(self.cfg):
            if name in base:  # vision layers above vision_feature_layer are never used by the path
                base[name].copy_(draw(name, shape, scale, shift).view(base[name].shape))
        for name, shape, scale, shift in llava_lora_specs(self.cfg):
            lora[name].copy_(draw(name, shape, scale, shift).view(shape))
        self.sync_master_from_params()
        if self.device.type == "cuda":
            torch.cuda.synchronize()

    def reset_adapters(self, seed: int = 0):
        """peft's LoRA init: A ~ kaiming_uniform(a=sqrt(5)) = U(-1/sqrt(in), 1/sqrt(in)), B = 0 (policy == reference)."""
        self.wait_optimizer()
        gen = torch.Generator().manual_seed(seed)
        for name, t in self.lora_views(self.policy).items():
            if name.endswith("lora_A"):
                bound = 1.0 / math.sqrt(t.shape[1])
                t.copy_(((torch.rand(t.shape, generator=gen) * 2 - 1) * bound).to(t.device, torch.bfloat16))
            else:
                t.zero_()
        self.sync_master_from_params()

    def load_hf_state_dict(self, sd: Dict[str, torch.Tensor], which: str = "policy"):
        dst = self.hf_state("policy")
        for k, v in sd.items():
            if k in dst:
                dst[k].copy_(v.to(device=self.device, dtype=torch.bfloat16).view(dst[k].shape))
        self.sync_master_from_params()

    # ------------------------------------------------------------------ decoder layer (base weights, adapters `lora` or None)
    def _layer_bufs(self, pre: str, sfx: str, m):
        b = super()._layer_bufs(pre, sfx, m)
        T, r = m.T, self.cfg.lora_r
        if pre == "a":
            b.update(ts_qkv=self.buf(f"a.ts_qkv{sfx}", (T, 3 * r)), ts_o=self.buf(f"a.ts_o{sfx}", (T, r)),
                     ts_gu=self.buf(f"a.ts_gu{sfx}", (T, 2 * r)), ts_d=self.buf(f"a.ts_d{sfx}", (T, r)))
        return b

    def _layer_fwd(self, w, i: int, x, b, m, xn, lora: Optional[Weights] = None):
        cfg, base = self.cfg, self.base
        d, T, ff, r = cfg.hidden, m.T, cfg.ff, cfg.lora_r
        H, KV, dh = cfg.heads, cfg.kv_heads, cfg.head_dim
        hd, kvd = H * dh, KV * dh
        h = self.buf("s.h", (T, d))
        qkv, att, xmid, gu = b["qkv"], b["att"], b["xmid"], b["gu"]

        def lora_t(xin, A, key, cols):
            """ts = bf16(s * xin A^T): kept in the saved set when the backward will need it, else in scratch"""
            ts = b[key] if key in b else self.buf(f"l.ts.{cols}", (T, cols))
            ops.gemm(xin, A, out=ts, alpha=cfg.lora_scale)
            return ts

        # every adapted linear is ONE launch: y = x W^T + ts B^T (second operand pair of the GEMM, no u in HBM)
        ops.rmsnorm_fwd(x, base[f"L{i}.ln1"], cfg.rms_eps, out=h, rstd=b["rstd1"])
        wqkv = base[f"L{i}.wqkv"]
        if lora is None:
            ops.gemm(h, wqkv, out=qkv)
        else:
            ts = lora_t(h, lora[f"L{i}.qkv.A"], "ts_qkv", 3 * r)
            for j, (lo, hi, n) in enumerate(((0, hd, "q"), (hd, hd + kvd, "k"), (hd + kvd, hd + 2 * kvd, "v"))):
                ops.gemm(h, wqkv[lo:hi], a2=ts[:, j * r:(j + 1) * r], b2=lora[f"L{i}.{n}.B"], out=qkv[:, lo:hi])
        ops.rope_(qkv, m.pos, self.rope_cos, self.rope_sin, H + KV, dh)
        attn_forward(m, qkv[:, :hd], qkv[:, hd:hd + kvd], qkv[:, hd + kvd:], att, b["lse"], H, KV, dh, 1.0 / math.sqrt(dh))
        if lora is None:
            ops.gemm(att, base[f"L{i}.wo"], out=xmid, residual=x)
        else:
            ts = lora_t(att, lora[f"L{i}.o.A"], "ts_o", r)
            ops.gemm(att, base[f"L{i}.wo"], a2=ts, b2=lora[f"L{i}.o.B"], out=xmid, residual=x)
        ops.rmsnorm_fwd(xmid, base[f"L{i}.ln2"], cfg.rms_eps, out=h, rstd=b["rstd2"])
        wgu = base[f"L{i}.wgu"]
        act = self.buf("s.act", (T, ff))
        if lora is None:   # reference pass: SwiGLU in the GEMM epilogue, gate|up never reaches HBM
            ops.gemm_swiglu(h, wgu, gu, act, write_gu=False)
        else:
            ts = lora_t(h, lora[f"L{i}.gu.A"], "ts_gu", 2 * r)
            ops.gemm(h, wgu[:ff], a2=ts[:, :r], b2=lora[f"L{i}.g.B"], out=gu[:, :ff])
            ops.gemm(h, wgu[ff:], a2=ts[:, r:], b2=lora[f"L{i}.u.B"], out=gu[:, ff:])
        if xn is not None or (lora is not None and "ts_d" in b):
            if lora is not None:
                ops.swiglu_fwd(gu, act)
            if lora is not None:
                ts = lora_t(act, lora[f"L{i}.d.A"], "ts_d", r)
            if xn is not None:
                if lora is None:
                    ops.gemm(act, base[f"L{i}.wd"], out=xn, residual=xmid)
                else:
                    ops.gemm(act, base[f"L{i}.wd"], a2=ts, b2=lora[f"L{i}.d.B"], out=xn, residual=xmid)

    # ------------------------------------------------------------------ forward of one pass
    def _forward(self, w, m, feats, tag: str, save: bool, ddpo_weight):
        cfg, base = self.cfg, self.base
        d, T = cfg.hidden, m.T
        lora = self.policy if tag == "policy" else None
        x = self._merged_embeddings(base, m, feats, False, "x.0" if save else "s.x0")   # frozen projector / embeddings
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
        H, KV, dh = cfg.heads, cfg.kv_heads, cfg.head_dim
        hd, kvd = H * dh, KV * dh
        s = cfg.lora_scale
        dx = self._head_backward(grad_logps, base["norm"], base["lm_head"], self._dw_scratch, None)
        dxf = self._bufs["b.dxf"]
        dx2 = self.buf("b.dx1", (T, d))
        h = self.buf("s.h", (T, d))
        act = self.buf("s.act", (T, ff))
        dact = None if self.fuse_swiglu_bwd else self.buf("b.dact", (T, ff))
        dnorm = dxf
        dqkv = self.buf("b.dqkv", (T, cfg.qkv_dim))
        datt = self.buf("b.datt", (T, hd))
        delta = self.buf("b.delta", (m.n_attn_seq, H, m.S), torch.float32)
        # dt = bf16(s * dy B) of the adapters that share an input, side by side (3r: q|k|v, 2r: gate|up, r: o, down)
        d3 = self.buf("b.d3", (T, 3 * r)); d2 = self.buf("b.d2", (T, 2 * r)); d1 = self.buf("b.d1", (T, r))
        scale = 1.0 / math.sqrt(dh)
        for i in reversed(range(cfg.layers)):
            x_in = self._bufs[f"x.{i}"]
            if self.tc.activation_checkpointing:
                sb = self._layer_bufs("a", ".ckpt", m)   # one recompute set incl. the LoRA intermediates
                self._layer_fwd(None, i, x_in, sb, m, None, lora)   # recomputes up to gu + act + ts_d
            else:
                sb = self._layer_bufs("a", f".{i}", m)
                if not self.fuse_swiglu_bwd:
                    ops.swiglu_fwd(sb["gu"], act)                    # recompute act
            xmid, gu, qkv, att = (sb[k] for k in ("xmid", "gu", "qkv", "att"))
            rstd1, rstd2, lse = (sb[k] for k in ("rstd1", "rstd2", "lse"))
            # ---- down_proj
            ops.gemm(dx, sb["ts_d"], a_kmajor=False, b_kmajor=False, out=g[f"L{i}.d.B"], accumulate=acc)      # dBd = dx^T ts_d
            ops.gemm(dx, lora[f"L{i}.d.B"], b_kmajor=False, out=d1, alpha=s)                  # dt = s dx Bd
            if self.fuse_swiglu_bwd:
                # dact = dx Wd + dt Ad stays in TMEM: SwiGLU backward in the epilogue (gu <- dgu in place); act is recomputed
                # there too unless the checkpointed forward above already produced it
                ops.gemm_swiglu_bwd(dx, base[f"L{i}.wd"], gu, None if self.tc.activation_checkpointing else act,
                                    a2=d1, b2=lora[f"L{i}.d.A"])
                ops.gemm(d1, act, a_kmajor=False, b_kmajor=False, out=g[f"L{i}.d.A"], accumulate=acc)         # dAd = dt^T act
            else:
                ops.gemm(d1, act, a_kmajor=False, b_kmajor=False, out=g[f"L{i}.d.A"], accumulate=acc)         # dAd = dt^T act
                ops.gemm(dx, base[f"L{i}.wd"], b_kmajor=False, a2=d1, b2=lora[f"L{i}.d.A"], out=dact)   # dact = dx Wd + dt Ad
            # ---- gate | up
            ops.rmsnorm_fwd(xmid, base[f"L{i}.ln2"], cfg.rms_eps, out=h)                      # recompute h2
            if not self.fuse_swiglu_bwd:
                ops.swiglu_bwd(gu, dact, out=gu)                                              # dgu (in place)
            tsg = sb["ts_gu"]
            ops.gemm(gu[:, :ff], tsg[:, :r], a_kmajor=False, b_kmajor=False, out=g[f"L{i}.g.B"], accumulate=acc)
            ops.gemm(gu[:, ff:], tsg[:, r:], a_kmajor=False, b_kmajor=False, out=g[f"L{i}.u.B"], accumulate=acc)
            ops.gemm(gu[:, :ff], lora[f"L{i}.g.B"], b_kmajor=False, out=d2[:, :r], alpha=s)
            ops.gemm(gu[:, ff:], lora[f"L{i}.u.B"], b_kmajor=False, out=d2[:, r:], alpha=s)
            ops.gemm(d2, h, a_kmajor=False, b_kmajor=False, out=g[f"L{i}.gu.A"], accumulate=acc)              # dA = dt^T h2  [2r, d]
            ops.gemm(gu, base[f"L{i}.wgu"], b_kmajor=False, a2=d2, b2=lora[f"L{i}.gu.A"], out=dnorm)   # dh2 = dgu Wgu + dt A
            ops.rmsnorm_bwd(dnorm, xmid, base[f"L{i}.ln2"], rstd2, self._dw_scratch, dres=dx, out=dx2)   # dxmid
            # ---- o_proj
            ops.gemm(dx2, sb["ts_o"], a_kmajor=False, b_kmajor=False, out=g[f"L{i}.o.B"], accumulate=acc)
            ops.gemm(dx2, lora[f"L{i}.o.B"], b_kmajor=False, out=d1, alpha=s)
            ops.gemm(d1, att, a_kmajor=False, b_kmajor=False, out=g[f"L{i}.o.A"], accumulate=acc)
            ops.gemm(dx2, base[f"L{i}.wo"], b_kmajor=False, a2=d1, b2=lora[f"L{i}.o.A"], out=datt)   # datt = dxmid Wo + dt Ao
            attn_backward(m, qkv[:, :hd], qkv[:, hd:hd + kvd], qkv[:, hd + kvd:], att, datt, lse, delta, dqkv[:, :hd],
                          dqkv[:, hd:hd + kvd], dqkv[:, hd + kvd:], H, KV, dh, scale)
            ops.rope_(dqkv, m.pos, self.rope_cos, self.rope_sin, H + KV, dh, inverse=True)
            # ---- q | k | v
            ops.rmsnorm_fwd(x_in, base[f"L{i}.ln1"], cfg.rms_eps, out=h)                      # recompute h1
            tsq = sb["ts_qkv"]
            cols = ((0, hd, "q"), (hd, hd + kvd, "k"), (hd + kvd, hd + 2 * kvd, "v"))
            for j, (lo, hi, n) in enumerate(cols):
                ops.gemm(dqkv[:, lo:hi], tsq[:, j * r:(j + 1) * r], a_kmajor=False, b_kmajor=False, out=g[f"L{i}.{n}.B"], accumulate=acc)
                ops.gemm(dqkv[:, lo:hi], lora[f"L{i}.{n}.B"], b_kmajor=False, out=d3[:, j * r:(j + 1) * r], alpha=s)
            ops.gemm(d3, h, a_kmajor=False, b_kmajor=False, out=g[f"L{i}.qkv.A"], accumulate=acc)             # [3r, d]
            if i > 0:  # nothing below decoder layer 0 is trainable: its input gradient is never read
                ops.gemm(dqkv, base[f"L{i}.wqkv"], b_kmajor=False, a2=d3, b2=lora[f"L{i}.qkv.A"], out=dnorm)
                ops.rmsnorm_bwd(dnorm, x_in, base[f"L{i}.ln1"], rstd1, self._dw_scratch, dres=dx2, out=dx)
            self._reduce_bucket(self.layout.offsets[f"L{i}.qkv.A"],
                                self.layout.offsets[f"L{i + 1}.qkv.A"] if i + 1 < cfg.layers else self.layout.size)
