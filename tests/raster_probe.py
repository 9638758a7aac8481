"""Development probe (not a pytest file): every distinct GEMM of the config-2 step (LLaVA-1.5-7B, T = 12792 rows) under a
sweep of the rasterisation budget, one process, CUDA events over back-to-back launches, two passes to show the noise."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlrlhf_b200  # noqa: E402,F401
from vlrlhf_b200 import ops  # noqa: E402

dev, bf = "cuda", torch.bfloat16
T, d, ff, V, R = 12792, 4096, 11008, 32064, 8184


def rnd(*shape):
    return (torch.randn(*shape, device=dev) * 0.05).to(bf)


x, x2 = rnd(T, d), rnd(T, d)
w_qkv, w_o, w_gu, w_d, w_lm = rnd(3 * d, d), rnd(d, d), rnd(2 * ff, d), rnd(d, ff), rnd(V, d)
t_qkv, t_gu, t_ff, t_r, t_v = rnd(T, 3 * d), rnd(T, 2 * ff), rnd(T, ff), rnd(R, d), rnd(R, V)
o_qkv, o_d, o_gu, o_ff, o_act = (torch.empty(T, n, dtype=bf, device=dev) for n in (3 * d, d, 2 * ff, ff, ff))
o_f32 = torch.empty(T, d, dtype=torch.float32, device=dev)
o_logits = torch.empty(R, V, dtype=torch.float32, device=dev)
g_qkv, g_o, g_gu, g_d, g_lm, o_r = (torch.empty_like(t) for t in (w_qkv, w_o, w_gu, w_d, w_lm, t_r))
cases = [  # name, launches per step, flops, fn
    ("fwd qkv", 64, 2.0 * T * 3 * d * d, lambda: ops.gemm(x, w_qkv, out=o_qkv)),
    ("fwd o+res", 64, 2.0 * T * d * d, lambda: ops.gemm(x, w_o, out=o_f32, residual=o_f32)),
    ("fwd gate_up swiglu", 64, 2.0 * T * 2 * ff * d, lambda: ops.gemm_swiglu(x, w_gu, o_gu, o_act, write_gu=True)),
    ("fwd down+res", 64, 2.0 * T * d * ff, lambda: ops.gemm(t_ff, w_d, out=o_f32, residual=o_f32)),
    ("fwd lm_head", 2, 2.0 * R * V * d, lambda: ops.gemm(t_r, w_lm, out=o_logits)),
    ("dgrad qkv", 32, 2.0 * T * 3 * d * d, lambda: ops.gemm(t_qkv, w_qkv, b_kmajor=False, out=o_d)),
    ("dgrad o", 32, 2.0 * T * d * d, lambda: ops.gemm(x, w_o, b_kmajor=False, out=o_d)),
    ("dgrad gate_up", 32, 2.0 * T * 2 * ff * d, lambda: ops.gemm(t_gu, w_gu, b_kmajor=False, out=o_d)),
    ("dgrad down", 32, 2.0 * T * d * ff, lambda: ops.gemm(x, w_d, b_kmajor=False, out=o_ff)),
    ("dgrad lm_head", 1, 2.0 * R * V * d, lambda: ops.gemm(t_v, w_lm, b_kmajor=False, out=o_r)),
    ("wgrad qkv", 32, 2.0 * T * 3 * d * d, lambda: ops.gemm(t_qkv, x, a_kmajor=False, b_kmajor=False, out=g_qkv)),
    ("wgrad o", 32, 2.0 * T * d * d, lambda: ops.gemm(x, x2, a_kmajor=False, b_kmajor=False, out=g_o)),
    ("wgrad gate_up", 32, 2.0 * T * 2 * ff * d, lambda: ops.gemm(t_gu, x, a_kmajor=False, b_kmajor=False, out=g_gu)),
    ("wgrad down", 32, 2.0 * T * d * ff, lambda: ops.gemm(x, t_ff, a_kmajor=False, b_kmajor=False, out=g_d)),
    ("wgrad lm_head", 1, 2.0 * R * V * d, lambda: ops.gemm(t_v, t_r, a_kmajor=False, b_kmajor=False, out=g_lm)),
]
# settings: raster budgets in MB (policy 0), or "model" = the LRU-model policy (vlb200_set_gemm_raster_policy(1))
budgets = [b if b == "model" else float(b) for b in (sys.argv[1].split(",") if len(sys.argv) > 1 else
                                                     "32,model,16,24,48,64,96,128,model,32".split(","))]
iters = 16
res = {}
for bi, mb in enumerate(budgets):
    ops.set_gemm_raster_policy(1 if mb == "model" else 0)
    if mb != "model":
        ops.set_gemm_raster_mb(mb)
    for name, per_step, fl, fn in cases:
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res.setdefault(name, []).append(round(e0.elapsed_time(e1) / iters, 4))
print("budgets_mb", budgets)
step = [0.0] * len(budgets)
for name, per_step, fl, fn in cases:
    ms = res[name]
    print(f"{name:20s} x{per_step:3d} " + " ".join(f"{m:7.3f}" for m in ms) + f"   best {budgets[ms.index(min(ms))]}", flush=True)
    for i, m in enumerate(ms):
        step[i] += per_step * m
print(f"{'GEMM ms per step':24s} " + " ".join(f"{m:7.1f}" for m in step))
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"budgets_mb": budgets, "ms": res, "launches_per_step": {c[0]: c[1] for c in cases}, "gemm_ms_per_step": step},
          open("gpurun_out/raster_probe.json", "w"), indent=1)
