"""Standalone launches of the dominant kernels at config-2 shapes for `ncu --set full` (one GPU, short)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlrlhf_b200  # noqa: E402,F401
from vlrlhf_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "gemm"
dev, bf = "cuda", torch.bfloat16
T, d, ff, V, H, dh, S, B = 12792, 4096, 11008, 32064, 32, 128, 1599, 8
if which == "gemm":
    x = torch.randn(T, d, device=dev).to(bf)
    w = torch.randn(2 * ff, d, device=dev).to(bf) * 0.02
    out = torch.empty(T, 2 * ff, dtype=bf, device=dev)
    for _ in range(4):
        ops.gemm(x, w, out=out)
elif which == "gemm_swiglu":   # gate|up GEMM with SwiGLU in the epilogue (reference pass: gate|up never written)
    x = torch.randn(T, d, device=dev).to(bf)
    w = torch.randn(2 * ff, d, device=dev).to(bf) * 0.02
    gu = torch.empty(T, 2 * ff, dtype=bf, device=dev)
    act = torch.empty(T, ff, dtype=bf, device=dev)
    for _ in range(4):
        ops.gemm_swiglu(x, w, gu, act, write_gu=False)
elif which == "lora":          # LoRA linear in one launch (second operand pair) and a split-K weight gradient
    r = 128
    x = torch.randn(T, d, device=dev).to(bf)
    w = torch.randn(ff, d, device=dev).to(bf) * 0.02
    ts = torch.randn(T, 2 * r, device=dev).to(bf)
    Bm = torch.randn(ff, r, device=dev).to(bf) * 0.02
    out = torch.empty(T, 2 * ff, dtype=bf, device=dev)
    dy = torch.randn(T, d, device=dev).to(bf)
    dB = torch.empty(d, r, dtype=bf, device=dev)
    for _ in range(4):
        ops.gemm(x, w, a2=ts[:, :r], b2=Bm, out=out[:, :ff])
        ops.gemm(dy, ts[:, :r], a_kmajor=False, b_kmajor=False, out=dB)
elif which == "attn_tc":
    qkv = torch.randn(T, 3 * d, device=dev).to(bf)
    out = torch.empty(T, d, dtype=bf, device=dev)
    lse = torch.empty(B, H, S, dtype=torch.float32, device=dev)
    sl = torch.tensor([1599, 1400, 1599, 1500, 1599, 1300, 1450, 1599], dtype=torch.int32, device=dev)
    sc = 1 / math.sqrt(dh)
    dout = torch.randn(T, d, device=dev).to(bf)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty_like(lse)
    for _ in range(3):
        ops.attn_fwd_tc(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], out, lse, sl, B, S, H, H, dh, True, sc)
        ops.attn_bwd_tc(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], out, dout, lse, delta, dqkv[:, :d], dqkv[:, d:2 * d],
                        dqkv[:, 2 * d:], sl, B, S, H, H, dh, True, sc)
elif which == "logps":
    R = 8 * 1023
    logits = torch.randn(R, V, device=dev)
    tgt = torch.randint(0, V, (R,), device=dev)
    for _ in range(3):
        _, _, lse = ops.logps_fwd(logits, tgt, 8)
        ops.logps_bwd(logits, tgt, 8, lse, torch.ones(8, device=dev))
torch.cuda.synchronize()
print("done", which)
