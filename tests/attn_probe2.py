"""Attention kernel probe (not a pytest file): times the tcgen05 forward / backward at the padded config-2 shape and at the
shared-prefix layout of the bench batch, prints TFLOP/s (causal-half FLOPs; backward = 2.5x forward).
    python tests/attn_probe2.py [time|ncu]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlrlhf_b200  # noqa: E402,F401
from vlrlhf_b200 import ops  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "time"
dev, bf = "cuda", torch.bfloat16
H, KV, dh = 32, 32, 128
sc = 1 / math.sqrt(dh)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def case(name, lens, starts, ctx, kids, S, reps):
    T, B = starts[-1], len(lens)
    torch.manual_seed(0)
    qkv = (torch.randn(T, (H + 2 * KV) * dh, device=dev) * 0.5).to(bf)
    q, k, v = qkv[:, :H * dh], qkv[:, H * dh:(H + KV) * dh], qkv[:, (H + KV) * dh:]
    out = torch.empty(T, H * dh, dtype=bf, device=dev)
    dout = (torch.randn(T, H * dh, device=dev) * 0.1).to(bf)
    dqkv = torch.empty_like(qkv)
    lse = torch.zeros(B, H, S, dtype=torch.float32, device=dev)
    delta = torch.zeros_like(lse)
    i32 = lambda x: torch.tensor(x, dtype=torch.int32, device=dev)  # noqa: E731
    kw = dict(row_starts=i32(starts), total_rows=T)
    if ctx is not None:
        kw.update(ctx=i32(ctx), kids=i32(kids))
    ld = i32(lens)
    fwd = lambda: ops.attn_fwd_tc(q, k, v, out, lse, ld, B, S, H, KV, dh, True, sc, **kw)  # noqa: E731
    bwd = lambda: ops.attn_bwd_tc(q, k, v, out, dout, lse, delta, dqkv[:, :H * dh], dqkv[:, H * dh:(H + KV) * dh],  # noqa: E731
                                  dqkv[:, (H + KV) * dh:], ld, B, S, H, KV, dh, True, sc, **kw)
    fl = 0.0
    for b in range(B):
        c = lens[ctx[b]] if ctx is not None and ctx[b] >= 0 else 0
        fl += 4.0 * dh * H * (lens[b] * lens[b] / 2 + lens[b] * c)
    if mode == "ncu":
        for _ in range(reps):
            fwd(); bwd()
        torch.cuda.synchronize()
        return
    tf, tb = timeit(fwd), timeit(bwd)
    print(f"[{name}] rows {T} seqs {B}: fwd {tf:.3f} ms = {fl / tf / 1e9:.0f} TFLOP/s   bwd {tb:.3f} ms = {2.5 * fl / tb / 1e9:.0f} TFLOP/s",
          flush=True)


# padded config 2 as packed rows of full length (8 x 1599)
n = 8
lens = [1599] * n
starts = [i * 1599 for i in range(n + 1)]
case("config2 8x1599", lens, starts, None, None, 1599, 2)
# the bench batch with shared prefixes: 4 pairs, prefix 703, chosen suffix 896, rejected suffixes ragged
pre, sc_, sr = [703] * 4, [896] * 4, [896, 810, 720, 850]
lens = sc_ + pre + sr
starts = [0]
for x in lens:
    starts.append(starts[-1] + x)
ctx = [4 + i for i in range(4)] + [-1] * 4 + [4 + i for i in range(4)]
kids = [-1] * 24
for i in range(4):
    kids[2 * (4 + i)], kids[2 * (4 + i) + 1] = i, 8 + i
case("shared 4 pairs", lens, starts, ctx, kids, 1599, 2)
