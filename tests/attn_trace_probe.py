"""Backward attention pipeline trace (not a pytest file): VLB200_ATTN_BWD_DBG=16 python tests/attn_trace_probe.py
clock64 timestamps of CTA 0's first tile iterations, printed relative to the first event, per pass."""
import ctypes
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlrlhf_b200  # noqa: E402,F401
from vlrlhf_b200 import _lib, ops  # noqa: E402

lib = ctypes.CDLL(_lib.LIB_PATH)
dev, bf = "cuda", torch.bfloat16
H, KV, dh, n, S = 32, 32, 128, 8, 1599
sc = 1 / math.sqrt(dh)
T = n * S
torch.manual_seed(0)
qkv = (torch.randn(T, (H + 2 * KV) * dh, device=dev) * 0.5).to(bf)
q, k, v = qkv[:, :H * dh], qkv[:, H * dh:(H + KV) * dh], qkv[:, (H + KV) * dh:]
out = torch.empty(T, H * dh, dtype=bf, device=dev)
dout = (torch.randn(T, H * dh, device=dev) * 0.1).to(bf)
dqkv = torch.empty_like(qkv)
lse = torch.zeros(n, H, S, dtype=torch.float32, device=dev)
delta = torch.zeros_like(lse)
ld = torch.full((n,), S, dtype=torch.int32, device=dev)
ops.attn_fwd_tc(q, k, v, out, lse, ld, n, S, H, KV, dh, True, sc)
for _ in range(3):
    ops.attn_bwd_tc(q, k, v, out, dout, lse, delta, dqkv[:, :H * dh], dqkv[:, H * dh:(H + KV) * dh], dqkv[:, (H + KV) * dh:], ld, n, S, H, KV,
                    dh, True, sc)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 4096)()
assert lib.vlbdbg_attn_bwd_trace(buf) == 0
names = ["ew waits T", "ew sees T", "ew publishes E", "mma sees E", "mma issued acc", "mma sees Y", "mma issued score"]
for mode in range(2):
    ev = [[buf[(mode * 32 + e) * 64 + i] for i in range(64)] for e in range(32)]
    t0 = min(x for row in ev[:7] for x in row if x > 0)
    print(f"== pass {mode} ({'dK/dV' if mode == 0 else 'dQ'}): cycles since the first event; tiles 8..23")
    print("tile " + " ".join(f"{nm:>16s}" for nm in names))
    for i in range(8, 16):
        print(f"{i:4d} " + " ".join(f"{ev[e][i] - t0:16d}" for e in range(7)))
    # derived latencies (averages over tiles 8..40)
    rng = range(8, 40)
    avg = lambda f: sum(f(i) for i in rng) / len(rng)  # noqa: E731
    print(f"  E published -> MMA warp sees it      : {avg(lambda i: ev[3][i] - ev[2][i]):7.0f}")
    print(f"  MMA sees E -> accumulate issued      : {avg(lambda i: ev[4][i] - ev[3][i]):7.0f}")
    print(f"  accumulate(t) issued -> score(t+2) issued : {avg(lambda i: ev[6][i + 2] - ev[4][i]):7.0f}")
    print(f"  score(t) issued -> elementwise sees T(t)  : {avg(lambda i: ev[1][i] - ev[6][i]):7.0f}")
    print(f"  elementwise: sees T -> publishes E   : {avg(lambda i: ev[2][i] - ev[1][i]):7.0f}")
    print(f"  elementwise: waits for T             : {avg(lambda i: ev[1][i] - ev[0][i]):7.0f}")
    for i in (12, 13, 20):
        base = ev[3][i]
        print(f"  tile {i}: warps see T at " + " ".join(f"{ev[16 + w][i] - base:6d}" for w in range(8)) + "   publish E at " +
              " ".join(f"{ev[8 + w][i] - base:6d}" for w in range(8)) + "  (relative to 'MMA sees E')")
    print("  MMA warp, per tile: sees E | first acc MMA | last acc MMA issued | after commit+syncwarp | sees Y(t+2) | first score MMA | last score MMA | done")
    for i in range(10, 16):
        b0 = ev[3][i]
        seq = [ev[3][i], ev[24][i], ev[25][i], ev[4][i], ev[5][i + 2], ev[26][i + 2], ev[27][i + 2], ev[6][i + 2]]
        print(f"    tile {i}: " + " ".join(f"{x - b0:6d}" for x in seq) + f"   next 'sees E' {ev[3][i + 1] - b0:6d}")
    print(f"  tile period                          : {avg(lambda i: ev[2][i + 1] - ev[2][i]):7.0f}")
