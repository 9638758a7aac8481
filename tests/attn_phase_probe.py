"""Attention phase probe (not a pytest file): runs the DBG-8 variants of the tcgen05 attention kernels (phase counters in
one softmax / elementwise thread per CTA) at the config-2 shape and prints the average cycles per tile spent in each phase.
    VLB200_ATTN_FWD_VARIANT=80|82 VLB200_ATTN_BWD_TS=1 VLB200_ATTN_BWD_DBG=8 python tests/attn_phase_probe.py"""
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
fwd = lambda: ops.attn_fwd_tc(q, k, v, out, lse, ld, n, S, H, KV, dh, True, sc)  # noqa: E731
bwd = lambda: ops.attn_bwd_tc(q, k, v, out, dout, lse, delta, dqkv[:, :H * dh], dqkv[:, H * dh:(H + KV) * dh],  # noqa: E731
                              dqkv[:, (H + KV) * dh:], ld, n, S, H, KV, dh, True, sc)


def run(fn, reader, width, names, reps=5):
    buf = (ctypes.c_ulonglong * width)()
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    reader(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    reader(buf, 1)
    ms = e0.elapsed_time(e1) / reps
    for base in range(0, width, 16):
        vals = [buf[base + i] / reps for i in range(10)]
        items, tiles = vals[8], vals[9]
        if tiles == 0:
            continue
        extra = [buf[base + i] / reps for i in (10, 11, 12, 13, 14, 15)]
        tot = sum(vals[:8]) + sum(extra)
        print(f"  [{names[base // 16]}] {ms:.3f} ms/launch-set; per CTA: {items / 148:.1f} items, {tiles / 148:.1f} tiles, {tot / 148:.0f} cycles in the loop")
        if any(extra):
            print(f"      item: wait for S of the first tile {extra[0] / items:8.0f} cycles per item; wait for the last PV {extra[1] / items:8.0f} cycles per item")
            if extra[2] or extra[3]:
                print(f"      item: epilogue: row-sum exchange + set-up {extra[2] / items:8.0f}; O chunks: TMEM load {extra[4] / items:6.0f}, scale/pack/staging {extra[5] / items:6.0f}, staging load + global store {extra[3] / items:6.0f} cycles per item")
        for i, nm in enumerate(names[-1]):
            if vals[i]:
                per = vals[i] / (items if nm.startswith("item:") else tiles)
                print(f"      {nm:52s} {per:8.0f} cycles per {'item' if nm.startswith('item:') else 'tile'}   ({100 * vals[i] / tot:4.1f} %)")


if os.environ.get("VLB200_ATTN_FWD_VARIANT", "0") in ("80", "81", "82", "84", "85", "125", "105", "245", "265"):
    run(fwd, lib.vlbdbg_attn_fwd_profile, 16, ["forward", ["item: start (plan, Q copy)", "wait for S", "TMEM -> registers (S)", "mask + row max",
                                                          "exp2, row sum, pack, P store issue", "wait for the previous PV",
                                                          "O correction, store completion, fences, publish", "item: epilogue"]])
if os.environ.get("VLB200_ATTN_BWD_DBG", "0") == "8":
    run(bwd, lib.vlbdbg_attn_bwd_profile, 32, ["backward dK/dV", "backward dQ", ["item: start", "tile coordinates, statistics barrier + prefetch",
                                                                             "wait for the score tiles", "TMEM -> registers", "elementwise",
                                                                             "E store, completion, fences, publish", "", "item: epilogue"]])
