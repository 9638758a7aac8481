"""Per-tile / per-item cost of the attention forward (not a pytest file): non-causal, B*H = 37 so that every CTA of the 148 gets
the same number of equal items: time = items_per_cta * (F + tiles * c)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlrlhf_b200  # noqa: E402,F401
from vlrlhf_b200 import ops  # noqa: E402

dev, bf = "cuda", torch.bfloat16
H = KV = 37
dh = 128
sc = 1 / math.sqrt(dh)
res = []
for S in (512, 1024, 2048, 4096):
    qkv = (torch.randn(S, 3 * H * dh, device=dev) * 0.5).to(bf)
    q, k, v = qkv[:, :H * dh], qkv[:, H * dh:2 * H * dh], qkv[:, 2 * H * dh:]
    out = torch.empty(S, H * dh, dtype=bf, device=dev)
    lse = torch.zeros(1, H, S, dtype=torch.float32, device=dev)
    fn = lambda: ops.attn_fwd_tc(q, k, v, out, lse, None, 1, S, H, KV, dh, False, sc)  # noqa: E731
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    nq = S // 128
    items_per_cta = H * nq / 148
    fl = 4.0 * S * S * dh * H
    res.append((S, ms, items_per_cta, nq))
    print(f"S={S}: {ms * 1e3:.1f} us  items/CTA {items_per_cta:.2f} x {nq} tiles  {fl / ms / 1e9:.0f} TFLOP/s", flush=True)
# fit F, c (microseconds) from the two largest
(_, t1, i1, n1), (_, t2, i2, n2) = res[-2], res[-1]
# t = i * (F + n c)
c = (t2 / i2 - t1 / i1) / (n2 - n1)
F = t1 / i1 - n1 * c
print(f"per tile c = {c * 1e3:.3f} us, per item F = {F * 1e3:.3f} us")
