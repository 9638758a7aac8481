"""Timing probe (not a pytest file): every GEMM shape one decoder layer of the LLaVA-1.5-7B + LoRA r=128 step issues
(engine_lora.py), timed alone with CUDA events; prints us, TFLOP/s and the GB/s of the operands/outputs it must move."""
import sys

import torch

T, d, ff, r = 12792, 4096, 11008, 128
qkv = 3 * d
# name, M, N, K, a_kmajor, b_kmajor, out dtype bytes, accumulate, residual, count per layer
SHAPES = [
    ("base qkv fwd (+u residual)", T, qkv, d, 1, 1, 2, 0, 1, 1),
    ("base gu fwd (+u residual)", T, 2 * ff, d, 1, 1, 2, 0, 1, 1),
    ("base down fwd (fp32 resid)", T, d, ff, 1, 1, 4, 0, 2, 1),
    ("t3 = h A3^T", T, 3 * r, d, 1, 1, 4, 0, 0, 1),
    ("t2 = h A2^T", T, 2 * r, d, 1, 1, 4, 0, 0, 1),
    ("t1 = x A^T (K=d)", T, r, d, 1, 1, 4, 0, 0, 1),
    ("t1 = act A^T (K=ff)", T, r, ff, 1, 1, 4, 0, 0, 1),
    ("u = ts B^T (N=d)", T, d, r, 1, 1, 2, 0, 0, 3),
    ("u = ts B^T (N=ff)", T, ff, r, 1, 1, 2, 0, 0, 2),
    ("x += ts B^T (fp32 acc, N=d)", T, d, r, 1, 1, 4, 1, 0, 2),
    ("dB = dy^T ts (M=d)", d, r, T, 0, 0, 2, 0, 0, 5),
    ("dB = dy^T ts (M=ff)", ff, r, T, 0, 0, 2, 0, 0, 2),
    ("dt = dy B (K=d)", T, r, d, 1, 0, 4, 0, 0, 5),
    ("dt = dy B (K=ff)", T, r, ff, 1, 0, 4, 0, 0, 2),
    ("dA = dt^T x (M=r, N=d)", r, d, T, 0, 0, 2, 0, 0, 1),
    ("dA = dt^T x (M=2r, N=d)", 2 * r, d, T, 0, 0, 2, 0, 0, 1),
    ("dA = dt^T x (M=3r, N=d)", 3 * r, d, T, 0, 0, 2, 0, 0, 1),
    ("dA = dt^T act (M=r, N=ff)", r, ff, T, 0, 0, 2, 0, 0, 1),
    ("dx += dt A (N=d, K=r)", T, d, r, 1, 0, 2, 1, 0, 1),
    ("dx += dt A (N=d, K=2r)", T, d, 2 * r, 1, 0, 2, 1, 0, 1),
    ("dx += dt A (N=d, K=3r)", T, d, 3 * r, 1, 0, 2, 1, 0, 1),
    ("dx += dt A (N=ff, K=r)", T, ff, r, 1, 0, 2, 1, 0, 1),
    ("base dgrad down", T, ff, d, 1, 0, 2, 0, 0, 1),
    ("base dgrad gu", T, d, 2 * ff, 1, 0, 2, 0, 0, 1),
    ("base dgrad o", T, d, d, 1, 0, 2, 0, 0, 1),
    ("base dgrad qkv", T, d, qkv, 1, 0, 2, 0, 0, 1),
]


def main():
    import vlrlhf_b200  # noqa: F401
    from vlrlhf_b200 import ops
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    total = 0.0
    for name, M, N, K, ak, bk, ob, acc, res, count in SHAPES:
        a = torch.randn((M, K) if ak else (K, M), device=dev).to(torch.bfloat16)
        b = torch.randn((N, K) if bk else (K, N), device=dev).to(torch.bfloat16)
        out = torch.zeros(M, N, device=dev, dtype=torch.float32 if ob == 4 else torch.bfloat16)
        kw = {}
        if res == 1:
            kw["residual"] = torch.randn(M, N, device=dev).to(torch.bfloat16)
        elif res == 2:
            kw["residual"] = torch.randn(M, N, device=dev)
        fn = lambda: ops.gemm(a, b, a_kmajor=bool(ak), b_kmajor=bool(bk), out=out, accumulate=bool(acc), **kw)  # noqa: E731
        for _ in range(3):
            fn()
        ts = []
        for _ in range(8):
            flush.zero_()   # L2 flush between timed iterations
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        byts = 2 * (M * K + N * K) + ob * M * N * (2 if acc else 1) + (2 * M * N if res == 1 else 4 * M * N if res == 2 else 0)
        print(f"{name:34s} M={M:6d} N={N:6d} K={K:6d} x{count}  {ms * 1e3:8.1f} us  {2.0 * M * N * K / ms / 1e9:7.1f} TF/s  "
              f"{byts / ms / 1e6:7.0f} GB/s  -> {ms * count:6.3f} ms/layer", flush=True)
        total += ms * count
    print(f"sum per layer {total:.3f} ms  (x32 layers = {total * 32:.1f} ms)")


if __name__ == "__main__":
    sys.exit(main())
