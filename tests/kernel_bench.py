"""Per-kernel timing at the config-2 shapes (LLaVA-1.5-7B, 8 sequences x 1599 merged tokens) -- a development
probe (not a pytest file, not the bench contract).  CUDA events on the launching stream, L2 flushed between
iterations by cycling through buffers larger than L2 where cheap."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlrlhf_b200  # noqa: E402,F401
from vlrlhf_b200 import ops  # noqa: E402


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = "cuda"
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf_peak = peaks.get("bf16_tflops", 1590.0)
    bw_peak = peaks.get("hbm_gbs", 6650.0)
    T, d, ff, V, H, dh, S, B = 12792, 4096, 11008, 32064, 32, 128, 1599, 8
    bf = torch.bfloat16
    x = torch.randn(T, d, device=dev).to(bf)
    res = {}

    def gemm_case(name, a, b, ak, bk, M, N, K, **kw):
        out = torch.empty(M, N, dtype=kw.pop("odt", bf), device=dev)
        ms = timeit(lambda: ops.gemm(a, b, a_kmajor=ak, b_kmajor=bk, out=out, **kw))
        tf = 2.0 * M * N * K / ms / 1e9
        print(f"gemm {name:18s} M={M:6d} N={N:6d} K={K:6d}  {ms:8.3f} ms  {tf:7.1f} TFLOP/s  ({tf / tf_peak:.2f} of measured cuBLAS)", flush=True)
        res[name] = tf

    wqkv = torch.randn(3 * d, d, device=dev).to(bf) * 0.02
    wo = torch.randn(d, d, device=dev).to(bf) * 0.02
    wgu = torch.randn(2 * ff, d, device=dev).to(bf) * 0.02
    wd = torch.randn(d, ff, device=dev).to(bf) * 0.02
    gemm_case("qkv fwd", x, wqkv, True, True, T, 3 * d, d)
    gemm_case("o fwd+res", x, wo, True, True, T, d, d, residual=x)
    gemm_case("gate_up fwd", x, wgu, True, True, T, 2 * ff, d)
    act = torch.randn(T, ff, device=dev).to(bf)
    gemm_case("down fwd+res", act, wd, True, True, T, d, ff, residual=x)
    dgu = torch.randn(T, 2 * ff, device=dev).to(bf)
    gemm_case("gate_up dgrad", dgu, wgu, True, False, T, d, 2 * ff)
    gemm_case("gate_up wgrad", dgu, x, False, False, 2 * ff, d, T)
    gemm_case("down dgrad", x, wd, True, False, T, ff, d)
    gemm_case("down wgrad", x, act, False, False, d, ff, T)
    R = 8 * 1023
    hsel = torch.randn(R, d, device=dev).to(bf)
    wlm = torch.randn(V, d, device=dev).to(bf) * 0.02
    gemm_case("lm_head f32out", hsel, wlm, True, True, R, V, d, odt=torch.float32)
    vx = torch.randn(4 * 577, 1024, device=dev).to(bf)
    vw = torch.randn(3072, 1024, device=dev).to(bf) * 0.02
    gemm_case("vit qkv", vx, vw, True, True, 4 * 577, 3072, 1024)
    del wgu, wd, dgu, act, wlm, hsel

    # attention
    qkv = torch.randn(T, 3 * d, device=dev).to(bf)
    out = torch.empty(T, d, dtype=bf, device=dev)
    lse = torch.empty(B, H, S, dtype=torch.float32, device=dev)
    seqlens = torch.tensor([1599, 1400, 1599, 1500, 1599, 1300, 1450, 1599], dtype=torch.int32, device=dev)
    sc = 1 / math.sqrt(dh)
    ms = timeit(lambda: ops.attn_fwd_tc(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], out, lse, seqlens, B, S, H, H, dh, True, sc))
    fl = sum(2.0 * 2 * int(l) * int(l) / 2 * dh * H for l in seqlens.tolist())
    print(f"attn fwd  causal B=8 S=1599 H=32 dh=128  {ms:8.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s (causal-half flops)", flush=True)
    dout = torch.randn(T, d, device=dev).to(bf)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty_like(lse)
    ms = timeit(lambda: ops.attn_bwd_tc(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], out, dout, lse, delta, dqkv[:, :d],
                                     dqkv[:, d:2 * d], dqkv[:, 2 * d:], seqlens, B, S, H, H, dh, True, sc))
    print(f"attn bwd  (2.5x fwd flops)               {ms:8.3f} ms  {2.5 * fl / ms / 1e9:7.1f} TFLOP/s", flush=True)
    vq = torch.randn(4 * 577, 3072, device=dev).to(bf)
    vo = torch.empty(4 * 577, 1024, dtype=bf, device=dev)
    ms = timeit(lambda: ops.attn_fwd_tc(vq[:, :1024], vq[:, 1024:2048], vq[:, 2048:], vo, None, None, 4, 577, 16, 16, 64, False, 0.125))
    print(f"attn fwd  vit B=4 S=577 H=16 dh=64        {ms:8.3f} ms  {4 * 16 * 4.0 * 577 * 577 * 64 / ms / 1e9:7.1f} TFLOP/s", flush=True)
    del qkv, dqkv, dout

    # HBM-bound kernels: bytes / time
    def bw_case(name, fn, nbytes):
        ms = timeit(fn)
        gbs = nbytes / ms / 1e6
        print(f"{name:28s} {ms:8.3f} ms  {gbs:8.1f} GB/s  ({gbs / bw_peak:.2f} of measured copy)", flush=True)
        res[name] = gbs

    w = torch.ones(d, dtype=bf, device=dev)
    y = torch.empty_like(x)
    rstd = torch.empty(T, dtype=torch.float32, device=dev)
    bw_case("rmsnorm fwd", lambda: ops.rmsnorm_fwd(x, w, 1e-5, out=y, rstd=rstd), 2 * T * d * 2)
    dw = torch.zeros(d, dtype=bf, device=dev)
    bw_case("rmsnorm bwd (+dres)", lambda: ops.rmsnorm_bwd(y, x, w, rstd, dw, dres=x, out=y), 4 * T * d * 2)
    x32 = x.float()
    y2 = torch.empty_like(x)
    bw_case("rmsnorm fwd (fp32 x)", lambda: ops.rmsnorm_fwd(x32, w, 1e-5, out=y, rstd=rstd), T * d * 6)
    bw_case("rmsnorm bwd (fp32 x,+dres)", lambda: ops.rmsnorm_bwd(y, x32, w, rstd, dw, dres=x, out=y2), T * d * 10)
    del x32, y2
    gu = torch.randn(T, 2 * ff, device=dev).to(bf)
    a2 = torch.empty(T, ff, dtype=bf, device=dev)
    bw_case("swiglu fwd", lambda: ops.swiglu_fwd(gu, a2), 3 * T * ff * 2)
    bw_case("swiglu bwd (in place)", lambda: ops.swiglu_bwd(gu, a2, out=gu), 5 * T * ff * 2)
    del gu, a2
    q2 = torch.randn(T, 3 * d, device=dev).to(bf)
    pos = (torch.arange(T, device=dev) % S).to(torch.int32)
    fr = torch.arange(4096, dtype=torch.float32)[:, None] / (10000.0 ** (torch.arange(0, dh, 2).float() / dh))[None]
    ct, st = fr.cos().to(dev), fr.sin().to(dev)
    bw_case("rope (q,k in place)", lambda: ops.rope_(q2, pos, ct, st, 2 * H, dh), 2 * T * 2 * d * 2)
    del q2
    logits = torch.randn(R, V, device=dev)
    tgt = torch.randint(0, V, (R,), device=dev)
    bw_case("logps fwd f32 logits", lambda: ops.logps_fwd(logits, tgt, 8), R * V * 4)
    _, _, l2 = ops.logps_fwd(logits, tgt, 8)
    g = torch.ones(8, device=dev)
    dl = torch.empty(R, V, dtype=bf, device=dev)
    bw_case("logps bwd f32->bf16", lambda: ops.logps_bwd(logits, tgt, 8, l2, g, out=dl), R * V * 6)
    lb = logits.to(bf)
    bw_case("logps fwd bf16 logits", lambda: ops.logps_fwd(lb, tgt, 8), R * V * 2)
    del logits, dl, lb
    n = 1 << 30
    p = torch.zeros(n, dtype=bf, device=dev)
    gr = torch.full((n,), 1e-3, dtype=bf, device=dev)
    ma, m1, m2 = (torch.zeros(n, device=dev) for _ in range(3))
    ss, ws = torch.zeros(1, device=dev), torch.zeros(1024, device=dev)
    bw_case("grad sumsq (1 Gi)", lambda: ops.sumsq(gr, ss, ws), n * 2)
    bw_case("adamw (1 Gi params)", lambda: ops.adamw_(p, gr, ma, m1, m2, 1e-6, 0.9, 0.98, 1e-6, 0.0, 1, grad_sumsq=ss, max_grad_norm=1.0), n * 28)
    print("KERNEL_BENCH_JSON " + json.dumps(res))


if __name__ == "__main__":
    main()
