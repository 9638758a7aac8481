"""Bring-up probe for the tcgen05 GEMM (not a pytest file): each variant runs in its own subprocess so a
trap / watchdog in one does not poison the others.  Prints error statistics that localise layout bugs."""
import subprocess
import sys

CASES = [
    # name, M, N, K, a_kmajor, b_kmajor, extra
    ("tn_1tile", 128, 256, 64, 1, 1, ""),
    ("tn_k256", 128, 256, 256, 1, 1, ""),
    ("tn_multi", 512, 1024, 512, 1, 1, ""),
    ("tn_ragged", 300, 320, 200, 1, 1, ""),
    ("tn_n128", 256, 128, 192, 1, 1, ""),
    ("nn_b_mn", 256, 512, 256, 1, 0, ""),
    ("tt_a_mn", 256, 512, 256, 0, 1, ""),
    ("nt_both_mn", 256, 512, 256, 0, 0, ""),
    ("tn_big", 4096, 4096, 4096, 1, 1, "time"),
    ("tn_epi", 384, 768, 320, 1, 1, "epi"),
    ("tn_odd", 12792, 4096, 4096, 1, 1, "time"),
    ("wgrad", 4096, 11008, 12792, 0, 0, "time"),
    ("dgrad", 12792, 4096, 22016, 1, 0, "time"),
    ("gate_up", 12792, 22016, 4096, 1, 1, "time"),
]


def run_case(name, M, N, K, ak, bk, extra):
    import torch
    import vlrlhf_b200  # noqa: F401
    from vlrlhf_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    a = (torch.randn((M, K) if ak else (K, M), device=dev) * 0.5).to(torch.bfloat16)
    b = (torch.randn((N, K) if bk else (K, N), device=dev) * 0.5).to(torch.bfloat16)
    A = a.float() if ak else a.float().t()
    B = b.float() if bk else b.float().t()
    torch.backends.cuda.matmul.allow_tf32 = False
    want = A @ B.t()
    kw = {}
    if extra == "epi":
        bias = torch.randn(N, device=dev).to(torch.bfloat16)
        res = torch.randn(M, N, device=dev).to(torch.bfloat16)
        kw = dict(bias=bias, act=ops.ACT_QUICK_GELU, residual=res)
        x = want + bias.float()
        want = x * torch.sigmoid(1.702 * x) + res.float()
    got = ops.gemm(a, b, a_kmajor=bool(ak), b_kmajor=bool(bk), out_dtype=torch.float32, **kw)
    torch.cuda.synchronize()
    err = (got - want).abs()
    tol = 1e-2 + 2e-3 * want.abs()
    bad = err > tol
    print(f"[{name}] M={M} N={N} K={K} ak={ak} bk={bk} max_err={err.max().item():.4g} "
          f"bad={bad.sum().item()}/{bad.numel()} ref_absmax={want.abs().max().item():.3g}", flush=True)
    if bad.any():
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print(f"   bad rows: n={rows.numel()} first={rows[:12].tolist()} last={rows[-4:].tolist()}")
        print(f"   bad cols: n={cols.numel()} first={cols[:12].tolist()} last={cols[-4:].tolist()}")
        print("   got[0,:8] ", got[0, :8].tolist())
        print("   want[0,:8]", want[0, :8].tolist())
        # is the result a permutation of K-slices / rows?  report simple diagnostics
        print("   mean|got|", got.abs().mean().item(), "mean|want|", want.abs().mean().item())
    if extra == "time":
        out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        for _ in range(3):
            ops.gemm(a, b, a_kmajor=bool(ak), b_kmajor=bool(bk), out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.gemm(a, b, a_kmajor=bool(ak), b_kmajor=bool(bk), out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"   time {ms:.3f} ms  -> {2 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
        if ak and bk:
            wt = b.t()
            ref = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
            for _ in range(3):
                torch.matmul(a, wt, out=ref)
            e0.record()
            for _ in range(10):
                torch.matmul(a, wt, out=ref)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"   cuBLAS {ms:.3f} ms -> {2 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
    return 0 if not bad.any() else 1


if __name__ == "__main__":
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if len(sys.argv) > 1 and sys.argv[1] not in ("1cta", "2cta"):
        c = [c for c in CASES if c[0] == sys.argv[1]][0]
        sys.exit(run_case(*c))
    mode = sys.argv[1] if len(sys.argv) > 1 else "1cta"
    env = dict(os.environ, VLB200_GEMM_2CTA="1" if mode == "2cta" else "0")
    print("gemm probe mode:", mode)
    fails = 0
    for c in CASES:
        try:
            r = subprocess.run([sys.executable, __file__, c[0]], timeout=120, capture_output=True, text=True, env=env)
            print(r.stdout, end="")
            if r.returncode != 0:
                fails += 1
                print(f"[{c[0]}] exit={r.returncode} stderr tail: {r.stderr[-600:]}")
        except subprocess.TimeoutExpired:
            fails += 1
            print(f"[{c[0]}] TIMEOUT")
    print("probe fails:", fails)
