"""Bring-up probe for the tcgen05 attention forward (not a pytest file): one subprocess per case."""
import math
import os
import subprocess
import sys

CASES = [
    # name, B, S, H, KV, dh, causal, lens, time
    ("one_tile", 1, 128, 1, 1, 128, 0, None, 0),
    ("two_tiles_noncausal", 1, 256, 1, 1, 128, 0, None, 0),
    ("causal_3tiles", 1, 384, 2, 2, 128, 1, None, 0),
    ("ragged_causal", 2, 200, 2, 2, 128, 1, [200, 131], 0),
    ("gqa", 3, 330, 4, 2, 128, 1, [330, 64, 1], 0),
    ("vit", 2, 577, 4, 4, 64, 0, None, 0),
    ("bigvals", 1, 512, 2, 2, 128, 1, None, 0),
    ("config2", 8, 1599, 32, 32, 128, 1, [1599, 1400, 1599, 1500, 1599, 1300, 1450, 1599], 1),
    ("vit_full", 4, 577, 16, 16, 64, 0, None, 1),
]


def run_case(name, B, S, H, KV, dh, causal, lens, do_time):
    import torch
    import vlrlhf_b200  # noqa: F401
    from vlrlhf_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    ld = (H + 2 * KV) * dh
    mult = 4.0 if name == "bigvals" else 1.0
    qkv = (torch.randn(B * S, ld, device=dev) * mult).to(torch.bfloat16)
    sl = torch.tensor(lens, device=dev, dtype=torch.int32) if lens is not None else None
    sc = 1 / math.sqrt(dh)
    q, k, v = qkv[:, :H * dh], qkv[:, H * dh:(H + KV) * dh], qkv[:, (H + KV) * dh:]
    out = torch.zeros(B * S, H * dh, dtype=torch.bfloat16, device=dev)
    lse = torch.zeros(B, H, S, dtype=torch.float32, device=dev)
    ops.attn_fwd_tc(q, k, v, out, lse, sl, B, S, H, KV, dh, bool(causal), sc)
    torch.cuda.synchronize()
    ref = torch.zeros_like(out)
    ref_lse = torch.zeros_like(lse)
    ops.attn_fwd(q, k, v, ref, ref_lse, sl, B, S, H, KV, dh, bool(causal), sc)  # mma.sync kernel (validated vs torch)
    torch.cuda.synchronize()
    valid = torch.ones(B, S, dtype=torch.bool, device=dev)
    if sl is not None:
        valid = torch.arange(S, device=dev)[None] < sl[:, None].long()
    o1 = out.view(B, S, H * dh).float()[valid]
    o2 = ref.view(B, S, H * dh).float()[valid]
    l1 = lse.permute(0, 2, 1)[valid]
    l2 = ref_lse.permute(0, 2, 1)[valid]
    eo = (o1 - o2).abs().max().item()
    el = (l1 - l2).abs().max().item()
    fin = bool(torch.isfinite(out.float()).all())
    bad = eo > 3e-2 * max(1.0, o2.abs().max().item()) or el > 2e-3 * max(1.0, l2.abs().max().item()) or not fin
    print(f"[{name}] B={B} S={S} H={H} KV={KV} dh={dh} causal={causal} max|dO|={eo:.4g} max|dLSE|={el:.4g} "
          f"finite={fin} refmax={o2.abs().max().item():.3g} {'BAD' if bad else 'ok'}", flush=True)
    if bad:
        d = (out.view(B, S, H, dh).float() - ref.view(B, S, H, dh).float()).abs()
        d = d * valid[:, :, None, None]
        per_row = d.amax(dim=(2, 3))
        rows = (per_row > 3e-2).nonzero()
        print("   first bad (b, row):", rows[:10].tolist(), " n_bad_rows:", rows.shape[0])
        per_col = d.amax(dim=(0, 1, 2))
        print("   bad dh cols:", (per_col > 3e-2).nonzero().flatten()[:16].tolist())
        print("   got ", out.view(B, S, H, dh)[0, 0, 0, :8].tolist())
        print("   want", ref.view(B, S, H, dh)[0, 0, 0, :8].tolist())
        print("   lse got/want", lse[0, 0, :4].tolist(), ref_lse[0, 0, :4].tolist())
    # ---- backward: tcgen05 vs the validated mma.sync kernels (same forward outputs as input)
    dout = (torch.randn(B * S, H * dh, device=dev)).to(torch.bfloat16) * valid.view(-1, 1)
    g_ref = torch.zeros_like(qkv)
    g_tc = torch.full_like(qkv, float("nan"))
    delta = torch.zeros(B, H, S, dtype=torch.float32, device=dev)
    hq, hk = H * dh, (H + KV) * dh
    ops.attn_bwd(q, k, v, ref, dout, ref_lse, delta, g_ref[:, :hq], g_ref[:, hq:hk], g_ref[:, hk:], sl, B, S, H, KV, dh, bool(causal), sc)
    torch.cuda.synchronize()
    bad_b = False
    try:
        ops.attn_bwd_tc(q, k, v, ref, dout, ref_lse, delta, g_tc[:, :hq], g_tc[:, hq:hk], g_tc[:, hk:], sl, B, S, H, KV, dh, bool(causal), sc)
        torch.cuda.synchronize()
        for nm, a0, a1 in (("dq", 0, hq), ("dk", hq, hk), ("dv", hk, ld)):
            x, y = g_tc[:, a0:a1].float(), g_ref[:, a0:a1].float()
            fin2 = bool(torch.isfinite(x).all())
            rel = ((x - y).norm() / y.norm().clamp(min=1e-9)).item() if fin2 else float("nan")
            flag = (not fin2) or rel > 2e-2
            bad_b |= flag
            print(f"   bwd {nm}: rel_l2={rel:.4g} finite={fin2} {'BAD' if flag else 'ok'}", flush=True)
            if flag and fin2:
                d = (x - y).abs().view(B, S, -1).amax(-1)
                rows = (d > 0.05 * y.abs().max()).nonzero()
                print("      first bad (b,row):", rows[:8].tolist(), "n:", rows.shape[0])
    except Exception as e:  # noqa: BLE001
        bad_b = True
        print("   bwd tc raised:", repr(e)[:300])
    bad = bad or bad_b
    if do_time:
        def t(fn):
            for _ in range(2):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / 10
        ls = lens if lens is not None else [S] * B
        fl = sum(4.0 * l * l * dh * H * (0.5 if causal else 1.0) for l in ls)
        ms = t(lambda: ops.attn_fwd_tc(q, k, v, out, lse, sl, B, S, H, KV, dh, bool(causal), sc))
        ms0 = t(lambda: ops.attn_fwd(q, k, v, ref, ref_lse, sl, B, S, H, KV, dh, bool(causal), sc))
        print(f"   fwd tcgen05 {ms:.3f} ms = {fl / ms / 1e9:.1f} TFLOP/s   | mma.sync {ms0:.3f} ms = {fl / ms0 / 1e9:.1f} TFLOP/s", flush=True)
        if not bad_b:
            msb = t(lambda: ops.attn_bwd_tc(q, k, v, ref, dout, ref_lse, delta, g_tc[:, :hq], g_tc[:, hq:hk], g_tc[:, hk:], sl, B, S, H, KV, dh, bool(causal), sc))
            msb0 = t(lambda: ops.attn_bwd(q, k, v, ref, dout, ref_lse, delta, g_ref[:, :hq], g_ref[:, hq:hk], g_ref[:, hk:], sl, B, S, H, KV, dh, bool(causal), sc))
            print(f"   bwd tcgen05 {msb:.3f} ms = {2.5 * fl / msb / 1e9:.1f} TFLOP/s (2.5x fwd flops) | mma.sync {msb0:.3f} ms = {2.5 * fl / msb0 / 1e9:.1f} TFLOP/s", flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if len(sys.argv) > 1:
        c = [c for c in CASES if c[0] == sys.argv[1]][0]
        sys.exit(run_case(*c))
    fails = 0
    for c in CASES:
        try:
            r = subprocess.run([sys.executable, __file__, c[0]], timeout=120, capture_output=True, text=True)
            print(r.stdout, end="")
            if r.returncode != 0:
                fails += 1
                print(f"[{c[0]}] exit={r.returncode} stderr tail: {r.stderr[-700:]}")
        except subprocess.TimeoutExpired:
            fails += 1
            print(f"[{c[0]}] TIMEOUT")
    print("attn_tc probe fails:", fails)
