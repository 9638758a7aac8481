"""gpurun_out/parity_r2.jsonl (written by tests/parity_log.py during `pytest -m gpu` on a B200) -> profiles/parity_r2.md."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity_r2.jsonl")
rows = {}
for line in open(src):
    r = json.loads(line)
    rows[(r["case"], r["quantity"])] = r   # the last run wins
out = ["# Parity table (round 2) -- CUDA path vs golden fixtures minted from the reference / the fp32 oracle", "",
       "Written by `tests/parity_log.py` during `pytest -m gpu` on a B200; one row per comparison.  `bound` is what the test asserts.",
       "", "| case | quantity | n | max abs err | max rel err | |want| max | bound |", "|---|---|---|---|---|---|---|"]
for (case, q), r in sorted(rows.items()):
    b = []
    if r["bound_rel"] is not None:
        b.append(f"rel {r['bound_rel']:g}")
    if r["bound_abs"] is not None:
        b.append(f"abs {r['bound_abs']:g}")
    out.append(f"| {case} | {q} | {r['n']} | {r['max_abs_err']:.3e} | {r['max_rel_err']:.3e} | {r['want_absmax']:.4g} | {', '.join(b)} |")
out += ["", "## Noise floor of the 7B-shape fixtures (why their log-prob bound is `tests/parity_log.py: RTOL_7B` = 2e-3)", "",
        "Maximum relative error of the per-sequence log-probs against the reference's fp32 CPU run, for four builds of the attention",
        "forward that differ only in arithmetic-neutral details (`profiles/r2u_experiment.log`, `gpurun_out` logs of the same day):", "",
        "| fixture | r2c build (P through smem, 2 row-sum accumulators) | P in TMEM, 4 accumulators | + denominator from the rounded bf16 P | shipped (`attn_fwd_tc2_kernel`) |",
        "|---|---|---|---|---|",
        "| g5 config 1 (LLaVA-1.5-7B, text 128) policy / reference | 3.2e-4 / 5.3e-4 | 1.01e-3 / - | 8.0e-4 / 5.1e-4 | see table above |",
        "| g12 config 3 (Qwen-VL 7B) policy / reference | 5.6e-4 / 3.1e-4 | 2.1e-4 / 1.9e-4 | 6.2e-4 / 2.0e-4 | see table above |",
        "| g13 config 5 (XC2 7B) policy / reference | 1.2e-3 (1.1e-3 with the combined adapter operand) / 4.3e-4 | 9.0e-4 / 6.0e-4 | 1.6e-3 / - | see table above |",
        "| g14 config 2 full length policy / reference | 5.8e-5 / 2.3e-4 | 1.2e-4 / 1.2e-4 | 1.2e-4 / 2.5e-4 | see table above |", "",
        "The same fixture moves by up to 5x between builds whose arithmetic differs at the last fp32 bit: a bf16 pipeline through 32",
        "random-weight layers sits at (1 +- 1)e-3 of an fp32 run at this size; the reference's own bf16 execution is at 4.3e-3 on g5",
        "(row `REFERENCE bf16 vs its fp32`).  The tiny / small fixtures (same kernels, fewer and narrower layers) are held to 1e-3 and sit at 0.5-3e-4."]
open(os.path.join(ROOT, "profiles", "parity_r2.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
