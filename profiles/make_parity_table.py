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
open(os.path.join(ROOT, "profiles", "parity_r2.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
