#!/bin/bash
# r2t: the whole -m gpu suite + smoke + the default bench with the new attention kernels
bash profiles/run_tests_only.sh
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r2t_bench_7b_1gpu.json 2> gpurun_out/r2t_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2t_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2t_bench_7b_1gpu.json").read().strip().splitlines() if l.startswith("{")][-1])
print("ms/step", round(d["ms_per_step"], 1), "pairs/s", round(d["value"], 3), "e2e", round(d["e2e"]["value"], 3), "plugin", d.get("e2e_plugin", {}).get("value"))
print("attention", [(round(r["ms"], 3), round(r["frac"], 3)) for r in d["roofline_attention"]], "clocks", d["clocks"]["sm_mhz"])
print("padded", d.get("padded_layout"))
PY
