#!/bin/bash
# r2ag: final verification of round 2: whole -m gpu suite, smoke, the default bench line (all arms), launch list of one step
rm -f gpurun_out/parity_r2.jsonl
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -s > gpurun_out/tests_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/tests_gpu.log
grep -n "^FAILED\|^ERROR\|passed\|failed" gpurun_out/tests_gpu.log | tail -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2ag_bench_7b_1gpu.json 2> gpurun_out/r2ag_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2ag_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2ag_bench_7b_1gpu.json").read().strip().splitlines() if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "plugin", d["e2e_plugin"]["value"],
      "padded", d["padded_layout"], "lib", d.get("gpu_library_baseline"), "cpu", d.get("cpu_baseline", {}).get("value"),
      "util", d["config"]["step_tensor_util_of_sustained_peak"], "roof", d["roofline"]["frac"],
      [x["frac"] for x in d["roofline_gemm_longk"]], [(x["ms"], x["frac"], x.get("sm_mhz")) for x in d["roofline_attention"]], d["clocks"])
PY
VLB_NVTX=1 timeout 900 ncu --nvtx --nvtx-include "vlb_step/" --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2ag_launches.csv python bench.py --steps 1 --warmup 3 --skip-e2e --skip-plugin --no-cpu-baseline --no-library-baseline \
    > gpurun_out/r2ag_ncu_bench.log 2>&1
echo "ncu launches rc=$?"; wc -l gpurun_out/r2ag_launches.csv
python profiles/summarize_launches.py gpurun_out/r2ag_launches.csv > gpurun_out/r2ag_launches_summary.md 2>/dev/null; head -24 gpurun_out/r2ag_launches_summary.md
