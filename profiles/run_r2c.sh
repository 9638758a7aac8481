#!/bin/bash
# round 2, GPU call C: the default bench line (share_prefix default, padded_layout beside it, plugin, library, cpu arms),
# the launch list of one timed step under ncu (shares), and the whole -m gpu suite for the parity table.
mkdir -p gpurun_out
rm -f gpurun_out/parity_r2.jsonl
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -s > gpurun_out/tests_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/tests_gpu.log
grep -n "^FAILED\|^ERROR\|passed\|failed" gpurun_out/tests_gpu.log | tail -20
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r2c_bench.err
VLB_NVTX=1 timeout 900 ncu --nvtx --nvtx-include "vlb_step/" --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2c_launches.csv python bench.py --steps 1 --warmup 3 --skip-e2e --no-cpu-baseline --no-library-baseline \
    > gpurun_out/r2c_ncu_bench.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r2c_launches.csv
python profiles/summarize_launches.py gpurun_out/r2c_launches.csv > gpurun_out/r2c_launches_summary.md 2>/dev/null; head -30 gpurun_out/r2c_launches_summary.md
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "plugin", d["e2e_plugin"]["value"],
      "padded", d["padded_layout"], "lib", d.get("gpu_library_baseline"), "cpu", d.get("cpu_baseline", {}).get("value"),
      "util", d["config"]["step_tensor_util_of_sustained_peak"], "roof", d["roofline"]["frac"],
      [x["frac"] for x in d["roofline_gemm_longk"]], [x["frac"] for x in d["roofline_attention"]], d["clocks"])
PY
