#!/bin/bash
# round 2, GPU call A: full -m gpu suite (with the parity table), the default bench (e2e, e2e_plugin, library baseline),
# and the long-K raster A/B (VLB200_RASTER_POLICY=model vs default) on the same box.
mkdir -p gpurun_out
rm -f gpurun_out/parity_r2.jsonl
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 -s > gpurun_out/r2a_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_tests.log
tail -25 gpurun_out/r2a_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2a_bench.err
VLB200_RASTER_POLICY=model timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline --skip-plugin \
    > gpurun_out/r2a_bench_raster_model.json 2> gpurun_out/r2a_bench_raster_model.err
echo "raster-model bench rc=$?"
python - <<'PY'
import json
for f in ("r2a_bench.json", "r2a_bench_raster_model.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"], 1), "pairs/s", round(d["value"], 3), "e2e", d["e2e"]["value"],
              "plugin", (d.get("e2e_plugin") or {}).get("value"), "lib", d.get("gpu_library_baseline"),
              "roofline", round(d["roofline"]["frac"], 3), [round(x["frac"], 3) for x in d.get("roofline_gemm_longk", [])],
              [round(x["frac"], 3) for x in d["roofline_attention"]], d["clocks"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
