#!/bin/bash
# round 2, GPU call B: shared-prefix kernels + engine tests, then the bench with and without --share-prefix on the same box
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_share_prefix.py tests/test_gpu_kernels.py -m gpu -q --maxfail=20 -s > gpurun_out/r2b_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_tests.log
grep -n "^FAILED\|^ERROR\|passed\|failed\|rel-l2\|watchdog" gpurun_out/r2b_tests.log | tail -60
for flag in "" "--share-prefix" "--pack"; do
  tag=${flag#--}; tag=${tag:-padded}
  timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-library-baseline $flag > gpurun_out/r2b_bench_${tag}.json 2> gpurun_out/r2b_bench_${tag}.err
  echo "bench $tag rc=$?"; tail -c 400 gpurun_out/r2b_bench_${tag}.err
done
python - <<'PY'
import json
for t in ("padded", "share-prefix", "pack"):
    try:
        d = json.loads(open(f"gpurun_out/r2b_bench_{t}.json").read().strip().splitlines()[-1])
        print(t, "ms/step", round(d["ms_per_step"], 1), "pairs/s", round(d["value"], 3), "e2e", round(d["e2e"]["value"], 3),
              "plugin", (d.get("e2e_plugin") or {}).get("value"), "rows", d["config"]["rows_per_step"],
              "loss", d["e2e"]["last_metrics"].get("loss"), "logps/chosen", d["e2e"]["last_metrics"].get("logps/chosen"), d["clocks"]["sm_mhz"])
    except Exception as e:
        print(t, "unreadable:", e)
PY
