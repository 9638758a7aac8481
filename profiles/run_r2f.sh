#!/bin/bash
# r2f: XC2 combined [lora|Plora] operand + shared-prefix rows on the Qwen-VL / XC2 engines: tests, then bench with and without sharing
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_xc2.py tests/test_gpu_qwen.py tests/test_gpu_zz_unvalidated.py -m gpu -q -s > gpurun_out/r2f_tests.log 2>&1
echo "pytest rc=$?"; grep -n "^FAILED\|^ERROR\|passed\|failed\|worst\|rel l2\|Error" gpurun_out/r2f_tests.log | tail -30
cp gpurun_out/parity_r2.jsonl gpurun_out/r2f_parity.jsonl 2>/dev/null
grep -h "\[parity\] g1[23].*logps " gpurun_out/r2f_tests.log | tail -12
for m in qwen7b xc2_7b; do
  for sp in "" "--no-share-prefix"; do
    tag=${m}$( [ -z "$sp" ] && echo _shared || echo _padded )
    timeout 600 python bench.py --gpus 1 --steps 6 --warmup 3 --model $m $sp --no-cpu-baseline --no-library-baseline > gpurun_out/r2f_bench_${tag}.json 2> gpurun_out/r2f_bench_${tag}.err
    echo "$tag rc=$?"; tail -3 gpurun_out/r2f_bench_${tag}.err
    python - "$tag" <<'PY'
import json, sys
t = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r2f_bench_{t}.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("  ->", t, "ms/step", round(d["ms_per_step"], 1), "pairs/s", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "launches", d["gpu_launches"], "loss", d["e2e"]["last_metrics"].get("loss"))
except Exception as e:
    print("  -> unreadable", e)
PY
  done
done
