#!/bin/bash
# round 2, GPU call D: whole -m gpu suite (incl. g12/g13 7B-shape parity, fused SwiGLU backward), default bench, fused vs unfused A/B
mkdir -p gpurun_out
rm -f gpurun_out/parity_r2.jsonl
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -s > gpurun_out/tests_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/tests_gpu.log
grep -n "^FAILED\|^ERROR\|passed\|failed" gpurun_out/tests_gpu.log | tail -20
for f in 1 0; do
  VLB200_FUSE_SWIGLU_BWD=$f timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline --skip-plugin > gpurun_out/r2d_bench_fuse$f.json 2> gpurun_out/r2d_bench_fuse$f.err
  echo "bench fuse=$f rc=$?"; tail -c 300 gpurun_out/r2d_bench_fuse$f.err
done
python - <<'PY'
import json
for f in (1, 0):
    d = json.loads(open(f"gpurun_out/r2d_bench_fuse{f}.json").read().strip().splitlines()[-1])
    print("fuse", f, "ms/step", round(d["ms_per_step"], 1), "pairs/s", round(d["value"], 3), "e2e", round(d["e2e"]["value"], 3),
          "padded", d["padded_layout"]["ms_per_step"], "launches", d["gpu_launches"], d["clocks"]["sm_mhz"])
PY
