#!/bin/bash
# bench.py on N GPUs of one box for the headline config and configs 3/4/5 (side measurements): usage run_multi.sh N [models...]
N=${1:-2}; shift
MODELS=${@:-"7b qwen7b xc2_7b next7b_lora 7b_lora"}
mkdir -p gpurun_out
port=29511
for m in $MODELS; do
  port=$((port+1))
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --steps 6 --warmup 3 --model $m --no-cpu-baseline --no-library-baseline > gpurun_out/r2_bench_${m}_${N}gpu.json 2> gpurun_out/r2_bench_${m}_${N}gpu.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 6 --warmup 3 --model $m > gpurun_out/r2_bench_${m}_${N}gpu.json 2> gpurun_out/r2_bench_${m}_${N}gpu.err
  fi
  echo "$m N=$N rc=$?"; tail -c 300 gpurun_out/r2_bench_${m}_${N}gpu.err | tail -3
  python - "$m" "$N" <<'PY'
import json, sys
m, n = sys.argv[1], sys.argv[2]
try:
    d = json.loads([l for l in open(f"gpurun_out/r2_bench_{m}_{n}gpu.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("  ->", m, "N", d["n_gpus"], "ms/step", round(d["ms_per_step"], 1), "pairs/s", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2) if d["e2e"]["value"] else None,
          "util", round(d["config"].get("step_tensor_util_of_sustained_peak", 0), 3), "loss", d["e2e"]["last_metrics"].get("loss"))
except Exception as e:
    print("  -> unreadable", e)
PY
done
