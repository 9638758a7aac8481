#!/bin/bash
# r2h: which side bounds the attention forward?  DBG variants: 1x = no MMAs (softmax side alone), 2x = no exponentials (tensor side alone)
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv,noheader -lms 200 > gpurun_out/r2h_clocks.log 2>&1 &
SMI=$!
for v in 2 40 42 30 32 70 72; do
  echo "== fwd variant $v"
  VLB200_ATTN_FWD_VARIANT=$v timeout 300 python tests/attn_probe2.py time 2>&1 | grep "^\["
done | tee gpurun_out/r2h_attn_timing.log
kill $SMI
sort gpurun_out/r2h_clocks.log | uniq -c | sort -rn | head -5
