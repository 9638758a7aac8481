#!/bin/bash
# r2u: does a softmax denominator summed from the rounded bf16 P (weights sum to exactly 1) tighten the 7B-shape parity?
mkdir -p gpurun_out
for v in 1 3; do
  VLB200_ATTN_FWD_VARIANT=$v timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_qwen.py tests/test_gpu_xc2.py tests/test_gpu_share_prefix.py -m gpu -q -s -k "config or 7b" > gpurun_out/r2u_tests_$v.log 2>&1
  echo "== variant $v: $(tail -1 gpurun_out/r2u_tests_$v.log)"
  grep "\[parity\].* policy_logps \|\[parity\].* ref_logps \|config4" gpurun_out/r2u_tests_$v.log | grep -v REFERENCE | awk '{print $2,$3,$4,$5,$6,$7,$8}' | cut -c1-90
done
VLB200_ATTN_FWD_VARIANT=3 timeout 300 python tests/attn_probe2.py time 2>&1 | grep "^\["
