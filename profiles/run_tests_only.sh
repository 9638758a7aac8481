#!/bin/bash
# the -m gpu suite alone (parity table lines into gpurun_out/parity_r2.jsonl)
mkdir -p gpurun_out
rm -f gpurun_out/parity_r2.jsonl
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -s "$@" > gpurun_out/tests_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/tests_gpu.log
grep -n "^FAILED\|^ERROR\|passed\|failed" gpurun_out/tests_gpu.log | tail -45
