#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_xc2.py tests/test_gpu_zz_unvalidated.py -m gpu -q -s -k "xc2" > gpurun_out/xc2_tests.log 2>&1
echo "pytest rc=$?"; grep -n "^FAILED\|passed\|failed\|g13.*logps \|worst" gpurun_out/xc2_tests.log | tail -20
bash profiles/run_multi.sh 1 xc2_7b 2>&1 | grep -v "^\*\*\*\|OMP_NUM"
