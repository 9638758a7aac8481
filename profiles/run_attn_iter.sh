#!/bin/bash
# attention kernel iteration: correctness (kernel + shared-prefix + engine parity tests), then timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_share_prefix.py tests/test_gpu_engine.py -m gpu -q --maxfail=10 -x -k "attention or ctx or shared or parity or config" > gpurun_out/attn_iter_tests.log 2>&1
echo "pytest rc=$?"; grep -n "^FAILED\|^ERROR\|passed\|failed\|watchdog\|Error" gpurun_out/attn_iter_tests.log | tail -15
python tests/attn_probe2.py time 2>&1 | tee gpurun_out/attn_probe_time.log
python tests/attn_probe3.py 2>&1 | tee -a gpurun_out/attn_probe_time.log
