#!/bin/bash
# r2am: backward: delta kernel with 32-bit index arithmetic, dQ pass copies the next item's stationary tiles ahead
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_share_prefix.py tests/test_gpu_engine.py tests/test_gpu_qwen.py -m gpu -q > gpurun_out/r2am_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r2am_tests.log)"
grep -n "^FAILED\|^E  .*rel l2\|watchdog\|Error" gpurun_out/r2am_tests.log | head -8
{
timeout 300 python tests/attn_probe2.py time 2>&1 | grep "^\["
echo "== phases: backward"; VLB200_ATTN_BWD_DBG=8 timeout 300 python tests/attn_phase_probe.py
} 2>&1 | grep -v Warning | tee gpurun_out/r2am_attn.log
