#!/bin/bash
# r2ai: packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2) in the forward softmax and the backward elementwise loops
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_share_prefix.py tests/test_gpu_engine.py -m gpu -q -k "attention or share_prefix or config1 or golden or shared" > gpurun_out/r2ai_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r2ai_tests.log)"
grep -n "^FAILED\|^E  .*rel l2\|watchdog\|Error" gpurun_out/r2ai_tests.log | head -8
{
VLB200_ATTN_FWD_VARIANT=5 timeout 300 python tests/attn_probe2.py time 2>&1 | grep "^\["
echo "== phases: forward variant 5"; VLB200_ATTN_FWD_VARIANT=85 timeout 300 python tests/attn_phase_probe.py
echo "== phases: backward"; VLB200_ATTN_BWD_DBG=8 timeout 300 python tests/attn_phase_probe.py
} 2>&1 | grep -v Warning | tee gpurun_out/r2ai_attn.log
