#!/bin/bash
# r2af: epilogue reorder (forward), kernel tests over every forward generation, measured loss bounds
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_share_prefix.py tests/test_gpu_engine.py tests/test_gpu_lora.py -m gpu -q > gpurun_out/r2af_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r2af_tests.log)"
grep -n "^FAILED\|^E  .*rel l2\|watchdog\|Error" gpurun_out/r2af_tests.log | head -8
{
for v in 5; do echo "== fwd variant $v"; VLB200_ATTN_FWD_VARIANT=$v timeout 300 python tests/attn_probe2.py time 2>&1 | grep "^\["; done
echo "== phases: forward variant 5"; VLB200_ATTN_FWD_VARIANT=85 timeout 300 python tests/attn_phase_probe.py
} 2>&1 | grep -v Warning | tee gpurun_out/r2af_attn.log
