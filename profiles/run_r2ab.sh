#!/bin/bash
# r2ab: third-generation forward with every fourth exponential as a polynomial (variant 45), finer epilogue profile
mkdir -p gpurun_out
T="tests/test_gpu_kernels.py tests/test_gpu_share_prefix.py tests/test_gpu_engine.py"
K="attention or share_prefix_rows_kernel or shared_step_equals or config1 or golden"
VLB200_ATTN_FWD_VARIANT=45 timeout 900 python -m pytest $T -m gpu -q -x -k "$K" > gpurun_out/r2ab_tests.log 2>&1
echo "tests[fwd 45] rc=$? $(tail -1 gpurun_out/r2ab_tests.log)"
grep -n "^FAILED\|^E  .*rel l2\|watchdog\|Error" gpurun_out/r2ab_tests.log | head -8
{
for v in 5 45; do echo "== fwd variant $v"; VLB200_ATTN_FWD_VARIANT=$v timeout 300 python tests/attn_probe2.py time 2>&1 | grep "^\["; done
echo "== phases: forward variant 5"; VLB200_ATTN_FWD_VARIANT=85 timeout 300 python tests/attn_phase_probe.py
echo "== phases: forward variant 45"; VLB200_ATTN_FWD_VARIANT=125 timeout 300 python tests/attn_phase_probe.py
} 2>&1 | grep -v Warning | tee gpurun_out/r2ab_attn.log
