#!/bin/bash
# r2s: forward FMNMX3 max chains + polynomial-exp variant; backward explicit ld.shared statistics
mkdir -p gpurun_out
T="tests/test_gpu_kernels.py tests/test_gpu_share_prefix.py"
K="attention or share_prefix_rows_kernel or shared_step_equals"
for v in 1 3; do
VLB200_ATTN_FWD_VARIANT=$v timeout 600 python -m pytest $T -m gpu -q -x -k "$K" > gpurun_out/r2s_tests_$v.log 2>&1
echo "tests[fwd $v] rc=$? $(tail -1 gpurun_out/r2s_tests_$v.log)"
grep -n "^FAILED\|^E  .*rel l2\|watchdog\|Error" gpurun_out/r2s_tests_$v.log | head -8
done
{
for v in 1 3; do echo "== fwd variant $v"; VLB200_ATTN_FWD_VARIANT=$v timeout 300 python tests/attn_probe2.py time 2>&1 | grep "^\["; done
echo "== phases: forward variant 1"; VLB200_ATTN_FWD_VARIANT=81 timeout 300 python tests/attn_phase_probe.py
echo "== phases: forward variant 3"; VLB200_ATTN_FWD_VARIANT=83 timeout 300 python tests/attn_phase_probe.py
echo "== phases: backward"; VLB200_ATTN_BWD_DBG=8 timeout 300 python tests/attn_phase_probe.py
} 2>&1 | grep -v Warning | tee gpurun_out/r2s_attn.log
