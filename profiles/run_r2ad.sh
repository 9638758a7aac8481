#!/bin/bash
# r2ad: explicit ld.shared / st.shared in the attention epilogues and exchanges
mkdir -p gpurun_out
T="tests/test_gpu_kernels.py tests/test_gpu_share_prefix.py tests/test_gpu_engine.py"
K="attention or share_prefix_rows_kernel or shared_step_equals or config1 or golden"
for v in 4 5; do
VLB200_ATTN_FWD_VARIANT=$v timeout 900 python -m pytest $T -m gpu -q -x -k "$K" > gpurun_out/r2ad_tests_$v.log 2>&1
echo "tests[fwd $v] rc=$? $(tail -1 gpurun_out/r2ad_tests_$v.log)"
grep -n "^FAILED\|^E  .*rel l2\|watchdog\|Error" gpurun_out/r2ad_tests_$v.log | head -8
done
{
for v in 4 5; do echo "== fwd variant $v"; VLB200_ATTN_FWD_VARIANT=$v timeout 300 python tests/attn_probe2.py time 2>&1 | grep "^\["; done
echo "== phases: forward variant 5"; VLB200_ATTN_FWD_VARIANT=85 timeout 300 python tests/attn_phase_probe.py
echo "== phases: backward"; VLB200_ATTN_BWD_DBG=8 timeout 300 python tests/attn_phase_probe.py
} 2>&1 | grep -v Warning | tee gpurun_out/r2ad_attn.log
