#!/bin/bash
# r2l: backward with producer-staged statistics (prefetched), MODE 1 stationary tiles in TMEM, staged coalesced epilogue stores
mkdir -p gpurun_out
T="tests/test_gpu_kernels.py tests/test_gpu_share_prefix.py"
K="attention or share_prefix_rows_kernel or shared_step_equals"
timeout 600 python -m pytest $T -m gpu -q -x -k "$K" > gpurun_out/r2l_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r2l_tests.log)"
grep -n "^FAILED\|^E  .*rel l2\|watchdog\|Error" gpurun_out/r2l_tests.log | head -8
{
VLB200_ATTN_FWD_VARIANT=1 timeout 300 python tests/attn_probe2.py time 2>&1 | grep "^\["
echo "== phases: backward"; VLB200_ATTN_BWD_DBG=8 timeout 300 python tests/attn_phase_probe.py
} 2>&1 | grep -v Warning | tee gpurun_out/r2l_attn.log
