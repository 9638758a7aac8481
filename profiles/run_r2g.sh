#!/bin/bash
# r2g: attention with operands in tensor memory (TS-mode MMAs): correctness per variant, then timing of every combination
mkdir -p gpurun_out
T="tests/test_gpu_kernels.py tests/test_gpu_share_prefix.py"
K="attention or share_prefix_rows_kernel or shared_step_equals"
run_tests() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m pytest $T -m gpu -q -x -k "$K" > gpurun_out/r2g_tests_$name.log 2>&1
  echo "tests[$name] rc=$? $(tail -1 gpurun_out/r2g_tests_$name.log)"
  grep -n "^FAILED\|^E  .*rel l2\|watchdog\|Error" gpurun_out/r2g_tests_$name.log | head -8
}
run_tests f1b0 VLB200_ATTN_FWD_VARIANT=1
run_tests f2b0 VLB200_ATTN_FWD_VARIANT=2
run_tests f0b1 VLB200_ATTN_BWD_TS=1
for v in "0 0" "1 0" "2 0" "0 1" "2 1"; do
  set -- $v
  echo "== fwd variant $1, bwd TS $2"
  VLB200_ATTN_FWD_VARIANT=$1 VLB200_ATTN_BWD_TS=$2 timeout 300 python tests/attn_probe2.py time 2>&1 | grep "^\[" 
done | tee gpurun_out/r2g_attn_timing.log
