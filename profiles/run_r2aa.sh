#!/bin/bash
# r2aa: third-generation forward (variant 5: eight softmax warps)
mkdir -p gpurun_out
T="tests/test_gpu_kernels.py tests/test_gpu_share_prefix.py tests/test_gpu_engine.py"
K="attention or share_prefix_rows_kernel or shared_step_equals or config1 or golden"
VLB200_ATTN_FWD_VARIANT=5 timeout 900 python -m pytest $T -m gpu -q -x -k "$K" > gpurun_out/r2aa_tests.log 2>&1
echo "tests[fwd 5] rc=$? $(tail -1 gpurun_out/r2aa_tests.log)"
grep -n "^FAILED\|^E  .*rel l2\|watchdog\|Error" gpurun_out/r2aa_tests.log | head -8
{
for v in 4 5; do echo "== fwd variant $v"; VLB200_ATTN_FWD_VARIANT=$v timeout 300 python tests/attn_probe2.py time 2>&1 | grep "^\["; done
echo "== phases: forward variant 5"; VLB200_ATTN_FWD_VARIANT=85 timeout 300 python tests/attn_phase_probe.py
} 2>&1 | grep -v Warning | tee gpurun_out/r2aa_attn.log
timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --model qwen_small --no-cpu-baseline --no-library-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('qwen_small', d['ms_per_step'], d['config'].get('step_tflop_executed'), d['config'].get('step_tflop_algorithmic'))"
