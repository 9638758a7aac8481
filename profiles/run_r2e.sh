#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_xc2.py -m gpu -q -k "swiglu_bwd or config5" 2>&1 | tail -4
bash profiles/run_multi.sh 1 qwen7b xc2_7b next7b_lora 7b_lora 2>&1 | grep -v "^\*\*\*\|OMP_NUM"
for f in 0 1 0 1; do
  VLB200_FUSE_SWIGLU_BWD=$f timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-library-baseline --skip-plugin --skip-e2e > gpurun_out/r2e_fuse$f.json 2>/dev/null
  python -c "
import json; d=json.loads(open('gpurun_out/r2e_fuse$f.json').read().strip().splitlines()[-1]); print('fuse $f ms/step', round(d['ms_per_step'],1), d['clocks']['sm_mhz'])"
done
