#!/bin/bash
mkdir -p gpurun_out
for n in 296 148 74 37; do echo "== max CTAs $n"; VLB200_ATTN_MAX_CTAS=$n python tests/attn_probe3.py 2>&1 | tail -3; done | tee gpurun_out/attn_exp.log
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -x -k "config1" 2>&1 | tail -30
