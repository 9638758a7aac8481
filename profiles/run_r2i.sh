#!/bin/bash
# r2i: which side bounds the attention backward (TS kernels)?  DBG mask: 1 = no MMAs, 2 = no exponentials, 4 = no streamed-tile loads
mkdir -p gpurun_out
for v in 0 1 2 4 3 7; do
  echo "== bwd TS dbg $v"
  VLB200_ATTN_BWD_TS=1 VLB200_ATTN_BWD_DBG=$v timeout 300 python tests/attn_probe2.py time 2>&1 | grep "^\[config2"
done | tee gpurun_out/r2i_attn_bwd_timing.log
