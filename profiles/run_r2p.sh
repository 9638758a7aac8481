#!/bin/bash
# r2p: tensor side of the backward alone (elementwise warps only hand the barriers over): DBG 32, and 36 = also no streamed-tile loads
mkdir -p gpurun_out
for v in 0 32 36; do
  echo "== bwd dbg $v"
  VLB200_ATTN_BWD_DBG=$v timeout 300 python tests/attn_probe2.py time 2>&1 | grep "^\[config2"
done | tee gpurun_out/r2p_attn_bwd.log
