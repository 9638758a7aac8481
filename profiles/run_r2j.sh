#!/bin/bash
# r2j: per-phase cycle counts of the softmax / elementwise threads (DBG-8 variants)
mkdir -p gpurun_out
{
echo "== forward, P through smem (variant 0)"; VLB200_ATTN_FWD_VARIANT=80 timeout 300 python tests/attn_phase_probe.py
echo "== forward, P and Q in TMEM (variant 2)"; VLB200_ATTN_FWD_VARIANT=82 timeout 300 python tests/attn_phase_probe.py
echo "== backward, TS"; VLB200_ATTN_BWD_TS=1 VLB200_ATTN_BWD_DBG=8 timeout 300 python tests/attn_phase_probe.py
} 2>&1 | grep -v Warning | tee gpurun_out/r2j_attn_phases.log
