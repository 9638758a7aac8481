#!/bin/bash
mkdir -p gpurun_out
python tests/attn_probe2.py time 2>&1 | tee gpurun_out/attn_probe_time.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 6 -c 3 -f -o gpurun_out/attn_prof python tests/attn_probe2.py ncu > gpurun_out/attn_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/attn_prof.ncu-rep
