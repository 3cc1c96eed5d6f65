#!/bin/bash
# ncu --set full with source-level stall sampling of the attention kernel (B = 64)
set -x
mkdir -p gpurun_out
timeout 300 /usr/local/cuda/bin/ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 2 -c 1 -f \
    -o gpurun_out/r2s3_prof_attention_b64 python tools/ncu_target.py attention 64 > gpurun_out/r2s3_ncu_attention.log 2>&1
ls -la gpurun_out/r2s3_prof_attention_b64.ncu-rep
