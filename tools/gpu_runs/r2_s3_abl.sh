#!/bin/bash
# attention chain ablations (diagnostic builds, wrong results): bf16 conversion vs row sums
set -x
mkdir -p gpurun_out
V=taiwan-tongues-asr-ce_b200/lib/variants
timeout 400 python tools/attn_ab.py base=$V/attn_base.so abl_pack=$V/attn_abl_pack.so abl_sum=$V/attn_abl_sum.so 32 > gpurun_out/r2s3_attn_abl2_ab.log 2>&1
tail -5 gpurun_out/r2s3_attn_abl2_ab.log
