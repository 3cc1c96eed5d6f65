#!/bin/bash
# in-kernel timelines of the attention variants (base = whole row max before the sweep; lm = late max)
set -x
mkdir -p gpurun_out
O=gpurun_out
V=taiwan-tongues-asr-ce_b200/lib/variants
for n in base lm lm_n; do
  timeout 120 python tools/attn_trace.py $V/trace_$n.so 32 > $O/r2s3_trace_$n.txt 2>&1
done
head -16 $O/r2s3_trace_base.txt; head -18 $O/r2s3_trace_lm.txt; head -18 $O/r2s3_trace_lm_n.txt
