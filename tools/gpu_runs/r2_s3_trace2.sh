#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 120 python tools/attn_trace.py taiwan-tongues-asr-ce_b200/lib/variants/trace_fake8.so 32 > gpurun_out/r2s3_trace_fake8.txt 2>&1
head -60 gpurun_out/r2s3_trace_fake8.txt
