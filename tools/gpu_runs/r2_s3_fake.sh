#!/bin/bash
# diagnostic: how much of the attention kernel's time is the exponentials? (2 / 4 / 8 of every 8 made free)
set -x
mkdir -p gpurun_out
V=taiwan-tongues-asr-ce_b200/lib/variants
timeout 400 python tools/attn_ab.py base=$V/attn_base.so fake2=$V/attn_fake2.so fake4=$V/attn_fake4.so fake8=$V/attn_fake8.so 32 > gpurun_out/r2s3_attn_fake_ab.log 2>&1
tail -6 gpurun_out/r2s3_attn_fake_ab.log
