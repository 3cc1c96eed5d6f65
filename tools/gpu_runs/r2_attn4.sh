#!/bin/bash
# four-warpgroup attention kernel: correctness on the op tests' shapes, then A/B against the two-warpgroup kernel
set -x
mkdir -p gpurun_out
O=gpurun_out
TTASR_ATTN_KERNEL=4wg timeout 200 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=60 -k "attention" 2>&1 | grep -v "^$" | tail -25 > $O/r2_attn4_pytest.log
V=taiwan-tongues-asr-ce_b200/lib/variants
timeout 300 python tools/attn_ab.py v2=$V/attn_v2.so v4=$V/attn_v4.so v4p2=$V/attn_v4p2.so v4p4=$V/attn_v4p4.so 32 > $O/r2_attn4_ab.log 2>&1
tail -12 $O/r2_attn4_pytest.log; tail -12 $O/r2_attn4_ab.log
