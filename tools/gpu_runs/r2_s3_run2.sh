#!/bin/bash
# round-2 session 3, call 2: attention with the row max taken beside quarter 0's exponentials (TTASR_ATTN_LATEMAX):
# op tests, A/B of the variants; GEMM tile cost model at small batch
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=120 -k "attention" 2>&1 | grep -v "^$" | tail -8 > $O/r2s3_attn_pytest.log
V=taiwan-tongues-asr-ce_b200/lib/variants
timeout 400 python tools/attn_ab.py base=$V/attn_base.so lm=$V/attn_lm.so lm_n=$V/attn_lm_n.so lm_p2=$V/attn_lm_p2.so lm_p2n=$V/attn_lm_p2n.so \
    lm_q2=$V/attn_lm_q2.so lm_p2q2=$V/attn_lm_p2q2.so lm_q4=$V/attn_lm_q4.so 32 > $O/r2s3_attn_ab.log 2>&1
timeout 200 python tools/latency_small_batch.py 1 2 4 > $O/r2s3_latency_model0.log 2>&1
TTASR_GEMM_TILE_MODEL=1 timeout 200 python tools/latency_small_batch.py 1 2 4 > $O/r2s3_latency_model1.log 2>&1
tail -4 $O/r2s3_attn_pytest.log; tail -12 $O/r2s3_attn_ab.log; cut -c1-200 $O/r2s3_latency_model0.log; cut -c1-200 $O/r2s3_latency_model1.log
