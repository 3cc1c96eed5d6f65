#!/bin/bash
# attention: one MMA-issuing warp per query tile (TTASR_ATTN_TWO_MMA)
set -x
mkdir -p gpurun_out
O=gpurun_out
V=taiwan-tongues-asr-ce_b200/lib/variants
TTASR_LIB_PATH=$PWD/$V/attn_mma2.so timeout 200 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=60 -k "attention" 2>&1 | grep -v "^$" | tail -3 > $O/r2s3_attn_mma2_pytest.log
timeout 300 python tools/attn_ab.py base=$V/attn_base.so mma2=$V/attn_mma2.so mma2_p0=$V/attn_mma2_p0.so mma2_st=$V/attn_mma2_st.so fake8=$V/attn_fake8.so mma2_fake8=$V/attn_mma2_fake8.so 32 > $O/r2s3_attn_mma2_ab.log 2>&1
tail -2 $O/r2s3_attn_mma2_pytest.log; tail -8 $O/r2s3_attn_mma2_ab.log
