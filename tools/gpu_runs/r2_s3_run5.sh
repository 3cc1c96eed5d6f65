#!/bin/bash
# attention: token passed per scheduler (TTASR_ATTN_SMSP_TOKEN) — op tests on the variant library, then interleaved A/B
set -x
mkdir -p gpurun_out
O=gpurun_out
V=taiwan-tongues-asr-ce_b200/lib/variants
TTASR_LIB_PATH=$PWD/$V/attn_st.so timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=120 -k "attention" 2>&1 | grep -v "^$" | tail -8 > $O/r2s3_attn_st_pytest.log
timeout 400 python tools/attn_ab.py base=$V/attn_base.so st=$V/attn_st.so st_p2=$V/attn_st_p2.so st_p0=$V/attn_st_p0.so st_q2=$V/attn_st_q2.so \
    st_lm=$V/attn_st_lm.so 32 > $O/r2s3_attn_st_ab.log 2>&1
tail -4 $O/r2s3_attn_st_pytest.log; tail -14 $O/r2s3_attn_st_ab.log
