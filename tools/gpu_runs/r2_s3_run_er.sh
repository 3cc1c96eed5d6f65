#!/bin/bash
# attention: token released one quarter early (TTASR_ATTN_EARLY_RELEASE)
set -x
mkdir -p gpurun_out
O=gpurun_out
V=taiwan-tongues-asr-ce_b200/lib/variants
TTASR_LIB_PATH=$PWD/$V/attn_er.so timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=120 -k "attention" 2>&1 | grep -v "^$" | tail -3 > $O/r2s3_attn_er_pytest.log
timeout 400 python tools/attn_ab.py base=$V/attn_base.so er=$V/attn_er.so er_p0=$V/attn_er_p0.so er_st=$V/attn_er_st.so er_st_p0=$V/attn_er_st_p0.so 32 > $O/r2s3_attn_er_ab.log 2>&1
tail -2 $O/r2s3_attn_er_pytest.log; tail -8 $O/r2s3_attn_er_ab.log
