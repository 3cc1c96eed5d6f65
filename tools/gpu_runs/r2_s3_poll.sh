#!/bin/bash
# attention: polling mbarrier waits (TTASR_ATTN_POLL)
set -x
mkdir -p gpurun_out
O=gpurun_out
V=taiwan-tongues-asr-ce_b200/lib/variants
TTASR_LIB_PATH=$PWD/$V/attn_poll.so timeout 200 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=60 -k "attention" 2>&1 | grep -v "^$" | tail -3 > $O/r2s3_attn_poll_pytest.log
timeout 300 python tools/attn_ab.py base=$V/attn_base.so poll=$V/attn_poll.so poll_mma2=$V/attn_poll_mma2.so fake8=$V/attn_fake8.so poll_fake8=$V/attn_poll_fake8.so 32 > $O/r2s3_attn_poll_ab.log 2>&1
tail -2 $O/r2s3_attn_poll_pytest.log; tail -7 $O/r2s3_attn_poll_ab.log
