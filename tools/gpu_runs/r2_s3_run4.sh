#!/bin/bash
# round-2 session 3, call 4: front end with the compiled-in mel projection: parity, timing, ncu; GEMM wave threshold at B = 1
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_frontend.py tests/test_ingest.py -m gpu -q --timeout=200 2>&1 | grep -v "^$" | tail -8 > $O/r2s3b_pytest_frontend.log
timeout 120 python tools/frontend_bench.py 256 128 > $O/r2s3b_frontend_bench.log 2>&1
timeout 120 python tools/frontend_bench.py 64 80 >> $O/r2s3b_frontend_bench.log 2>&1
TTASR_FRONTEND_MEL=generic timeout 120 python tools/frontend_bench.py 256 128 > $O/r2s3b_frontend_bench_generic.log 2>&1
timeout 200 /usr/local/cuda/bin/ncu --set full --clock-control none --import-source on -k regex:logmel_frames_kernel -s 2 -c 1 -f \
    -o $O/r2s3b_prof_frontend_b256 python tools/ncu_target.py frontend 256 > $O/r2s3b_ncu_frontend.log 2>&1
for pct in 60 40 30; do
  TTASR_GEMM_WAVE_MIN_PCT=$pct timeout 200 python tools/latency_small_batch.py 1 2 > $O/r2s3b_latency_wave$pct.log 2>&1
done
tail -4 $O/r2s3b_pytest_frontend.log; cat $O/r2s3b_frontend_bench.log $O/r2s3b_frontend_bench_generic.log; cut -c1-160 $O/r2s3b_latency_wave*.log
