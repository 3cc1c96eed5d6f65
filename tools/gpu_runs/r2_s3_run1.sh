#!/bin/bash
# round-2 session 3, call 1: whole GPU suite, front end after the packed-butterfly / twiddle-table commits (timing + ncu),
# driver-style bench lines for configs[2] and configs[1], small-batch latency with per-stage breakdown
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 2>&1 | grep -v "^$" | tail -15 > $O/r2s3_pytest_gpu.log
timeout 120 python tools/frontend_bench.py 256 128 > $O/r2s3_frontend_bench.log 2>&1
timeout 120 python tools/frontend_bench.py 64 80 >> $O/r2s3_frontend_bench.log 2>&1
timeout 200 /usr/local/cuda/bin/ncu --set full --clock-control none --import-source on -k regex:logmel_frames_kernel -s 2 -c 1 -f \
    -o $O/r2s3_prof_frontend_b256 python tools/ncu_target.py frontend 256 > $O/r2s3_ncu_frontend.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r2s3_bench_large-v3_b256.json 2> $O/r2s3_bench_large.err
timeout 300 python bench.py --workload small --batch 64 --steps 200 --warmup 5 > $O/r2s3_bench_small_b64.json 2> $O/r2s3_bench_small.err
timeout 200 python tools/latency_small_batch.py > $O/r2s3_latency.log 2>&1
tail -4 $O/r2s3_pytest_gpu.log; cat $O/r2s3_frontend_bench.log; cat $O/r2s3_bench_large-v3_b256.json; tail -3 $O/r2s3_bench_large.err
