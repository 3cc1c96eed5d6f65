#!/bin/bash
# round-2 final state: whole GPU suite + smoke, driver-style bench lines (configs[2], configs[1]), front end alone + its
# ncu capture, launch list of one timed-shape step, small-batch latency
set -x
mkdir -p gpurun_out
O=gpurun_out
NCU=/usr/local/cuda/bin/ncu
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 2>&1 | grep -v "^$" | tail -6 > $O/r2f_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/r2f_smoke.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 > $O/r2f_bench_large-v3_b256.json 2> $O/r2f_bench_large.err
timeout 300 python bench.py --workload small --batch 64 --steps 200 --warmup 5 > $O/r2f_bench_small_b64.json 2> $O/r2f_bench_small.err
timeout 120 python tools/frontend_bench.py 256 128 > $O/r2f_frontend_bench.log 2>&1
cp $O/frontend_bench.json $O/r2f_frontend_bench_b256_128.json
timeout 120 python tools/frontend_bench.py 64 80 >> $O/r2f_frontend_bench.log 2>&1
cp $O/frontend_bench.json $O/r2f_frontend_bench_b64_80.json
timeout 200 python tools/latency_small_batch.py > $O/r2f_latency.log 2>&1
cp $O/latency_small_batch.json $O/r2f_latency_small_batch.json
timeout 200 $NCU --set full --clock-control none --import-source on -k regex:logmel_frames_kernel -s 2 -c 1 -f \
    -o $O/r2f_prof_frontend_b256 python tools/ncu_target.py frontend 256 > $O/r2f_ncu_frontend.log 2>&1
TTASR_PROFILE_STEP=1 timeout 420 $NCU --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/r2f_launches_bench_b256.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-baseline \
    > $O/r2f_bench_under_ncu.log 2>&1
tail -3 $O/r2f_pytest_gpu.log; tail -2 $O/r2f_smoke.log; cat $O/r2f_frontend_bench.log; cut -c1-300 $O/r2f_bench_large-v3_b256.json
