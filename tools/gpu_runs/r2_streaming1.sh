#!/bin/bash
# streaming load shape on one GPU (levels 1 / 10 / 64 / 256) + final single-GPU bench lines for profiles/
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python tools/streaming_sim.py --concurrency 1,10,64,256 --utterances 384 > $O/r2_streaming_1gpu.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r2_bench_large-v3_b256.json 2> $O/r2_bench_large.err
timeout 300 python bench.py --workload small --batch 64 --steps 200 --warmup 5 > $O/r2_bench_small_b64.json 2> $O/r2_bench_small.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err
tail -5 $O/r2_streaming_1gpu.log
