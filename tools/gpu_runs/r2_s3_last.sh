#!/bin/bash
# last check of the round: whole GPU suite + smoke on the committed build, configs[1] bench line
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 2>&1 | grep -v "^$" | tail -4 > $O/r2l_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/r2l_smoke.log 2>&1
timeout 300 python bench.py --workload small --batch 64 --steps 200 --warmup 5 > $O/r2l_bench_small_b64.json 2> $O/r2l_bench_small.err
tail -3 $O/r2l_pytest_gpu.log; tail -1 $O/r2l_smoke.log; cut -c1-200 $O/r2l_bench_small_b64.json
