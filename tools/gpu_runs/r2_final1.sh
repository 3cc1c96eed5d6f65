#!/bin/bash
# bench contract test + the single-GPU bench lines kept under profiles/ (driver-style invocations)
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_bench_contract.py -q -m gpu --timeout=900 2>&1 | tail -5 > $O/r2_pytest_bench_contract.log
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r2_bench_large-v3_b256.json 2> $O/r2_bench_large.err
timeout 300 python bench.py --workload small --batch 64 --steps 200 --warmup 5 > $O/r2_bench_small_b64.json 2> $O/r2_bench_small.err
tail -3 $O/r2_pytest_bench_contract.log
