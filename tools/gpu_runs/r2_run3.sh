#!/bin/bash
# round-2 GPU call 3: split-GEMM staging fix + PDL: ops / encoder tests, bench A/B, small workload, small-batch latency
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_ops.py tests/test_gpu_encoder.py tests/test_gpu_bench_parity.py -m gpu -q -s --timeout=300 -x 2>&1 | grep -v "^$" | tail -60 > $O/r2c_pytest.log
for mode in split f32; do
  timeout 240 python bench.py --steps 6 --warmup 3 --residual $mode --no-cpu-baseline --no-library-baseline > $O/r2c_bench_$mode.json 2> $O/r2c_bench_$mode.err
done
timeout 200 python bench.py --workload small --batch 64 --steps 200 --warmup 5 --no-cpu-baseline > $O/r2c_bench_small_b64.json 2> $O/r2c_bench_small.err
timeout 300 python tools/latency_small_batch.py > $O/r2c_latency.log 2>&1
tail -3 $O/r2c_pytest.log; tail -7 $O/r2c_latency.log
