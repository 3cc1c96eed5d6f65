#!/bin/bash
# round-2 GPU call 4: 32-column split epilogue: ops / encoder tests, bench split vs f32
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_ops.py tests/test_gpu_encoder.py -m gpu -q --timeout=300 -x 2>&1 | grep -v "^$" | tail -15 > $O/r2d_pytest.log
for mode in split f32; do
  timeout 240 python bench.py --steps 6 --warmup 3 --residual $mode --no-cpu-baseline --no-library-baseline > $O/r2d_bench_$mode.json 2> $O/r2d_bench_$mode.err
done
tail -3 $O/r2d_pytest.log
