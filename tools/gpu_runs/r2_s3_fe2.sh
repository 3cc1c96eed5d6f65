#!/bin/bash
# front end: Hermitian split from the thread's own registers (half the Z round trip), twiddle symmetry in pass 1
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_frontend.py tests/test_ingest.py tests/test_gpu_bench_parity.py -m gpu -q --timeout=300 2>&1 | grep -v "^$" | tail -5 > $O/r2s3e_pytest_frontend.log
timeout 120 python tools/frontend_bench.py 256 128 > $O/r2s3e_frontend_bench.log 2>&1
timeout 120 python tools/frontend_bench.py 64 80 >> $O/r2s3e_frontend_bench.log 2>&1
timeout 120 python tools/parity_report.py --frontend-only > $O/r2s3e_parity_fe.log 2>&1 || true
tail -3 $O/r2s3e_pytest_frontend.log; cat $O/r2s3e_frontend_bench.log
