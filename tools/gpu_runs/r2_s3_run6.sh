#!/bin/bash
# front end: software-pipelined generic mel loop
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_frontend.py -m gpu -q --timeout=200 2>&1 | grep -v "^$" | tail -5 > $O/r2s3c_pytest_frontend.log
timeout 120 python tools/frontend_bench.py 256 128 > $O/r2s3c_frontend_bench.log 2>&1
TTASR_FRONTEND_MEL=generic timeout 120 python tools/frontend_bench.py 64 80 >> $O/r2s3c_frontend_bench.log 2>&1
timeout 120 python tools/frontend_bench.py 64 80 >> $O/r2s3c_frontend_bench.log 2>&1
tail -3 $O/r2s3c_pytest_frontend.log; cat $O/r2s3c_frontend_bench.log
