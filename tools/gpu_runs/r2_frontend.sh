#!/bin/bash
# front end after the packed-butterfly change: parity tests, timing alone, ncu full capture
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_frontend.py tests/test_ingest.py -m gpu -q --timeout=200 2>&1 | tail -5 > $O/r2_pytest_frontend.log
timeout 120 python tools/frontend_bench.py 256 128 > $O/r2_frontend_bench.log 2>&1
timeout 120 python tools/frontend_bench.py 64 80 >> $O/r2_frontend_bench.log 2>&1
timeout 200 /usr/local/cuda/bin/ncu --set full --clock-control none --import-source on -k regex:logmel_frames_kernel -s 2 -c 1 -f \
    -o $O/r2_prof_frontend_b256 python tools/ncu_target.py frontend 256 > $O/r2_ncu_frontend.log 2>&1
tail -3 $O/r2_pytest_frontend.log; cat $O/r2_frontend_bench.log
