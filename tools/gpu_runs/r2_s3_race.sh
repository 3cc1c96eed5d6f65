#!/bin/bash
# racecheck of the front end after the staging-pitch fix (baked 80-filter path, generic program, padding tiles), then timing
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 600 $CS --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_frontend.py -q -x --timeout=500 -k "config1_clips or padding_tiles or compiled_in" > gpurun_out/r2_sanitizer_racecheck_frontend.log 2>&1
echo "rc=$?" >> gpurun_out/r2_sanitizer_racecheck_frontend.log
grep -E "passed|failed|RACECHECK SUMMARY|rc=" gpurun_out/r2_sanitizer_racecheck_frontend.log | tail -4
timeout 120 python tools/frontend_bench.py 256 128 2>&1 | head -2
timeout 120 python tools/frontend_bench.py 64 80 2>&1 | head -2
