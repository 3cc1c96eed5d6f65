#!/bin/bash
# compute-sanitizer pass over the kernels (memcheck on the per-kernel op tests + a small encoder / front-end subset,
# racecheck on the front end).  Writes gpurun_out/r2_sanitizer_*.log; summarised under profiles/.
#   gpurun --timeout 1500 -- bash tools/gpu_runs/sanitize.sh
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
SEL_OPS='test_gemm and (m300n384k384 or m128n128k64) or test_attention and (b1t128h1 or b2t300h2) or test_layernorm and 7-128 or test_gemm_split_residual and m300n384k384 or test_gemm_with_folded_layernorm and m200n512k128 or test_conv_stem_against_conv1d and b1t128d128'
timeout 900 $CS --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_ops.py -q -x --timeout=600 -k "$SEL_OPS" > gpurun_out/r2_sanitizer_memcheck_ops.log 2>&1
echo "rc=$?" >> gpurun_out/r2_sanitizer_memcheck_ops.log
timeout 900 $CS --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_encoder.py -q -x --timeout=800 -k "config1_clips or ragged or clamp_pass or (matches_fp32_oracle and micro) or (residual_modes and micro)" > gpurun_out/r2_sanitizer_memcheck_path.log 2>&1
echo "rc=$?" >> gpurun_out/r2_sanitizer_memcheck_path.log
timeout 600 $CS --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_frontend.py -q -x --timeout=500 -k "config1_clips or padding_tiles or compiled_in" > gpurun_out/r2_sanitizer_racecheck_frontend.log 2>&1
echo "rc=$?" >> gpurun_out/r2_sanitizer_racecheck_frontend.log
tail -4 gpurun_out/r2_sanitizer_*.log
