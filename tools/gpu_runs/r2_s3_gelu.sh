#!/bin/bash
# what the GELU epilogue costs fc1: the same GEMM with and without the activation, B = 256 and 32, isolated
set -x
mkdir -p gpurun_out
L=taiwan-tongues-asr-ce_b200/lib/libttasr_b200.so
for B in 32 256; do
  for shape in fc1 fc1_noact qkv; do
    echo "== $shape B=$B" >> gpurun_out/r2s3_gelu_cost.log
    timeout 120 python tools/gemm_ab.py cur=$L --shape=$shape $B >> gpurun_out/r2s3_gelu_cost.log 2>&1
  done
done
cat gpurun_out/r2s3_gelu_cost.log
