#!/bin/bash
# B = 1 (M = 1500) GEMMs under ncu: is the L2 the limiter? (lts throughput, tensor pipe, time)
set -x
mkdir -p gpurun_out
O=gpurun_out
NCU=/usr/local/cuda/bin/ncu
for tgt in gemm_fc2_split gemm_qkv_ln gemm_fc1_ln gemm_out_split; do
  timeout 200 $NCU --set full --clock-control none -k regex:gemm_kernel -s 2 -c 1 -f -o $O/r2s3_b1_${tgt} \
      python tools/ncu_target.py $tgt 1 > $O/r2s3_b1_ncu_${tgt}.log 2>&1
done
ls -la $O/r2s3_b1_*.ncu-rep
