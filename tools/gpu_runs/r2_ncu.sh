#!/bin/bash
# ncu evidence for round 2: launch list of ONE timed-shape step of the bench command (cudaProfilerStart/Stop around it:
# TTASR_PROFILE_STEP) + `--set full` captures of each kernel family at B = 256 (tools/ncu_target.py)
set -x
mkdir -p gpurun_out
O=gpurun_out
NCU=/usr/local/cuda/bin/ncu
TTASR_PROFILE_STEP=1 timeout 420 $NCU --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/r2_launches_bench_b256.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-baseline \
    > $O/r2_bench_under_ncu.log 2>&1
for tgt in ${NCU_TARGETS:-attention gemm_qkv_ln gemm_out_split gemm_fc1_ln gemm_fc2_split frontend}; do
  pat="gemm_kernel|attention_kernel|logmel_frames_kernel"
  timeout 200 $NCU --set full --clock-control none --import-source on -k regex:"$pat" -s 2 -c 1 -f -o $O/r2_prof_${tgt}_b256 \
      python tools/ncu_target.py $tgt 256 > $O/r2_ncu_${tgt}.log 2>&1
done
ls -la $O/*.ncu-rep
