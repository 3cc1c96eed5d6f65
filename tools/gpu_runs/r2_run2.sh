#!/bin/bash
# round-2 GPU call: ubench, attention A/B, the GPU test suite, parity report, bench A/B of the residual modes
set -x
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/r2_gpuinfo.txt
timeout 90 tools/ubench/sweep_poly > $O/r2_sweep_poly.txt 2>&1
V=taiwan-tongues-asr-ce_b200/lib/variants
timeout 300 python tools/attn_ab.py base=$V/attn_base.so nreg=$V/attn_nreg.so q0p8=$V/attn_q0p8.so all2=$V/attn_all2.so \
   all4=$V/attn_all4.so q01p8_pre2=$V/attn_q01p8_pre2.so q0p8_rest2=$V/attn_q0p8_rest2.so all4_pre4=$V/attn_all4_pre4.so \
   all2_pre4=$V/attn_all2_pre4.so q0p8_nonreg=$V/attn_q0p8_nonreg.so 32 > $O/r2_attn_ab.log 2>&1
timeout 500 python -m pytest tests -m gpu -q --timeout=240 -x -k "ops" 2>&1 | tail -40 > $O/r2_pytest_ops.log
timeout 800 python -m pytest tests -m gpu -q --timeout=300 -k "not ops" 2>&1 | tail -80 > $O/r2_pytest_rest.log
timeout 400 python tools/parity_report.py > $O/r2_parity_report.json 2> $O/r2_parity_report.err
for mode in split f32 bf16; do
  timeout 240 python bench.py --steps 5 --warmup 3 --residual $mode --no-cpu-baseline > $O/r2_bench_$mode.json 2> $O/r2_bench_$mode.err
done
tail -3 $O/r2_pytest_ops.log $O/r2_pytest_rest.log; cat $O/r2_attn_ab.log | tail -14
