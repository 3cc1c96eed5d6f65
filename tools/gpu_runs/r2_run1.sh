set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_gpuinfo.txt
timeout 120 tools/ubench/sweep_poly > gpurun_out/r2_sweep_poly.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -x -k "ops" 2>&1 | tail -40 > gpurun_out/r2_pytest_ops.log
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 -k "not ops" 2>&1 | tail -60 > gpurun_out/r2_pytest_rest.log
timeout 600 python tools/parity_report.py > gpurun_out/r2_parity_report.json 2> gpurun_out/r2_parity_report.err
for mode in split f32 bf16; do
  timeout 300 python bench.py --steps 5 --warmup 3 --residual $mode --no-cpu-baseline > gpurun_out/r2_bench_$mode.json 2> gpurun_out/r2_bench_$mode.err
done
tail -5 gpurun_out/r2_pytest_ops.log gpurun_out/r2_pytest_rest.log
