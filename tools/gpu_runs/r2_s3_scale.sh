#!/bin/bash
# round-2 final build: bench at N = 2 and N = 4 on one box (the N = 8 line is profiles/r2_bench_large-v3_b256_8gpu.json)
set -x
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
CUDA_VISIBLE_DEVICES=0,1 timeout 300 $TR --nproc-per-node 2 --master-port 29631 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2f_bench_2gpu.json 2> $O/r2f_bench_2gpu.err
timeout 300 $TR --nproc-per-node 4 --master-port 29632 bench.py --gpus 4 --steps 10 --warmup 3 > $O/r2f_bench_4gpu.json 2> $O/r2f_bench_4gpu.err
cut -c1-260 $O/r2f_bench_2gpu.json $O/r2f_bench_4gpu.json; tail -2 $O/r2f_bench_4gpu.err
