#!/bin/bash
# round-2 multi-GPU measurements on one 8 x B200 box (gpurun --gpus 8): bench at N = 8 with per-rank attribution and the
# cross-rank probe, long-form (configs[3]) at 8 and 4 GPUs, streaming load shape (configs[4]) over 8 single-GPU servers
set -x
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name --format=csv > $O/r2_8gpu_info.txt
timeout 300 $TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 8 --warmup 3 > $O/r2_bench_8gpu.json 2> $O/r2_bench_8gpu.err
timeout 240 $TR --nproc-per-node 8 --master-port 29612 tools/longform_bench.py > $O/r2_longform_8gpu.log 2>&1
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 300 $TR --nproc-per-node 4 --master-port 29613 tools/longform_bench.py > $O/r2_longform_4gpu.log 2>&1
timeout 420 $TR --nproc-per-node 8 --master-port 29614 tools/streaming_sim.py --concurrency 1,10,64,256 --utterances 512 > $O/r2_streaming_8gpu.log 2>&1
tail -2 $O/r2_longform_8gpu.log $O/r2_longform_4gpu.log; tail -5 $O/r2_streaming_8gpu.log
