#!/bin/bash
# 2-GPU sanity run of the streaming simulator: with one client per server the latency cannot be below one B = 1 forward
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29621 tools/streaming_sim.py --concurrency 1,2,4,16 --utterances 128 > gpurun_out/r2s3_streaming_2gpu.log 2>&1
tail -6 gpurun_out/r2s3_streaming_2gpu.log
