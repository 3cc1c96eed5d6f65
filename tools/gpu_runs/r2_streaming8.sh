#!/bin/bash
# streaming load shape (configs[4]) over 8 single-GPU servers, after the decode-stub device fix
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 420 $TR --nproc-per-node 8 --master-port 29614 tools/streaming_sim.py --concurrency 1,10,64,256 --utterances 512 > gpurun_out/r2_streaming_8gpu.log 2>&1
tail -5 gpurun_out/r2_streaming_8gpu.log
