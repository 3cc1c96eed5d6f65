"""BASELINE.json configs[3]: long-form path — N synthetic 5-minute files cut into independent
30 s chunks (int16 PCM in pinned host memory, file-major order), sharded block-wise BY FILE over the GPUs of one box,
streamed through H2D -> log-mel -> large-v3 encoder with the copies overlapped; hidden states stay on the device
(a per-batch checksum stands in for the decoder that would consume them).  Timed region: everything after the host
buffers exist.  One process per GPU:
    python tools/longform_bench.py [--chunks 4096] [--batch 256]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/longform_bench.py"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "taiwan-tongues-asr-ce_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
import ttasr  # noqa: E402
from ttasr import dp  # noqa: E402

CHUNKS_PER_FILE = 10  # 5 min = 10 x 30 s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=4096)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--workload", default="large-v3")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    lo, hi = dp.shard_bounds(args.chunks, rank, world, keep_together=CHUNKS_PER_FILE)
    n = hi - lo
    cfg = ttasr.EncoderConfig.named(args.workload)
    fe = ttasr.B200WhisperFeatureExtractor(feature_size=cfg.num_mel_bins)
    enc = ttasr.B200WhisperEncoder(cfg, bench.make_gpu_weights(cfg, dev))
    pipe = ttasr.B200LogMelEncoder(fe, enc)
    # this rank's chunks, int16, pinned; the last file of the job is short (SURVEY: "last one short")
    g = torch.Generator(device=dev).manual_seed(4321 + rank)
    host = torch.empty((n, 480000), dtype=torch.int16).pin_memory()
    for i in range(0, n, 64):
        blk = (torch.randn((min(64, n - i), 480000), device=dev, generator=g) * 3277.0).clamp_(-32768, 32767).to(torch.int16)
        host[i: i + blk.shape[0]].copy_(blk)
    if hi == args.chunks and n > 0:
        host[-1, 240000:] = 0
    torch.cuda.synchronize()
    nb = n // args.batch
    batches = [host[k * args.batch: (k + 1) * args.batch] for k in range(nb)]
    tail = host[nb * args.batch:] if n % args.batch else None
    sums = []

    def consume(k, hidden):
        sums.append(hidden[:, ::97, ::31].float().sum())  # stands in for the decoder reading the batch on the GPU

    def run():
        sums.clear()
        pipe.stream_host(batches, None, consume)
        if tail is not None and tail.shape[0]:
            sums.append(pipe.encode_host(tail)[:, ::97, ::31].float().sum())

    run()  # warm-up pass (allocations, attribute setup)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    a.record()
    run()
    b.record()
    torch.cuda.synchronize()
    ms = dp.all_max(a.elapsed_time(b))
    wall = dp.all_max(time.time() - t0)
    checks = dp.gather_host({"rank": rank, "chunks": [lo, hi], "checksum": float(torch.stack(sums).sum().item())})
    if rank == 0:
        line = {"workload": f"long-form: {args.chunks} x 30 s chunks of {(args.chunks + 9) // 10} synthetic 5-min files, int16 PCM in "
                            f"pinned host memory, {args.workload}, block-wise by file over {world} GPU(s), batch {args.batch}",
                "n_gpus": world, "audio_s_per_s": args.chunks * 30.0 / (ms / 1e3), "ms": ms, "wall_s": wall,
                "h2d_bytes": args.chunks * 480000 * 2, "shards": checks}
        print(json.dumps(line))
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(line, open(os.path.join(ROOT, "gpurun_out", f"longform_{world}gpu.json"), "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
