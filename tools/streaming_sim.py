"""BASELINE.json configs[4]: the streaming server's load shape on one GPU, or — under
`python -m torch.distributed.run --nproc-per-node N tools/streaming_sim.py ...` — on N GPUs of one box, one server
process per GPU with the clients of each concurrency level dealt round-robin over the processes (what a load balancer
in front of N single-GPU servers does; there is no cross-GPU traffic on the data path) — utterances of
U ~ Uniform[1 s, 5 s] (int16, as they sit in client.scratch_buffer) arrive as a Poisson process from `concurrency`
clients; B200ASR's cross-client micro-batcher (window 5 ms) encodes whatever is ready in one launch group (CUDA graph
per batch-size bucket).  The decoder is a stub that waits for the hidden states on the GPU (decode is out of scope).
Reports utterances/s, real-speech-seconds/s, audio-s/s (30 s windows) and the latency distribution submit -> hidden
states complete.   python tools/streaming_sim.py [--concurrency 1,10,64] [--utterances 300] [--eager]"""
import argparse
import asyncio
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "taiwan-tongues-asr-ce_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import ttasr  # noqa: E402
import ttasr.asr_plugin  # noqa: E402
from ttasr.asr_plugin import B200ASR  # noqa: E402


async def client_loop(asr, cid, n_utts, rng, lat, speech):
    for _ in range(n_utts):
        n = int(rng.integers(16000, 80001))
        pcm = (rng.standard_normal(n) * 3000).astype("<i2")
        c = types.SimpleNamespace(scratch_buffer=bytearray(pcm.tobytes()), samples_width=2, last_start_time=0.0, client_id=cid)
        t0 = time.perf_counter()
        r = await asr.transcribe(c)
        if r is None:  # the plugin swallows errors as the reference's wrapper does; surface the cause here
            hidden, _ = asr.batcher.encode_batch([ttasr.asr_plugin.Utterance(ttasr.asr_plugin.pcm_bytes_to_tensor(c.scratch_buffer, 2), c)])
            raise RuntimeError("transcribe returned None")
        lat.append(time.perf_counter() - t0)
        speech.append(n / 16000.0)
        await asyncio.sleep(float(rng.exponential(0.002)))  # think time between a client's utterances


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--concurrency", default="1,10,64")
    ap.add_argument("--utterances", type=int, default=300, help="total utterances per concurrency level")
    ap.add_argument("--workload", default="large-v3")
    ap.add_argument("--eager", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
    cfg = ttasr.EncoderConfig.named(args.workload)
    fe = ttasr.B200WhisperFeatureExtractor(feature_size=cfg.num_mel_bins)
    enc = ttasr.B200WhisperEncoder(cfg, bench.make_gpu_weights(cfg, dev))
    pipe = ttasr.B200LogMelEncoder(fe, enc)

    def decode(hidden, info):  # stand-in for the host decoder: the hidden states must be complete on the GPU
        # (runs in the plugin's decode worker thread, whose current device is not this server's: name the device)
        torch.cuda.current_stream(hidden.device).synchronize()
        return {"text": "x", "words": []}

    out = {}
    from ttasr.dp import gather_host

    for total_conc in [int(c) for c in args.concurrency.split(",")]:
        conc = total_conc // world + (1 if rank < total_conc % world else 0)   # this server's share of the clients
        if world > 1:
            dist.barrier()
        if conc == 0:
            gather_host(None)
            continue
        asr = B200ASR(pipe, decode, batch_window_s=0.005, max_batch=max(conc, 1), use_graphs=not args.eager)
        asr.warm_up()
        rng = np.random.default_rng(777)
        per_client = max(1, args.utterances // total_conc)

        async def warm():
            await asyncio.gather(*(client_loop(asr, i, 1, np.random.default_rng(i), [], []) for i in range(conc)))

        asyncio.run(warm())  # builds the graph buckets this level uses
        lat, speech = [], []

        async def run():
            await asyncio.gather(*(client_loop(asr, i, per_client, np.random.default_rng(777 + rank * 1000 + i), lat, speech)
                                   for i in range(conc)))

        l0, e0 = asr.batcher.launches, asr.batcher.encoded
        t0 = time.perf_counter()
        asyncio.run(run())
        wall = time.perf_counter() - t0
        parts = gather_host({"lat": lat, "speech": float(np.sum(speech)), "wall": wall, "clients": conc,
                             "mean_batch": (asr.batcher.encoded - e0) / max(1, asr.batcher.launches - l0)})
        if rank != 0:
            continue
        parts = [p for p in parts if p]
        lat_all = [x for p in parts for x in p["lat"]]
        wall = max(p["wall"] for p in parts)
        ls = np.sort(np.array(lat_all)) * 1e3
        out[total_conc] = {"utterances": len(lat_all), "servers": len(parts), "utterances_per_s": len(lat_all) / wall,
                           "speech_s_per_s": sum(p["speech"] for p in parts) / wall,
                           "audio_s_per_s_30s_windows": 30.0 * len(lat_all) / wall,
                           "mean_batch": float(np.mean([p["mean_batch"] for p in parts])),
                           "latency_ms": {"p50": float(ls[len(ls) // 2]), "p90": float(ls[int(0.9 * len(ls))]),
                                          "p99": float(ls[min(len(ls) - 1, int(0.99 * len(ls)))]), "max": float(ls[-1])}}
        print(total_conc, json.dumps(out[total_conc]), flush=True)
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        name = "streaming_sim" + ("_eager" if args.eager else "") + (f"_{world}gpu" if world > 1 else "") + ".json"
        json.dump({"mode": "eager" if args.eager else "cuda-graph buckets", "workload": args.workload, "n_gpus": world,
                   "levels": out}, open(os.path.join(ROOT, "gpurun_out", name), "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
