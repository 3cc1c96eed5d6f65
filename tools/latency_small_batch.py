"""Config-5 shape: latency of PCM -> hidden states for small batches of short utterances (large-v3), eager vs CUDA
graph replay.   python tools/latency_small_batch.py  -> gpurun_out/latency_small_batch.json"""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "taiwan-tongues-asr-ce_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
import ttasr  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    cfg = ttasr.EncoderConfig.named("large-v3")
    fe = ttasr.B200WhisperFeatureExtractor(feature_size=cfg.num_mel_bins)
    enc = ttasr.B200WhisperEncoder(cfg, bench.make_gpu_weights(cfg, dev))
    pipe = ttasr.B200LogMelEncoder(fe, enc)
    enc.reserve_workspace(32)  # captured graphs bake the workspace address in: size it once for the largest batch
    out = {}
    batches = tuple(int(a) for a in sys.argv[1:] if a.isdigit()) or (1, 2, 4, 8, 16, 32)
    for B in batches:
        pcm = (torch.randn((B, 80000), device=dev) * 3000).to(torch.int16)  # 5 s utterances, int16 wire format
        nv = torch.full((B,), 80000, dtype=torch.int32, device=dev)

        def run():
            return pipe.encode_device(pcm, n_valid=nv)

        for _ in range(3):
            run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            o = run()
        tg = []
        for _ in range(20):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b.record(); torch.cuda.synchronize()
            tg.append(a.elapsed_time(b))
        stages = None
        if B in (1, 8):  # where the time goes, eager (CUDA events around every launch, so the sum exceeds the graph time)
            enc.profile(True); enc.profile_read(reset=True)
            for _ in range(10):
                run()
            torch.cuda.synchronize()
            stages = {k: {"us_per_launch": round(v[0] / max(v[1], 1) * 1e3, 2), "launches_per_forward": v[1] // 10,
                          "ms_per_forward": round(v[0] / 10, 4)} for k, v in enc.profile_read(reset=True).items() if v[1]}
            enc.profile(False)
        fl = cfg.flops_per_chunk() * B
        out[B] = {"eager_ms": statistics.median(ts), "graph_ms": statistics.median(tg),
                  "graph_tflops": fl / statistics.median(tg) / 1e9,
                  "utterances_per_s_graph": B / statistics.median(tg) * 1e3}
        if stages:
            out[B]["stages_eager"] = stages
        print(B, out[B], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "latency_small_batch.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
