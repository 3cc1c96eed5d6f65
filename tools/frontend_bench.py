"""Front-end timing on one B200: ttasr_frontend_run (memset + frames kernel + conditional clamp kernel) on B synthetic
30 s chunks, CUDA events, median of 20 after 5 warm-ups; noise input (nothing clamped) and short clips (padding tiles).
python tools/frontend_bench.py [B] [n_mels]"""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "taiwan-tongues-asr-ce_b200"))
import torch  # noqa: E402

import ttasr  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    n_mels = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    dev = torch.device("cuda", 0)
    fe = ttasr.B200WhisperFeatureExtractor(feature_size=n_mels)
    g = torch.Generator(device=dev).manual_seed(1234)
    pcm = (0.1 * torch.randn((B, 480000), device=dev, generator=g)).clamp_(-1, 1)
    short = pcm.clone()
    short[:, 48000:] = 0  # 3 s of audio, 27 s of padding
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6555.2}
    out = {}
    nv = torch.full((B,), 48000, dtype=torch.int32, device=dev)
    for name, x, with_bf16, n_valid in (("noise f32+bf16", pcm, True, None), ("noise f32 only", pcm, False, None),
                                        ("3 s + zeros in data", short, True, None), ("3 s + n_valid", short, True, nv)):
        ts = []
        for i in range(25):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fe.extract(x, n_valid=n_valid, return_time_major=with_bf16)
            b.record()
            torch.cuda.synchronize()
            if i >= 5:
                ts.append(a.elapsed_time(b))
        med = statistics.median(ts)
        alg = B * (480000 * 4 + n_mels * 3000 * 4)
        out[name] = {"ms": med, "us_per_chunk": med * 1e3 / B, "alg_GBps": alg / med / 1e6, "frac_hbm": alg / med / 1e6 / peaks["hbm_gbs"]}
        print(f"{name:22s} B={B} n_mels={n_mels}: {med:.3f} ms  ({med * 1e3 / B:.2f} us/chunk, {alg / med / 1e6:.0f} GB/s algorithmic = "
              f"{alg / med / 1e6 / peaks['hbm_gbs']:.3f} of measured HBM peak)")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "frontend_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
