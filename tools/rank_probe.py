#!/usr/bin/env python
"""Cross-rank determinism probe: every rank encodes the SAME probe batch (seeded PCM, seeded weights)
on its own GPU and the SHA-256 digests of the hidden states are compared over the control group.

    torchrun --nproc-per-node N tools/rank_probe.py [--workload tiny] [--out FILE]

Prints / writes "<sha256> identical=<bool>"; exit code 1 when a rank disagrees.  bench.py --gpus N runs the same probe
(ttasr.dp.probe_digest) before its timed region and reports it as `rank_probe`."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "taiwan-tongues-asr-ce_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def main():
    import torch
    import torch.distributed as dist

    import bench
    import ttasr
    from ttasr.dp import probe_digest

    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="tiny")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        dist.init_process_group("gloo")
    cfg = ttasr.EncoderConfig.named(args.workload)
    pipe = ttasr.B200LogMelEncoder(ttasr.B200WhisperFeatureExtractor(feature_size=cfg.num_mel_bins),
                                   ttasr.B200WhisperEncoder(cfg, bench.make_gpu_weights(cfg, dev)))
    digest, same = probe_digest(pipe)
    line = f"{digest} identical={same}"
    print(line, flush=True)
    if args.out:
        with open(args.out, "w") as f:
            f.write(line + "\n")
    if dist.is_initialized():
        dist.destroy_process_group()
    return 0 if same else 1


if __name__ == "__main__":
    sys.exit(main())
