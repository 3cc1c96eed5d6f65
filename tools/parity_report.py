#!/usr/bin/env python
"""Encoder parity report on the B200 (the numbers the frozen gates in tests/ are derived from).

For each architecture: this library in each residual mode (split = default, f32, bf16) and the Hugging Face encoder run
entirely in bf16 on the same GPU (the "HF-bf16 control" of SURVEY.md 8d), all against the fp32 oracle run on the GPU
with TF32 off, on the same bf16-rounded random-init weights and the oracle's fp32 log-mel features.

    python tools/parity_report.py [--archs tiny small large-v3] [--chunks 4] > profiles/r2_parity_report.json
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "taiwan-tongues-asr-ce_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def main():
    import numpy as np
    import torch

    import ttasr
    from oracle import encoder as OE
    from oracle import frontend as OF
    from oracle.gen_golden import hf_encoder

    ap = argparse.ArgumentParser()
    ap.add_argument("--archs", nargs="+", default=["tiny", "small", "large-v3"])
    ap.add_argument("--chunks", type=int, default=4)
    args = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")
    dev = torch.device("cuda", 0)
    report = {"device": torch.cuda.get_device_name(0), "chunks": args.chunks, "archs": {}}
    clips = [OF.synth_noise, OF.synth_tones, OF.synth_short, lambda: OF.synth_noise(77)]
    for name in args.archs:
        arch = OE.ARCHS[name]
        w = OE.round_weights_bf16(OE.init_weights(arch, seed=0, ln_jitter=0.02))
        feats = np.stack([OF.log_mel(OF.pad_or_trim(clips[i % len(clips)]()), arch.n_mels) for i in range(args.chunks)])
        fdev = torch.from_numpy(feats).to(dev)
        hf = hf_encoder(arch, w).to(dev)
        with torch.no_grad():
            ref = torch.cat([hf(fdev[i:i + 2]).last_hidden_state for i in range(0, args.chunks, 2)]).cpu()
            hf16 = hf.to(torch.bfloat16)
            ctl = torch.cat([hf16(fdev[i:i + 2].to(torch.bfloat16)).last_hidden_state
                             for i in range(0, args.chunks, 2)]).float().cpu()
        del hf, hf16
        torch.cuda.empty_cache()
        entry = {"hf_bf16_control": OE.parity_stats(ctl, ref)}
        cfg = dict(d_model=arch.d_model, encoder_layers=arch.layers, encoder_attention_heads=arch.heads,
                   encoder_ffn_dim=arch.ffn, num_mel_bins=arch.n_mels, max_source_positions=arch.n_ctx)
        for mode in ("split", "f32", "bf16"):
            enc = ttasr.B200WhisperEncoder(cfg, w, residual=mode)
            got32 = enc.encode(feats, out_dtype=torch.float32).cpu()
            got16 = enc.encode(feats).float().cpu()
            entry[f"ttasr_{mode}_f32out"] = OE.parity_stats(got32, ref)
            entry[f"ttasr_{mode}_bf16out"] = OE.parity_stats(got16, ref)
            del enc
            torch.cuda.empty_cache()
        report["archs"][name] = entry
        print(name, json.dumps(entry), file=sys.stderr, flush=True)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
