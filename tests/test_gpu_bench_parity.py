"""Parity AT THE BENCHMARKED CONFIGURATIONS (BASELINE.json configs[1] and configs[2]), through the same objects and
batch sizes bench.py times: sampled rows of the real 256 x large-v3 and 64 x small batches are compared with the fp32
oracle run ON THE B200 with TF32 off (SURVEY.md section 8c allows exactly this), so the 256-wide cta_group::2 GEMM
tiles, the full-occupancy attention schedule and the batched front end are tied to the oracle, not to themselves.

The oracle here is Hugging Face's own `WhisperEncoder` module (fp32, eager attention) when transformers imports, and
oracle/encoder.py (pinned to it by tests/test_oracle_encoder.py) otherwise; features come from the numpy oracle.
"""
import numpy as np
import pytest

from oracle import encoder as OE
from oracle import frontend as OF

pytestmark = pytest.mark.gpu

# frozen gates (round 2): ~2x the errors measured on the B200 for the default path with bf16 output
# (profiles/r2_parity_report.json: large-v3 0.029 / 0.0040 / 0.9999875, small 0.022 / 0.0030 / 0.9999930), per architecture
GATES = {
    "large-v3": dict(max_abs=0.07, mean_abs=0.008, cosine=0.999975),
    "small": dict(max_abs=0.055, mean_abs=0.006, cosine=0.999985),
}


def _oracle_on_gpu(arch, w32, feats, dev):
    """fp32 hidden states [n, 1500, d] for `feats` [n, n_mels, 3000] (CUDA tensor), TF32 disabled."""
    import torch

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")
    try:
        from oracle.gen_golden import hf_encoder

        enc = hf_encoder(arch, {k: v.cpu() for k, v in w32.items()}).to(dev)
        enc.config._attn_implementation = "eager"
        with torch.no_grad():
            return torch.cat([enc(feats[i:i + 2]).last_hidden_state for i in range(0, feats.shape[0], 2)]), "hf"
    except ImportError:
        wd = {k: v.to(dev) for k, v in w32.items()}
        return torch.cat([OE.encoder_forward(feats[i:i + 2], wd, arch) for i in range(0, feats.shape[0], 2)]), "port"


def _run(workload, B, rows, dev):
    import torch
    import bench
    import ttasr

    cfg = ttasr.EncoderConfig.named(workload)
    arch = OE.ARCHS[workload]
    weights = bench.make_gpu_weights(cfg, dev)          # the bench's own random-init bf16 weights
    fe = ttasr.B200WhisperFeatureExtractor(feature_size=cfg.num_mel_bins)
    enc = ttasr.B200WhisperEncoder(cfg, weights)
    pipe = ttasr.B200LogMelEncoder(fe, enc)
    g = torch.Generator(device=dev).manual_seed(1234)   # bench.py's seed for rank 0
    pcm = (0.1 * torch.randn((B, 480000), generator=g, device=dev)).clamp_(-1, 1)
    hidden = pipe.encode_device(pcm)                    # the timed path: bf16 out, whole batch in one call
    assert hidden.shape == (B, 1500, cfg.d_model) and hidden.dtype == torch.bfloat16
    got = hidden[rows].float()
    feats_gpu = fe.extract(pcm[rows].contiguous())
    del hidden, pipe, enc
    torch.cuda.empty_cache()
    feats_ref = np.stack([OF.log_mel(pcm[r].cpu().numpy(), cfg.num_mel_bins) for r in rows])
    fe_err = float(np.abs(feats_gpu.cpu().numpy() - feats_ref).max())
    assert fe_err <= 1e-4, f"log-mel max abs err {fe_err}"
    w32 = {k: v.float() for k, v in weights.items()}
    ref, kind = _oracle_on_gpu(arch, w32, torch.from_numpy(feats_ref).to(dev), dev)
    return got, ref, kind, fe_err


def test_large_v3_b256_sampled_rows_vs_fp32_oracle(cuda_device):
    """configs[2]: whisper-large-v3, 128-bin, 256 x 30 s in one step; 8 sampled rows (first / last tile of the batch,
    both CTAs of a pair, both epilogue warpgroups) against the fp32 oracle on the same device."""
    rows = [0, 1, 37, 100, 128, 129, 200, 255]
    got, ref, kind, fe_err = _run("large-v3", 256, rows, cuda_device)
    s = OE.parity_stats(got.cpu(), ref.cpu())
    print(f"large-v3 x 256, rows {rows} vs {kind} fp32 on the B200: {s}; log-mel {fe_err:.2e}")
    gate = GATES["large-v3"]
    assert s["max_abs"] <= gate["max_abs"] and s["mean_abs"] <= gate["mean_abs"] and s["cosine"] >= gate["cosine"], s


def test_small_b64_sampled_rows_vs_fp32_oracle(cuda_device):
    """configs[1]: whisper-small, 80-bin, 64 x 30 s in one step; 8 sampled rows against the fp32 oracle."""
    rows = [0, 1, 13, 31, 32, 33, 50, 63]
    got, ref, kind, fe_err = _run("small", 64, rows, cuda_device)
    s = OE.parity_stats(got.cpu(), ref.cpu())
    print(f"small x 64, rows {rows} vs {kind} fp32 on the B200: {s}; log-mel {fe_err:.2e}")
    gate = GATES["small"]
    assert s["max_abs"] <= gate["max_abs"] and s["mean_abs"] <= gate["mean_abs"] and s["cosine"] >= gate["cosine"], s


def test_two_ranks_produce_identical_bits(cuda_device, tmp_path):
    """the same chunk gives the same bits on any rank.  Two processes (torchrun-style env, gloo control
    group, one GPU each when the box has two, else sharing cuda:0) encode the same probe batch with the same weights and
    compare SHA-256 digests of the hidden states."""
    import os
    import subprocess
    import sys

    import torch

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = os.path.join(root, "tools", "rank_probe.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29500 + os.getpid() % 2000), WORLD_SIZE="2")
    out = [str(tmp_path / f"r{r}.txt") for r in range(2)]
    procs = [subprocess.Popen([sys.executable, script, "--out", out[r]],
                              env=dict(env, RANK=str(r), LOCAL_RANK=str(r % max(torch.cuda.device_count(), 1))))
             for r in range(2)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    digests = [open(o).read().split() for o in out]
    assert digests[0][0] == digests[1][0] and len(digests[0][0]) == 64, digests
    assert digests[0][1] == "identical=True" and digests[1][1] == "identical=True"
