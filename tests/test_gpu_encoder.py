"""Encoder parity through the drop-in Python surface -> C ABI -> sm_100a kernels, against the fp32 oracle on the same
bf16-rounded weights.  Stated tolerance: max-abs <= 0.10, mean-abs <= 0.01, cosine >= 0.9999."""
import os

import numpy as np
import pytest

from oracle import encoder as OE
from oracle import frontend as OF

pytestmark = pytest.mark.gpu
MAX_ABS, MEAN_ABS, COS = 0.10, 0.01, 0.9999


def _cfg(arch):
    return dict(d_model=arch.d_model, encoder_layers=arch.layers, encoder_attention_heads=arch.heads,
                encoder_ffn_dim=arch.ffn, num_mel_bins=arch.n_mels, max_source_positions=arch.n_ctx)


def _build(arch_name, seed=0):
    from ttasr import B200WhisperEncoder

    arch = OE.ARCHS[arch_name]
    w = OE.round_weights_bf16(OE.init_weights(arch, seed=seed, ln_jitter=0.02))
    return arch, w, B200WhisperEncoder(_cfg(arch), w)


def _check(got, ref):
    s = OE.parity_stats(got, ref)
    assert s["max_abs"] <= MAX_ABS and s["mean_abs"] <= MEAN_ABS and s["cosine"] >= COS, s
    return s


@pytest.mark.parametrize("arch_name", ["micro", "tiny"])
def test_encoder_matches_fp32_oracle(cuda_device, arch_name):
    import torch

    arch, w, enc = _build(arch_name)
    feats = np.stack([OF.log_mel(OF.synth_noise(), arch.n_mels), OF.log_mel(OF.synth_tones(), arch.n_mels),
                      OF.log_mel(OF.synth_short(), arch.n_mels)])
    ref = OE.encoder_forward(torch.from_numpy(feats), w, arch).numpy()
    got = enc.encode(feats, out_dtype=torch.float32).cpu().numpy()
    assert got.shape == (3, 1500, arch.d_model)
    _check(got, ref)
    got16 = enc.encode(feats)
    assert got16.dtype == torch.bfloat16
    _check(got16.float().cpu().numpy(), ref)
    # same chunk, other batch position -> identical bits (SURVEY.md section 8e determinism)
    again = enc.encode(feats[[2, 0]], out_dtype=torch.float32).cpu().numpy()
    assert np.array_equal(again[1], got[0]) and np.array_equal(again[0], got[2])


def test_golden_fixture_tiny(cuda_device):
    import torch

    arch, w, enc = _build("tiny")
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "encoder_hf.npz"))
    feats = OF.log_mel(OF.synth_noise(), arch.n_mels)[None]
    got = enc.encode(feats, out_dtype=torch.float32)[0].cpu().numpy()
    _check(got[::25], g["enc_tiny_sub"])


def test_forward_is_a_drop_in_for_hf_encoder(cuda_device):
    """encoder(input_features) -> .last_hidden_state, usable as encoder_outputs by the reference's decode code."""
    import torch

    arch, w, enc = _build("micro")
    feats = torch.from_numpy(OF.log_mel(OF.synth_noise(), arch.n_mels)[None])
    out = enc(feats, attention_mask=None)
    assert out.last_hidden_state.shape == (1, 1500, arch.d_model) and out[0] is out.last_hidden_state
    with pytest.raises(ValueError):
        enc(feats[..., :2000])


def test_pcm_to_hidden_pipeline(cuda_device):
    import torch
    from ttasr import B200LogMelEncoder, B200WhisperFeatureExtractor

    arch, w, enc = _build("tiny")
    pipe = B200LogMelEncoder(B200WhisperFeatureExtractor(feature_size=arch.n_mels), enc)
    clips = np.stack([OF.pad_or_trim(OF.synth_noise(3)), OF.pad_or_trim(OF.synth_short(4))])
    got = pipe.encode_device(torch.from_numpy(clips).to(cuda_device), out_dtype=torch.float32).cpu().numpy()
    feats = np.stack([OF.log_mel(c, arch.n_mels) for c in clips])
    ref = OE.encoder_forward(torch.from_numpy(feats), w, arch).numpy()
    _check(got, ref)
    host = torch.from_numpy(clips).pin_memory()
    out_host = torch.empty((2, 1500, arch.d_model), dtype=torch.bfloat16).pin_memory()
    pipe.encode_host(host, out_host)
    torch.cuda.synchronize()
    _check(out_host.float().numpy(), ref)


@pytest.mark.parametrize("arch_name", ["small"])
def test_encoder_small_config2_sample(cuda_device, arch_name):
    """Config 2 architecture (whisper-small, 80-bin) on two chunks; the oracle needs ~2 s per chunk on CPU."""
    import torch

    arch, w, enc = _build(arch_name)
    feats = np.stack([OF.log_mel(OF.synth_noise(11), arch.n_mels), OF.log_mel(OF.synth_tones(12), arch.n_mels)])
    ref = OE.encoder_forward(torch.from_numpy(feats), w, arch).numpy()
    got = enc.encode(feats, out_dtype=torch.float32).cpu().numpy()
    _check(got, ref)
