"""Decode hand-off: the decoder the reference already uses consumes the B200 encoder's output on
the GPU.  CPU part: the transformers contract the hand-off relies on (generate(encoder_outputs=...) skips the encoder
and equals generate(input_features=...)).  GPU part: same, with the hidden states coming from ttasr."""
import numpy as np
import pytest

from oracle import encoder as OE
from oracle import frontend as OF


def _tiny_whisper(seed=0):
    import torch
    from transformers import WhisperConfig, WhisperForConditionalGeneration

    arch = OE.ARCHS["micro"]
    torch.manual_seed(seed)
    cfg = WhisperConfig(vocab_size=200, d_model=arch.d_model, encoder_layers=arch.layers,
                        encoder_attention_heads=arch.heads, encoder_ffn_dim=arch.ffn, decoder_layers=2,
                        decoder_attention_heads=arch.heads, decoder_ffn_dim=256, num_mel_bins=arch.n_mels,
                        max_source_positions=1500, max_target_positions=64, pad_token_id=0, bos_token_id=1,
                        eos_token_id=2, decoder_start_token_id=3, suppress_tokens=None, begin_suppress_tokens=None)
    model = WhisperForConditionalGeneration(cfg).eval()
    # encoder weights rounded to bf16 (what the B200 encoder holds), so both encoders see the same parameters
    with torch.no_grad():
        for p in model.model.encoder.parameters():
            p.copy_(p.to(torch.bfloat16).float())
    return arch, model


def _features(arch):
    clips = np.stack([OF.pad_or_trim(OF.synth_noise()), OF.pad_or_trim(OF.synth_tones())])
    return np.stack([OF.log_mel(c, arch.n_mels) for c in clips]), clips


def test_hf_generate_contract_cpu():
    import torch
    from ttasr.decode_handoff import encoder_outputs

    arch, model = _tiny_whisper()
    feats, _ = _features(arch)
    feats = torch.from_numpy(feats)
    with torch.no_grad():
        want = model.generate(input_features=feats, max_new_tokens=12, do_sample=False)
        hidden = model.get_encoder()(feats).last_hidden_state
        got = model.generate(encoder_outputs=encoder_outputs(hidden), max_new_tokens=12, do_sample=False)
    assert torch.equal(want, got)


@pytest.mark.gpu
def test_hf_decoder_consumes_b200_encoder_output(cuda_device):
    import torch
    import ttasr
    from ttasr.decode_handoff import hf_generate

    arch, model = _tiny_whisper()
    model = model.to(cuda_device)
    feats, clips = _features(arch)
    sd = {k: v.detach().float().cpu() for k, v in model.model.encoder.state_dict().items()}
    enc = ttasr.B200WhisperEncoder(
        dict(d_model=arch.d_model, encoder_layers=arch.layers, encoder_attention_heads=arch.heads,
             encoder_ffn_dim=arch.ffn, num_mel_bins=arch.n_mels), sd)
    fe = ttasr.B200WhisperFeatureExtractor(feature_size=arch.n_mels)
    pipe = ttasr.B200LogMelEncoder(fe, enc)
    pcm = torch.from_numpy(clips).to(cuda_device)
    with torch.no_grad():
        ref_hidden = model.get_encoder()(torch.from_numpy(feats).to(cuda_device)).last_hidden_state
        hidden = pipe.encode_device(pcm, out_dtype=torch.float32)
        stats = OE.parity_stats(hidden.cpu(), ref_hidden.cpu())
        assert stats["max_abs"] <= 0.10 and stats["cosine"] >= 0.9999, stats
        # decoder logits for a fixed prefix: the hand-off must not change what the decoder sees beyond bf16 noise
        prefix = torch.tensor([[3, 5, 7, 11]] * 2, device=cuda_device)
        lg_ref = model(encoder_outputs=(ref_hidden,), decoder_input_ids=prefix).logits
        lg = model(encoder_outputs=(hidden,), decoder_input_ids=prefix).logits
        assert (lg - lg_ref).abs().max().item() <= 0.05 * lg_ref.abs().max().item()
        want = model.generate(input_features=torch.from_numpy(feats).to(cuda_device), max_new_tokens=12, do_sample=False)
        got = hf_generate(model, pipe, pcm=pcm, max_new_tokens=12, do_sample=False)
    assert got.shape == want.shape and got.is_cuda
    assert (got == want).float().mean().item() >= 0.9, (got, want)  # greedy ties may flip a token on random weights
