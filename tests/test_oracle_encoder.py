"""The fp32 encoder oracle, pinned against golden vectors generated from the Hugging Face WhisperEncoder
(tests/golden/encoder_hf.npz), a live re-check when `transformers` is importable, and structural known answers."""
import os

import numpy as np
import pytest
import torch

from oracle import encoder as OE
from oracle import frontend as OF

GOLD = os.path.join(os.path.dirname(__file__), "golden", "encoder_hf.npz")


def _setup(name):
    arch = OE.ARCHS[name]
    w = OE.round_weights_bf16(OE.init_weights(arch, seed=0, ln_jitter=0.02))
    feats = torch.from_numpy(OF.log_mel(OF.synth_noise(), arch.n_mels))[None]
    return arch, w, feats


@pytest.mark.parametrize("name", ["micro", "tiny"])
def test_oracle_reproduces_golden_vectors(name):
    arch, w, feats = _setup(name)
    out = OE.encoder_forward(feats, w, arch)[0].numpy()
    g = np.load(GOLD)
    assert out.shape == (1500, arch.d_model)
    assert np.abs(out[::25] - g[f"enc_{name}_sub"]).max() < 2e-5
    assert abs(out.astype(np.float64).sum() - g[f"enc_{name}_stats"][0]) < 1e-4 * out.size


def test_oracle_matches_transformers_live():
    pytest.importorskip("transformers")
    from oracle.gen_golden import hf_encoder

    arch, w, feats = _setup("micro")
    with torch.no_grad():
        ref = hf_encoder(arch, w)(feats).last_hidden_state
    s = OE.parity_stats(OE.encoder_forward(feats, w, arch), ref)
    assert s["max_abs"] < 1e-5 and s["cosine"] > 0.999999


def test_k6_sinusoid_table():
    t = OE.sinusoids(1500, 384)
    assert t.shape == (1500, 384)
    assert torch.all(t[0, :192] == 0) and torch.all(t[0, 192:] == 1)
    assert torch.allclose(t[7, 0], torch.sin(torch.tensor(7.0))) and torch.allclose(t[7, 192], torch.cos(torch.tensor(7.0)))
    assert torch.allclose(t[1, 191], torch.tensor(1e-4), rtol=1e-4)  # slowest channel: sin(1/10000)


def test_k7_zero_weights_reduce_to_layernorm_of_positions():
    arch = OE.ARCHS["micro"]
    w = {k: torch.zeros_like(v) for k, v in OE.init_weights(arch).items()}
    w["embed_positions.weight"] = OE.sinusoids(1500, arch.d_model)
    w["layer_norm.weight"] = torch.ones(arch.d_model)
    out = OE.encoder_forward(torch.randn(1, arch.n_mels, 3000), w, arch)
    ref = torch.nn.functional.layer_norm(w["embed_positions.weight"], (arch.d_model,))
    assert torch.allclose(out[0], ref, atol=1e-6)


def test_wrong_length_raises_like_the_reference():
    arch, w, feats = _setup("micro")
    with pytest.raises(ValueError):
        OE.encoder_forward(feats[..., :2999], w, arch)


def test_flops_formula_matches_survey_table():
    assert OE.ARCHS["tiny"].flops_per_chunk() == 36_937_728_000
    assert OE.ARCHS["small"].flops_per_chunk() == 344_162_304_000
    assert OE.ARCHS["large-v3"].flops_per_chunk() == 2_273_771_520_000
