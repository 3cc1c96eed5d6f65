"""The numpy oracle of the log-mel front end, pinned three ways (no GPU needed):
golden vectors generated from the Hugging Face extractor (tests/golden/frontend_hf.npz, oracle/gen_golden.py),
a live re-check against `transformers` when importable, and the analytic cases K1-K5 of SURVEY.md section 8c."""
import os

import numpy as np
import pytest

from oracle import frontend as OF

GOLD = os.path.join(os.path.dirname(__file__), "golden", "frontend_hf.npz")
CLIPS = {"noise": OF.synth_noise, "tones": OF.synth_tones, "short": OF.synth_short}


@pytest.mark.parametrize("n_mels", [80, 128])
@pytest.mark.parametrize("name", list(CLIPS))
def test_oracle_reproduces_golden_vectors_bit_exact(n_mels, name):
    g = np.load(GOLD)
    out = OF.log_mel(CLIPS[name](), n_mels)
    key = f"logmel{n_mels}_{name}"
    assert out.shape == (n_mels, 3000) and out.dtype == np.float32
    assert np.array_equal(out[:, ::37], g[key + "_sub"])
    assert np.array_equal(out[:, :8], g[key + "_head"])
    assert np.array_equal(out[:, -8:], g[key + "_tail"])
    o64 = out.astype(np.float64)
    assert np.allclose([o64.sum(), o64.min(), o64.max()], g[key + "_stats"][:3], rtol=0, atol=1e-9)


@pytest.mark.parametrize("n_mels", [80, 128])
def test_oracle_matches_transformers_live(n_mels):
    tf = pytest.importorskip("transformers")
    fe = tf.WhisperFeatureExtractor(feature_size=n_mels)
    assert np.array_equal(fe.mel_filters, OF.mel_filter_bank(n_mels))
    rng = np.random.default_rng(42)
    x = (0.3 * rng.standard_normal(200000)).astype(np.float32)
    ref = fe._np_extract_fbank_features(OF.pad_or_trim(x)[None], "cpu")[0]
    assert np.array_equal(ref, OF.log_mel(x, n_mels))
    # K8: the torch path the HF __call__ dispatches to agrees with the numpy path to ~1e-5
    alt = fe(x, sampling_rate=16000, return_tensors="np")["input_features"][0]
    assert np.abs(alt - ref).max() < 5e-5


def test_k1_all_zero_pcm_is_minus_one_point_five():
    assert np.all(OF.log_mel(np.zeros(480000, np.float32), 80) == -1.5)
    assert np.all(OF.log_mel(np.zeros(0, np.float32), 128) == -1.5)


def test_k2_padded_tail_sits_at_the_clamp_floor():
    out = OF.log_mel(OF.synth_short(), 128)
    tail = out[:, 400:]
    assert np.all(tail == out.min()) and np.isclose(out.min(), (out.max() * 4 - 4 - 8 + 4) / 4, atol=1e-6)


def test_k3_shape_and_frame_centering():
    x = np.zeros(480000, np.float32)
    x[160 * 1000] = 1.0  # an impulse at the centre of frame 1000
    out = OF.log_mel(x, 80)
    energy = out.sum(axis=0)
    assert out.shape == (80, 3000) and energy.argmax() == 1000
    assert np.all(energy[:998] == energy[0]) and np.all(energy[1003:] == energy[0])


def test_k5_bin_centred_sine_closed_form():
    # amplitude-A sine on bin k: |X[k]| = A * sum(window) / 2 = 100 A; neighbours k +- 1 carry 50 A (Hann)
    k, A = 50, 0.5
    t = np.arange(480000) / 16000.0
    x = (A * np.sin(2 * np.pi * (k * 40.0) * t)).astype(np.float32)
    fb = OF.mel_filter_bank(80)
    power = np.zeros(201)
    power[k], power[k - 1], power[k + 1] = (100 * A) ** 2, (50 * A) ** 2, (50 * A) ** 2
    expect = (np.log10(fb.T @ power) + 4) / 4
    out = OF.log_mel(x, 80)
    m = int(np.argmax(fb[k]))
    assert abs(out[m, 1500] - expect[m]) < 2e-4 and out[:, 1500].argmax() == expect.argmax()


def test_padding_and_truncation_semantics():
    x = OF.synth_noise(9, 500000)
    assert np.array_equal(OF.log_mel(x, 80), OF.log_mel(x[:480000], 80))
    assert OF.pad_or_trim(x[:10]).shape == (480000,) and OF.pad_or_trim(x[:10])[10:].sum() == 0


def test_window_is_periodic_hann():
    w = OF.hann_window()
    assert w[0] == 0 and np.isclose(w[200], 1.0) and np.isclose(w.sum(), 200.0)
