"""numpy restatement of the Whisper log-mel front end (TEST INFRASTRUCTURE — see oracle/__init__.py).

Follows the numpy path of the Hugging Face extractor the reference calls at
`train_asr.py:607-616` (HF = site-packages/transformers, v5.5.0):

* constants            HF/models/whisper/feature_extraction_whisper.py:69-103
* mel scale (slaney)   HF/audio_utils.py:263-332
* triangular filters   HF/audio_utils.py:356-375, :453-544
* periodic Hann        HF/audio_utils.py:560-620
* STFT / power / mel   HF/audio_utils.py:624-832   (centre reflect pad :769-771, fp64 :774-775,
                       complex64 spectrum :781,803, |.|**2 in fp64 :808, mel floor 1e-10 :813, log10 :819)
* drop frame, clamp    HF/models/whisper/feature_extraction_whisper.py:128-130

Arithmetic is fp64 between an fp32 input and an fp32 output, exactly as the reference does it;
only the Python per-frame loop is replaced by one batched `np.fft.rfft` over a strided view.
"""
from __future__ import annotations

import numpy as np

SAMPLING_RATE = 16000
N_FFT = 400
HOP = 160
CHUNK_SECONDS = 30
N_SAMPLES = CHUNK_SECONDS * SAMPLING_RATE  # 480000
N_FRAMES = N_SAMPLES // HOP  # 3000
N_FREQ = N_FFT // 2 + 1  # 201
MEL_FLOOR = 1e-10


def hertz_to_mel_slaney(freq):
    """HF/audio_utils.py:263-296 (mel_scale="slaney")."""
    freq = np.asarray(freq, dtype=np.float64)
    min_log_hertz, min_log_mel = 1000.0, 15.0
    logstep = 27.0 / np.log(6.4)
    mels = 3.0 * freq / 200.0
    log_region = freq >= min_log_hertz
    if np.ndim(mels) == 0:
        return float(min_log_mel + np.log(freq / min_log_hertz) * logstep) if log_region else float(mels)
    mels = mels.copy()
    mels[log_region] = min_log_mel + np.log(freq[log_region] / min_log_hertz) * logstep
    return mels


def mel_to_hertz_slaney(mels):
    """HF/audio_utils.py:299-332 (mel_scale="slaney")."""
    mels = np.asarray(mels, dtype=np.float64)
    min_log_hertz, min_log_mel = 1000.0, 15.0
    logstep = np.log(6.4) / 27.0
    freq = 200.0 * mels / 3.0
    log_region = mels >= min_log_mel
    freq = freq.copy()
    freq[log_region] = min_log_hertz * np.exp(logstep * (mels[log_region] - min_log_mel))
    return freq


def mel_filter_bank(n_mels: int, n_freq: int = N_FREQ, sampling_rate: int = SAMPLING_RATE,
                    f_min: float = 0.0, f_max: float = 8000.0) -> np.ndarray:
    """[n_freq, n_mels] fp64, slaney scale + slaney area norm. HF/audio_utils.py:453-544, :356-375."""
    mel_min = hertz_to_mel_slaney(f_min)
    mel_max = hertz_to_mel_slaney(f_max)
    mel_freqs = np.linspace(mel_min, mel_max, n_mels + 2)
    filter_freqs = mel_to_hertz_slaney(mel_freqs)
    fft_freqs = np.linspace(0, sampling_rate // 2, n_freq)
    filter_diff = np.diff(filter_freqs)
    slopes = np.expand_dims(filter_freqs, 0) - np.expand_dims(fft_freqs, 1)
    down = -slopes[:, :-2] / filter_diff[:-1]
    up = slopes[:, 2:] / filter_diff[1:]
    fb = np.maximum(np.zeros(1), np.minimum(down, up))
    enorm = 2.0 / (filter_freqs[2:n_mels + 2] - filter_freqs[:n_mels])
    fb = fb * np.expand_dims(enorm, 0)
    return fb


def hann_window(n: int = N_FFT) -> np.ndarray:
    """Periodic Hann, fp64: np.hanning(n+1)[:-1]. HF/audio_utils.py:597-608."""
    return np.hanning(n + 1)[:-1].astype(np.float64)


def pad_or_trim(pcm: np.ndarray, n_samples: int = N_SAMPLES) -> np.ndarray:
    """Right-pad with 0.0 / truncate to n_samples (HF feature_extraction_sequence_utils pad/_truncate,
    as called from feature_extraction_whisper.py:296-303)."""
    pcm = np.asarray(pcm, dtype=np.float32).reshape(-1)
    if pcm.shape[0] >= n_samples:
        return pcm[:n_samples].copy()
    out = np.zeros(n_samples, dtype=np.float32)
    out[: pcm.shape[0]] = pcm
    return out


def log_mel_unclamped(pcm: np.ndarray, n_mels: int) -> np.ndarray:
    """log10(max(1e-10, mel power)) for all 1 + len/160 frames, fp32 [n_mels, n_frames+1]."""
    pcm = np.asarray(pcm, dtype=np.float32).reshape(-1)
    wave = np.pad(pcm, (N_FFT // 2, N_FFT // 2), mode="reflect").astype(np.float64)
    n_frames = 1 + (wave.size - N_FFT) // HOP
    frames = np.lib.stride_tricks.as_strided(
        wave, shape=(n_frames, N_FFT), strides=(wave.strides[0] * HOP, wave.strides[0]), writeable=False
    )
    spec = np.fft.rfft(frames * hann_window()[None, :], axis=-1).astype(np.complex64)
    power = np.abs(spec, dtype=np.float64) ** 2.0  # [frames, 201]
    mel = np.maximum(MEL_FLOOR, np.dot(mel_filter_bank(n_mels).T, power.T))
    return np.asarray(np.log10(mel), dtype=np.float32)


def log_mel(pcm: np.ndarray, n_mels: int) -> np.ndarray:
    """One padded/truncated 30 s chunk -> [n_mels, 3000] fp32 (feature_extraction_whisper.py:105-133)."""
    x = log_mel_unclamped(pad_or_trim(pcm), n_mels)[:, :-1]
    x = np.maximum(x, x.max() - 8.0)
    x = (x + 4.0) / 4.0
    return x.astype(np.float32)


def log_mel_batch(pcm_batch, n_mels: int) -> np.ndarray:
    """[B, *] -> [B, n_mels, 3000]; the clamp max is per row (per 30 s chunk)."""
    return np.stack([log_mel(p, n_mels) for p in pcm_batch], axis=0)


# ----------------------------------------------------------------------------------------------
# Synthetic inputs shared by tests / bench
# ----------------------------------------------------------------------------------------------
def synth_noise(seed: int = 0, n: int = N_SAMPLES, sigma: float = 0.1) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return (sigma * rng.standard_normal(n)).astype(np.float32)


def synth_tones(seed: int = 1, n: int = N_SAMPLES) -> np.ndarray:
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / SAMPLING_RATE
    x = sum(0.1 * np.sin(2 * np.pi * f * t) for f in (220.0, 440.0, 1000.0, 3000.0, 7000.0))
    x = x + 0.01 * rng.standard_normal(n)
    return x.astype(np.float32)


def synth_short(seed: int = 2, seconds: float = 3.0) -> np.ndarray:
    return synth_noise(seed, int(seconds * SAMPLING_RATE))
