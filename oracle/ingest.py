"""TEST INFRASTRUCTURE — CPU oracle of the ingest step.  Only tests/, smoke() and the
cpu_baseline leg of bench.py may import this module; the product path never does.

What it restates: `librosa.load(path, sr=16000, mono=True)` as called by the reference (asr_core.py:156 with
mono=False + faster-whisper's own handling, api/file_asr.py:271-275 with mono=True), minus the file decoding.
librosa is a third-party dependency that is NOT on disk (requirements.txt:3 `librosa>=0.9.0`,
api/requirements.txt:16 `librosa>=0.10.0`; no lockfile), so its published algorithm is restated:

  * soundfile decodes PCM_16 to float32 as sample / 32768                       (librosa/core/audio.py `__soundfile_load`)
  * `to_mono(y)` = `np.mean(y, axis=0)` over the channel axis                   (librosa/core/audio.py `to_mono`)
  * `resample(y, orig_sr, target_sr, res_type)`; for res_type="polyphase":
        gcd = np.gcd(orig_sr, target_sr)
        y_hat = scipy.signal.resample_poly(y, target_sr // gcd, orig_sr // gcd, axis=-1)
    followed by `util.fix_length(y_hat, size=ceil(n * target_sr / orig_sr))`   (librosa/core/audio.py `resample`)

Parity pin: scipy.signal.resample_poly is installed here (scipy 1.18), so the resampler is checked against the very
function librosa calls.  librosa's DEFAULT res_type is "soxr_hq" (libsoxr, not installable offline): that filter
design differs from the polyphase one, so parity with an unmodified `librosa.load` call is *unpinned*; the reference
call sites can be switched with `res_type="polyphase"`."""
from __future__ import annotations

import math

import numpy as np

N_SAMPLES = 480000


def to_float_mono(frames: np.ndarray) -> np.ndarray:
    """frames [n] or [n, channels], int16 or float32 -> float32 [n]."""
    x = np.asarray(frames)
    if x.dtype == np.int16:
        x = x.astype(np.float32) / np.float32(32768.0)
    x = x.astype(np.float32)
    if x.ndim == 2:
        x = np.mean(x.T, axis=0)   # librosa keeps [channels, n]; to_mono averages axis 0
    return np.ascontiguousarray(x, dtype=np.float32)


def resample_polyphase(y: np.ndarray, orig_sr: int, target_sr: int = 16000) -> np.ndarray:
    import scipy.signal

    if orig_sr == target_sr:
        return y.copy()
    g = math.gcd(int(orig_sr), int(target_sr))
    y_hat = scipy.signal.resample_poly(y, target_sr // g, orig_sr // g, axis=-1)
    n = int(math.ceil(y.shape[-1] * target_sr / orig_sr))
    if y_hat.shape[-1] > n:
        y_hat = y_hat[..., :n]
    elif y_hat.shape[-1] < n:
        y_hat = np.pad(y_hat, (0, n - y_hat.shape[-1]))
    return np.ascontiguousarray(y_hat, dtype=np.float32)


def load_like_librosa(frames: np.ndarray, orig_sr: int, target_sr: int = 16000) -> np.ndarray:
    return resample_polyphase(to_float_mono(frames), orig_sr, target_sr)


def chunk(y: np.ndarray, chunk_samples: int = N_SAMPLES):
    """Independent zero-padded 30 s rows + real samples per row."""
    n_chunks = max(1, -(-len(y) // chunk_samples))
    out = np.zeros((n_chunks, chunk_samples), np.float32)
    out.reshape(-1)[: len(y)] = y
    n_valid = np.clip(len(y) - np.arange(n_chunks) * chunk_samples, 0, chunk_samples).astype(np.int32)
    return out, n_valid


def synth_frames(orig_sr: int, seconds: float, channels: int, dtype, seed: int = 0) -> np.ndarray:
    """Deterministic test signal: a few sines + noise, different per channel."""
    rng = np.random.default_rng(seed)
    n = int(round(orig_sr * seconds))
    t = np.arange(n) / orig_sr
    chans = []
    for c in range(channels):
        x = 0.2 * np.sin(2 * np.pi * (220.0 + 110.0 * c) * t) + 0.1 * np.sin(2 * np.pi * 3000.0 * t + c)
        x += 0.05 * np.sin(2 * np.pi * 7300.0 * t) + 0.02 * rng.standard_normal(n)
        chans.append(x)
    x = np.stack(chans, axis=1) if channels > 1 else chans[0]
    if dtype == np.int16:
        return np.clip(np.round(x * 32767.0), -32768, 32767).astype(np.int16)
    return x.astype(np.float32)
