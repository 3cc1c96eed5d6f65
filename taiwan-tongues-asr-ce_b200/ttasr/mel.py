"""Host-side constants of the Whisper front end: slaney mel filter bank and periodic Hann window.

Same construction as the extractor the reference loads with AutoFeatureExtractor (train_asr.py:518-527):
transformers/audio_utils.py:263-332 (slaney mel scale), :356-375 + :453-544 (area-normalised triangles over
linspace(0, sr/2, n_fft/2+1)), :560-620 (np.hanning(n+1)[:-1]).  Computed once in fp64 and handed to
ttasr_frontend_create as fp32.
"""
from __future__ import annotations

import numpy as np

_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = 15.0
_LOGSTEP = np.log(6.4) / 27.0


def _hz_to_mel(f: np.ndarray) -> np.ndarray:
    f = np.atleast_1d(np.asarray(f, dtype=np.float64))
    m = 3.0 * f / 200.0
    hi = f >= _MIN_LOG_HZ
    m[hi] = _MIN_LOG_MEL + np.log(f[hi] / _MIN_LOG_HZ) / _LOGSTEP
    return m


def _mel_to_hz(m: np.ndarray) -> np.ndarray:
    m = np.atleast_1d(np.asarray(m, dtype=np.float64))
    f = 200.0 * m / 3.0
    hi = m >= _MIN_LOG_MEL
    f[hi] = _MIN_LOG_HZ * np.exp(_LOGSTEP * (m[hi] - _MIN_LOG_MEL))
    return f


def slaney_mel_filters(n_mels: int, n_fft: int = 400, sampling_rate: int = 16000, f_min: float = 0.0,
                       f_max: float = 8000.0) -> np.ndarray:
    """[n_fft//2 + 1, n_mels] fp64."""
    n_freq = n_fft // 2 + 1
    edges = _mel_to_hz(np.linspace(_hz_to_mel(f_min)[0], _hz_to_mel(f_max)[0], n_mels + 2))
    bins = np.linspace(0, sampling_rate // 2, n_freq)
    width = np.diff(edges)
    rel = edges[None, :] - bins[:, None]
    falling = -rel[:, :-2] / width[:-1]
    rising = rel[:, 2:] / width[1:]
    fb = np.clip(np.minimum(falling, rising), 0.0, None)
    return fb * (2.0 / (edges[2:] - edges[:-2]))[None, :]


def periodic_hann(n: int = 400) -> np.ndarray:
    return np.hanning(n + 1)[:-1].astype(np.float64)
