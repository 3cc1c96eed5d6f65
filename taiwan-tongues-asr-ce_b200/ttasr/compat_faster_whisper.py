"""Seams of `faster_whisper.WhisperModel` (REF asr_core.py:141,159-167; api/file_asr.py:188,457-465) re-pointed at
the B200 kernels.  faster-whisper computes the log-mel once over the WHOLE (VAD-concatenated) waveform — so the
`max - 8` clamp is per file, not per 30 s chunk — slices 3000-frame windows, and calls `self.encode(segment)` per
window.  Both behaviours are reproduced here on top of the chunked kernels:

  * `FileFeatureExtractor(fe)(waveform)` -> [n_mels, n_frames] float32 with the GLOBAL clamp: the per-chunk features
    are un-normalised back to raw log10, the file maximum is applied, then `(x + 4) / 4` again.  Chunks are formed
    with a 200-sample halo so frame values equal the whole-file STFT away from the two file ends.
  * `patch_model(model, encoder)` swaps `model.feature_extractor` and `model.encode`; `encode` returns whatever
    `wrap` makes of the CUDA tensor (identity by default; pass `ctranslate2.StorageView.from_array` when the CT2
    decoder consumes it).
"""
from __future__ import annotations

import numpy as np

from .encoder import B200WhisperEncoder
from .feature_extractor import B200WhisperFeatureExtractor


class FileFeatureExtractor:
    """faster-whisper `FeatureExtractor.__call__(waveform, padding=160, chunk_length=None)` semantics."""

    def __init__(self, fe: B200WhisperFeatureExtractor):
        self.fe = fe
        self.sampling_rate = fe.sampling_rate
        self.hop_length = fe.hop_length
        self.n_fft = fe.n_fft
        self.n_samples = fe.n_samples
        self.nb_max_frames = fe.nb_max_frames
        self.time_per_frame = fe.hop_length / fe.sampling_rate

    def __call__(self, waveform, padding: int = 160, chunk_length=None):
        import torch

        x = np.asarray(waveform, dtype=np.float32).reshape(-1)
        if padding:
            x = np.concatenate([x, np.zeros(padding, np.float32)])
        hop, N = self.hop_length, self.n_samples
        n_frames = x.shape[0] // hop  # whole-file STFT frames minus the dropped last one
        # A chunk row starting `lead` frames early reproduces the whole-file frames [c*F, (c+1)*F) at local indices
        # [lead, lead + F): local frames 0-1 (left reflect) and 2999 (right reflect) of an interior row are not file
        # frames, hence lead = 2 and F = 2997.  Row 0 starts at the file start, where the reflect IS the file's own.
        F = self.nb_max_frames - 3
        n_chunks = max(1, -(-n_frames // F))
        rows = np.zeros((n_chunks, N), np.float32)
        lens = np.zeros(n_chunks, np.int32)
        leads = [0] + [2] * (n_chunks - 1)
        for c in range(n_chunks):
            start = (c * F - leads[c]) * hop
            seg = x[start: start + N]
            rows[c, : seg.shape[0]] = seg
            lens[c] = seg.shape[0]
        dev = self.fe._torch_device()
        # per-chunk clamp, (x + 4) / 4; the zero tail of the last row is declared as padding so its tiles skip the FFT
        feats = self.fe.extract(torch.from_numpy(rows).to(dev), n_valid=torch.from_numpy(lens).to(dev))
        raw = torch.as_tensor(feats) * 4.0 - 4.0                  # back to (chunk-clamped) log10
        # the per-chunk clamp only raises values to (chunk max - 8) <= (file max - 8), so clamping again with the
        # file maximum gives exactly the whole-file result
        full = torch.cat([raw[c, :, leads[c]: leads[c] + F] for c in range(n_chunks)], dim=1)[:, :n_frames]
        full = torch.maximum(full, full.max() - 8.0)
        return ((full + 4.0) / 4.0).cpu().numpy()


def patch_model(model, encoder: B200WhisperEncoder, fe: B200WhisperFeatureExtractor | None = None, wrap=None):
    """Re-point `model.feature_extractor` and `model.encode` of a faster_whisper.WhisperModel-like object."""
    fe = fe or B200WhisperFeatureExtractor(feature_size=encoder.config.num_mel_bins)
    model.feature_extractor = FileFeatureExtractor(fe)

    def encode(features):
        f = np.asarray(features, dtype=np.float32)
        if f.ndim == 2:
            f = f[None]
        t_in = 2 * encoder.config.max_source_positions
        if f.shape[-1] < t_in:  # pad_or_trim in feature space, as faster-whisper does for the last window
            f = np.concatenate([f, np.zeros(f.shape[:-1] + (t_in - f.shape[-1],), np.float32)], axis=-1)
        hidden = encoder.encode(f[..., :t_in])
        return wrap(hidden) if wrap is not None else hidden

    model.encode = encode
    return model
