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
    """faster-whisper `FeatureExtractor.__call__` semantics.

    Assumed upstream version: faster-whisper >= 1.1.0, `__call__(waveform, padding=160, chunk_length=None)` (numpy STFT,
    160 zero samples appended, features of the last window padded in FEATURE space by `pad_or_trim`).  Releases up to
    1.0.3 had `__call__(waveform, padding=True, chunk_length=None)` and appended `n_samples` (30 s) of zero AUDIO: pass
    `padding=True` (or `padding=self.n_samples`) for that scheme.  Any integer `padding` >= 0 is reproduced exactly.
    The reference pins only `faster-whisper>=0.9.0` (requirements.txt:10), so which scheme it ran depends on the install
    date; the package is not installable offline, hence parity with the real class is unpinned —
    tests/test_gpu_encoder.py::test_faster_whisper_seams checks this against the oracle's restatement of both schemes,
    and tests/test_faster_whisper_live.py against the real class whenever `import faster_whisper` works (e.g. from
    baseline/_ref)."""

    def __init__(self, fe: B200WhisperFeatureExtractor):
        self.fe = fe
        self.sampling_rate = fe.sampling_rate
        self.hop_length = fe.hop_length
        self.n_fft = fe.n_fft
        self.chunk_length = fe.chunk_length
        self.n_samples = fe.n_samples
        self.nb_max_frames = fe.nb_max_frames
        self.time_per_frame = fe.hop_length / fe.sampling_rate

    def __call__(self, waveform, padding=160, chunk_length=None):
        import torch

        if chunk_length is not None and int(chunk_length) != int(self.chunk_length):
            # upstream resizes n_samples / nb_max_frames here; the encoder behind this object takes 3000-frame windows
            raise NotImplementedError(f"chunk_length={chunk_length}: only the model's own {self.chunk_length} s window "
                                      "is implemented by the B200 encoder")
        x = np.asarray(waveform, dtype=np.float32).reshape(-1)
        pad = self.n_samples if padding is True else int(padding or 0)
        if pad < 0:
            raise ValueError("padding must be >= 0")
        if pad:
            x = np.concatenate([x, np.zeros(pad, np.float32)])
        hop, N = self.hop_length, self.n_samples
        n_frames = x.shape[0] // hop  # whole-file STFT frames minus the dropped last one
        if n_frames == 0:
            return np.zeros((self.fe.feature_size, 0), np.float32)
        # The whole-file STFT reflects 200 samples at the file END; the chunk kernel only knows "zeros after n_valid".
        # The kept frames reach at most 40 samples into that reflection, so materialise it: with padding >= 40 it is
        # all zeros anyway (the default scheme), with less it carries real audio.
        half = self.n_fft // 2
        if x.shape[0] > 1:
            refl = x[-2: -2 - half: -1] if x.shape[0] > half + 1 else np.resize(x[-2::-1], half)
            x = np.concatenate([x, refl.astype(np.float32)])
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
        # unclamped (log10 + 4) / 4 per chunk (the zero tail of the last row is declared as padding so its tiles skip the
        # FFT); the clamp is then taken over the frames of the FILE, as faster-whisper does — not per 30 s chunk, and
        # not over the helper frames of the rows (leads, reflection) that are cut away here
        feats = self.fe.extract(torch.from_numpy(rows).to(dev), n_valid=torch.from_numpy(lens).to(dev),
                                clamp_decades=float("inf"))
        full = torch.cat([feats[c, :, leads[c]: leads[c] + F] for c in range(n_chunks)], dim=1)[:, :n_frames]
        full = torch.maximum(full, full.max() - 2.0)              # 8 decades of log10 after the (x + 4) / 4 map
        return full.cpu().numpy()


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
