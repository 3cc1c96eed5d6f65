"""Ingest step in front of the log-mel path: decoded PCM -> mono float32 -> 16 kHz ->
zero-padded 30 s chunks, on the GPU.

Replaces the numeric part of `librosa.load(path, sr=16000, mono=True)` at the reference's call sites
(asr_core.py:156; api/file_asr.py:271-275 which then forces a contiguous 1-D float32 array) and the independent
30 s chunking of SURVEY.md section 8d config 4.  Semantics: samples / 32768 for int16, `librosa.to_mono` (mean over
channels), `librosa.resample(..., res_type="polyphase")` = `scipy.signal.resample_poly(y, up, down)` with
up/down = 16000/orig_sr reduced by their gcd, output length ceil(n * 16000 / orig_sr).  File decoding (soundfile /
audioread inside librosa) stays on the host: this class takes the decoded frames.

The arithmetic runs in `ttasr_ingest_run` (csrc/ingest_resample.cu); there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib


def resample_poly_filter(up: int, down: int, beta: float = 5.0) -> np.ndarray:
    """The prototype low-pass scipy.signal.resample_poly designs (signal/_signaltools.py `resample_poly`, default
    window=("kaiser", 5.0)): firwin(2 * 10 * max(up, down) + 1, 1 / max(up, down)) * up, in float32 as scipy uses
    for float32 input.  numpy only (np.sinc, np.kaiser): firwin's windowed ideal low-pass scaled to unit DC gain."""
    max_rate = max(up, down)
    half_len = 10 * max_rate
    numtaps = 2 * half_len + 1
    f_c = 1.0 / max_rate
    m = np.arange(numtaps, dtype=np.float64) - half_len
    h = f_c * np.sinc(f_c * m) * np.kaiser(numtaps, beta)
    h /= h.sum()
    return (h.astype(np.float32) * np.float32(up)).astype(np.float32)


class B200AudioIngest:
    """`B200AudioIngest(orig_sr).load(frames)` -> (chunks [n_chunks, 480000] float32 CUDA, n_valid int32 [n_chunks])."""

    def __init__(self, orig_sr: int, target_sr: int = 16000, chunk_length: int = 30):
        if int(orig_sr) != orig_sr or orig_sr <= 0:
            raise ValueError(f"orig_sr must be a positive integer sampling rate, got {orig_sr}")
        self.orig_sr, self.target_sr = int(orig_sr), int(target_sr)
        g = math.gcd(self.orig_sr, self.target_sr)
        self.up, self.down = self.target_sr // g, self.orig_sr // g
        self.chunk_samples = chunk_length * self.target_sr
        self._handle = None

    def _native(self):
        if self._handle is None:
            h = C.c_void_p()
            if self.up == 1 and self.down == 1:
                _lib.check(_lib.lib().ttasr_ingest_create(1, 1, None, 0, C.byref(h)))
            else:
                taps = np.ascontiguousarray(resample_poly_filter(self.up, self.down))
                _lib.check(_lib.lib().ttasr_ingest_create(self.up, self.down, taps.ctypes.data, int(taps.size), C.byref(h)))
            self._handle = h
        return self._handle

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h is not None:
            try:
                _lib.lib().ttasr_ingest_destroy(h)
            except Exception:
                pass

    def out_len(self, n_in: int) -> int:
        x = int(n_in) * self.up
        return x // self.down + (x % self.down != 0)

    def load(self, frames, pad_to_chunks: bool = True):
        """frames: CUDA tensor [n] or [n, channels] (interleaved, as decoded), int16 or float32.
        Returns (chunks, n_valid): the 16 kHz mono signal laid out as zero-padded 30 s rows plus the number of real
        samples per row — exactly the (pcm, n_valid) pair `B200WhisperFeatureExtractor.extract` takes.  With
        pad_to_chunks=False returns the flat [n_out] signal instead (librosa.load's return value)."""
        import torch

        frames = _lib.require_cuda_tensor(frames, "frames")
        if frames.dim() == 1:
            frames = frames.unsqueeze(1)
        if frames.dim() != 2:
            raise ValueError("frames must be [n] or [n, channels]")
        if frames.dtype == torch.int16:
            dtype = _lib.PCM_I16
        elif frames.dtype == torch.float32:
            dtype = _lib.PCM_F32
        else:
            raise _lib.TtasrError(-1, f"frames dtype must be int16 or float32, got {frames.dtype}")
        frames = frames.contiguous()
        n_in, ch = int(frames.shape[0]), int(frames.shape[1])
        n_out = self.out_len(n_in)
        n_chunks = max(1, -(-n_out // self.chunk_samples)) if pad_to_chunks else 1
        cap = n_chunks * self.chunk_samples if pad_to_chunks else n_out
        with torch.cuda.device(frames.device):
            out = torch.empty((cap,), dtype=torch.float32, device=frames.device)
            _lib.check(_lib.lib().ttasr_ingest_run(self._native(), frames.data_ptr(), dtype, ch, n_in, out.data_ptr(),
                                                   cap, _lib.current_stream_ptr(frames.device)))
        if not pad_to_chunks:
            return out
        starts = torch.arange(n_chunks, dtype=torch.int64) * self.chunk_samples
        n_valid = torch.clamp(n_out - starts, min=0, max=self.chunk_samples).to(torch.int32).to(frames.device)
        return out.view(n_chunks, self.chunk_samples), n_valid
