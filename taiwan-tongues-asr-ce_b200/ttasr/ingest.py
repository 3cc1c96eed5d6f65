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


def read_wav_pcm16(path: str):
    """Decode a PCM-16 WAV with the standard library -> (frames int16 [n, channels], sampling rate).  File decoding is
    host work on both sides of the boundary (the reference goes through soundfile / audioread inside librosa)."""
    import wave

    with wave.open(path, "rb") as w:
        if w.getsampwidth() != 2 or w.getcomptype() != "NONE":
            raise ValueError(f"{path}: only uncompressed 16-bit PCM WAV is decoded here")
        sr, ch, n = w.getframerate(), w.getnchannels(), w.getnframes()
        frames = np.frombuffer(w.readframes(n), dtype="<i2").reshape(-1, ch).copy()
    return frames, sr


def pause_aligned_cuts(db: np.ndarray, hop_s: float = 0.01, max_len_s: float = 30.0, min_silence_s: float = 0.2,
                       silence_db: float | None = None, search_back_s: float = 10.0):
    """Chunk boundaries for a long recording such that no chunk exceeds `max_len_s` and cuts fall inside pauses.

    db: per-hop energy track in dB (ttasr_ingest_frame_energy).  A pause is a run of >= min_silence_s hops below
    `silence_db` (default: 35 dB under the 95th percentile of the track).  Greedy, left to right: a chunk starting at s
    ends at the centre of the LAST pause whose centre lies in (s + max_len - search_back, s + max_len]; without one it
    ends at s + max_len (a hard cut, what `ttasr` round 1 always did).  Chunks tile the recording — nothing is dropped,
    so a chunk's timestamps are its start offset plus the decoder's.
    Returns a list of (start_hop, end_hop), end exclusive.

    Stand-in for the Silero VAD that faster-whisper runs behind `vad_filter=True` (asr_core.py:159-167,
    api/file_asr.py:457-465); that network is not reproduced here — callers who have its speech timestamps pass their
    own `cut_points` to `B200AudioIngest.load`."""
    db = np.asarray(db, dtype=np.float32).reshape(-1)
    n = int(db.shape[0])
    if n == 0:
        return []
    max_h = max(1, int(round(max_len_s / hop_s)))
    min_sil = max(1, int(round(min_silence_s / hop_s)))
    back = min(max_h - 1, int(round(search_back_s / hop_s)))
    thr = float(np.percentile(db, 95) - 35.0) if silence_db is None else float(silence_db)
    quiet = db < thr
    # centres of the pauses
    edges = np.flatnonzero(np.diff(np.concatenate([[0], quiet.view(np.int8), [0]])))
    starts, ends = edges[0::2], edges[1::2]
    keep = (ends - starts) >= min_sil
    centres = ((starts[keep] + ends[keep]) // 2).astype(np.int64)
    cuts, s = [], 0
    while s < n:
        hard = s + max_h
        if hard >= n:
            cuts.append((s, n))
            break
        lo = np.searchsorted(centres, hard - back, side="right")
        hi = np.searchsorted(centres, hard, side="right")
        e = int(centres[hi - 1]) if hi > lo else hard
        if e <= s:
            e = hard
        cuts.append((s, e))
        s = e
    return cuts


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

    def load_aligned(self, frames, cut_points=None, **vad_kwargs):
        """Long-form ingest with chunk boundaries inside pauses instead of every 480 000 samples.

        frames as for `load`.  cut_points: optional list of (start_sample, end_sample) at 16 kHz, each <= 30 s (e.g.
        derived from a VAD the host already runs); default = `pause_aligned_cuts` over the GPU-computed 10 ms energy
        track.  Returns (chunks [n, 480000] float32, n_valid int32 [n], starts int64 [n]): ragged rows for
        `extract(..., n_valid=...)`; `starts / 16000` are the chunk offsets to add to decoder timestamps."""
        import torch

        flat = self.load(frames, pad_to_chunks=False)
        dev = flat.device
        n = int(flat.shape[0])
        hop, win = self.target_sr // 100, self.target_sr // 50
        if cut_points is None:
            n_frames = max(1, -(-n // hop))
            db = torch.empty((n_frames,), dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                _lib.check(_lib.lib().ttasr_ingest_frame_energy(flat.data_ptr(), n, win, hop, db.data_ptr(), n_frames,
                                                                _lib.current_stream_ptr(dev)))
            cuts = pause_aligned_cuts(db.cpu().numpy(), hop_s=hop / self.target_sr,
                                      max_len_s=self.chunk_samples / self.target_sr, **vad_kwargs)
            cut_points = [(a * hop, min(b * hop, n)) for a, b in cuts]
        cut_points = [(int(a), int(b)) for a, b in cut_points]
        for a, b in cut_points:
            if not (0 <= a <= b <= n) or b - a > self.chunk_samples:
                raise ValueError(f"cut ({a}, {b}) is outside the signal or longer than {self.chunk_samples} samples")
        rows = max(1, len(cut_points))
        starts = torch.tensor([a for a, _ in cut_points] or [0], dtype=torch.int64)
        lens = torch.tensor([b - a for a, b in cut_points] or [0], dtype=torch.int32)
        starts_d, lens_d = starts.to(dev), lens.to(dev)
        out = torch.empty((rows, self.chunk_samples), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ttasr_ingest_gather_rows(flat.data_ptr(), starts_d.data_ptr(), lens_d.data_ptr(),
                                                           out.data_ptr(), rows, self.chunk_samples,
                                                           _lib.current_stream_ptr(dev)))
        return out, lens_d, starts

    def load(self, frames, pad_to_chunks: bool = True):
        """frames: CUDA tensor [n] or [n, channels] (interleaved, as decoded), int16 or float32.
        Returns (chunks, n_valid): the 16 kHz mono signal laid out as zero-padded 30 s rows plus the number of real
        samples per row — exactly the (pcm, n_valid) pair `B200WhisperFeatureExtractor.extract` takes.  With
        pad_to_chunks=False returns the flat [n_out] signal instead (librosa.load's return value).
        See `load_aligned` for chunk boundaries placed in pauses (long-form files)."""
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
