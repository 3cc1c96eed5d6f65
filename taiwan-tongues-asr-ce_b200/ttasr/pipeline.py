"""The hot path end to end: 16 kHz PCM chunks -> log-mel -> Whisper encoder hidden states, one object per GPU.

This is what the reference's three inference call sites do per 30 s window inside `WhisperModel.transcribe`
(asr_core.py:159-167, api/file_asr.py:457-465, api/stt_streaming/src/asr/faster_whisper_asr.py:170-172), batched:
features never leave the GPU between the two stages (the front end writes the bf16 time-major tensor the conv stem
reads through TMA).
"""
from __future__ import annotations

from . import _lib
from .encoder import B200WhisperEncoder
from .feature_extractor import B200WhisperFeatureExtractor


class PinnedStaging:
    """Reusable pinned host staging for ragged int16 micro-batches: two flat buffers used alternately, each guarded by
    the CUDA event of the last copy that read it, so the serving path pays for `cudaHostAlloc` only when a batch is
    larger than anything seen before (never per request)."""

    def __init__(self):
        self._bufs = [None, None]
        self._events = [None, None]
        self._turn = 0

    def rows(self, n_rows: int, width: int, dtype):
        """A zeroed, contiguous, pinned [n_rows, width] tensor; call `.sent(stream)` after the copy is enqueued."""
        import torch

        i = self._turn
        need = n_rows * width
        if self._events[i] is not None:
            self._events[i].synchronize()        # normally long complete: two batches ago
        buf = self._bufs[i]
        if buf is None or buf.numel() < need or buf.dtype != dtype:
            grow = max(need, 2 * buf.numel() if buf is not None and buf.dtype == dtype else need)
            buf = torch.empty(grow, dtype=dtype).pin_memory()
            self._bufs[i] = buf
        view = buf[:need].view(n_rows, width)
        view.zero_()
        return view

    def sent(self, device=None):
        import torch

        if device is not None and torch.device(device).type != "cuda":   # host-only pipelines (tests): nothing in flight
            self._turn ^= 1
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        self._events[self._turn] = ev
        self._turn ^= 1


class B200LogMelEncoder:
    def __init__(self, feature_extractor: B200WhisperFeatureExtractor, encoder: B200WhisperEncoder):
        if feature_extractor.feature_size != encoder.config.num_mel_bins:
            raise _lib.TtasrError(-2, f"feature_size {feature_extractor.feature_size} != encoder num_mel_bins "
                                      f"{encoder.config.num_mel_bins}")
        self.feature_extractor = feature_extractor
        self.encoder = encoder
        self.device = encoder.device
        self._stage = None

    @property
    def launches_per_call(self) -> int:
        # frames + clamp kernels (front end) and the encoder's sequence minus its unused fp32->bf16 transpose
        return 2 + self.encoder.launches_per_forward - 1

    def encode_device(self, pcm, n_valid=None, out_dtype=None):
        """pcm: CUDA [B, n_samples] float32 / int16 -> [B, 1500, d] CUDA."""
        # only the bf16 time-major tensor the conv stem reads is produced: the fp32 `input_features` would be written
        # (1.5 MB per chunk at 128 bins) and never read
        _, tm = self.feature_extractor.extract(pcm, n_valid=n_valid, return_time_major=True, features=False)
        return self.encoder.encode(tm, out_dtype=out_dtype, time_major_ld=tm.shape[2])

    def encode_host(self, pcm_host, out_host=None, n_valid=None):
        """pcm_host: CPU tensor [B, n_samples] (pinned for full-rate copies).  Returns the CUDA tensor of hidden states,
        or — when `out_host` (a pinned CPU tensor [B, 1500, d] bf16) is given — `out_host` with the device-to-host copy
        ENQUEUED on the current stream, not complete: synchronise the stream (or the device) before reading it."""
        import torch

        B = pcm_host.shape[0]
        if self._stage is None or self._stage.shape[0] < B or self._stage.dtype != pcm_host.dtype or \
                self._stage.shape[1] != pcm_host.shape[1]:
            self._stage = torch.empty((B, pcm_host.shape[1]), dtype=pcm_host.dtype, device=self.device)
        dev = self._stage[:B]
        dev.copy_(pcm_host, non_blocking=True)
        hidden = self.encode_device(dev, n_valid=n_valid)
        if out_host is not None:
            out_host.copy_(hidden, non_blocking=True)
            return out_host
        return hidden

    def stream_host(self, batches, outs=None, consume=None):
        """Pipelined host-to-host run over a sequence of batches: while batch k is in the kernels, batch k+1's PCM is
        copied in and batch k-1's hidden states are copied out (three streams, double-buffered device staging).

        batches: sequence of pinned CPU tensors [B, n_samples] (float32 or int16, same shape/dtype);
        outs:    sequence of pinned CPU tensors [B, 1500, d] bf16 receiving the hidden states, or None to leave the
                 results on the device (then `consume(k, hidden)` — if given — is called with each batch's CUDA
                 tensor, stream-ordered on the current stream, before its buffer is reused two batches later).
        Returns when everything has been enqueued; synchronise the current stream (or the device) to wait."""
        import torch

        batches = list(batches)
        outs = list(outs) if outs is not None else None
        if outs is not None and len(batches) != len(outs):
            raise _lib.TtasrError(-2, "stream_host needs one output buffer per batch")
        if not batches:
            return
        main = torch.cuda.current_stream(self.device)
        if getattr(self, "_copy_streams", None) is None:
            self._copy_streams = (torch.cuda.Stream(self.device), torch.cuda.Stream(self.device))
        h2d, d2h = self._copy_streams
        shape, dtype = tuple(batches[0].shape), batches[0].dtype
        if getattr(self, "_ring", None) is None or self._ring[0].shape != shape or self._ring[0].dtype != dtype:
            self._ring = [torch.empty(shape, dtype=dtype, device=self.device) for _ in range(2)]
        in_ready = [torch.cuda.Event() for _ in batches]
        in_free = [torch.cuda.Event() for _ in batches]
        out_ready = [torch.cuda.Event() for _ in batches]
        hidden = [None, None]
        h2d.wait_stream(main)
        d2h.wait_stream(main)
        for k, pcm in enumerate(batches):
            buf = self._ring[k & 1]
            with torch.cuda.stream(h2d):
                if k >= 2:
                    h2d.wait_event(in_free[k - 2])  # the front end has consumed this staging buffer
                buf.copy_(pcm, non_blocking=True)
                in_ready[k].record(h2d)
            main.wait_event(in_ready[k])
            if k >= 2 and outs is not None:
                main.wait_event(out_ready[k - 2])  # hidden[k & 1] has been copied out before it is overwritten
            _, tm = self.feature_extractor.extract(buf, return_time_major=True, features=False)
            in_free[k].record(main)
            hidden[k & 1] = self.encoder.encode(tm, time_major_ld=tm.shape[2])
            if consume is not None:
                consume(k, hidden[k & 1])
            if outs is None:
                continue
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(d2h):
                d2h.wait_event(done)
                outs[k].copy_(hidden[k & 1], non_blocking=True)
                out_ready[k].record(d2h)
        main.wait_stream(d2h)


class GraphedLogMelEncoder:
    """Small-batch serving path (BASELINE.json configs[4]): one CUDA graph of the whole PCM -> hidden
    launch sequence per batch-size bucket, replayed for every micro-batch.  A forward is ~230 kernel launches and ~700
    tensor-map encodes on the host; at one to a few utterances that host work is a visible part of the latency
    (4.3 ms eager vs 3.85 ms replayed at B = 1, large-v3) and it would otherwise run on the server's event loop.

    Rows are ragged int16 utterances; a batch is padded up to the next bucket with empty rows (n_valid = 0: their
    front-end tiles are skipped).  Static buffers: PCM [bucket, n_samples] int16 and n_valid [bucket] per bucket."""

    def __init__(self, pipeline: B200LogMelEncoder, buckets=(1, 2, 4, 8, 16, 32, 64)):
        self.pipeline = pipeline
        self.buckets = tuple(sorted(set(int(b) for b in buckets)))
        self.n_samples = pipeline.feature_extractor.n_samples
        self._graphs = {}
        self._staging = PinnedStaging()
        self._pool = None   # one memory pool for all buckets: they are replayed one at a time and the result is copied out
        self.replays = 0
        # the graphs bake the encoder's workspace address in: size it for the largest bucket before any capture
        self._ws_generation = pipeline.encoder.reserve_workspace(self.buckets[-1])

    def _bucket(self, n: int) -> int:
        for b in self.buckets:
            if n <= b:
                return b
        raise _lib.TtasrError(-2, f"batch of {n} utterances exceeds the largest graph bucket {self.buckets[-1]}")

    def _build(self, bucket: int):
        import torch

        dev = self.pipeline.device
        with torch.cuda.device(dev):
            pcm = torch.zeros((bucket, self.n_samples), dtype=torch.int16, device=dev)
            n_valid = torch.zeros((bucket,), dtype=torch.int32, device=dev)
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):       # warm-up outside the capture (attribute setup, allocator)
                self.pipeline.encode_device(pcm, n_valid=n_valid)
            torch.cuda.current_stream(dev).wait_stream(side)
            if self._pool is None:
                self._pool = torch.cuda.graph_pool_handle()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, pool=self._pool):
                hidden = self.pipeline.encode_device(pcm, n_valid=n_valid)
        entry = (graph, pcm, n_valid, hidden)
        self._graphs[bucket] = entry
        return entry

    def build_all(self):
        """Capture every bucket now (server warm-up), so that no request pays for a capture."""
        for b in self.buckets:
            if b not in self._graphs:
                self._build(b)
        return len(self._graphs)

    def encode(self, rows, lens=None):
        """rows: list of 1-D int16 CPU tensors (or one [n, width] int16 CPU tensor, pinned for async copies);
        lens: real samples per row (default: row lengths, capped at 30 s).  Returns [n, 1500, d] bf16 CUDA (a copy:
        the graph's own output buffer is overwritten by the next replay of the same bucket)."""
        import torch

        if isinstance(rows, (list, tuple)):
            n = len(rows)
            lens = [min(int(r.numel()), self.n_samples) for r in rows] if lens is None else list(lens)
            width = (max(max(lens), 1) + 7) // 8 * 8
            host = self._staging.rows(n, width, torch.int16)
            for i, r in enumerate(rows):
                host[i, : lens[i]] = r[: lens[i]]
            staged = True
        else:
            host = rows
            staged = False
            n, width = host.shape
            lens = [min(width, self.n_samples)] * n if lens is None else list(lens)
        if n == 0:
            raise _lib.TtasrError(-2, "empty batch")
        width = min(width, self.n_samples)
        bucket = self._bucket(n)
        gen = self.pipeline.encoder.reserve_workspace(self.buckets[-1])
        if gen != self._ws_generation:   # somebody ran a larger eager batch: the captured workspace address is stale
            self._graphs.clear()
            self._ws_generation = gen
        graph, pcm, n_valid, hidden = self._graphs.get(bucket) or self._build(bucket)
        pcm[:n, :width].copy_(host[:, :width], non_blocking=True)
        if staged:
            self._staging.sent(self.pipeline.device)
        nv = torch.zeros((bucket,), dtype=torch.int32)
        nv[:n] = torch.tensor(lens, dtype=torch.int32)
        n_valid.copy_(nv, non_blocking=False)
        graph.replay()
        self.replays += 1
        return hidden[:n].clone()
