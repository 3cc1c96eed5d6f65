"""`B200ASR`: the reference's streaming ASR plugin (`ASRInterface`, REF api/stt_streaming/src/asr/asr_interface.py:1-15)
on the B200 hot path, selectable from `ASRFactory.create_asr_pipeline("b200", ...)` (asr_factory.py:9-30).

What changes against `FasterWhisperASR.transcribe` (faster_whisper_asr.py:151-267):
  * the int16 PCM of `client.scratch_buffer` goes to the GPU as is (no temp WAV round trip, audio_utils.py:5-29);
    right-padding to 30 s is virtual (`n_valid`), the kernel never reads the padding;
  * utterances of all connected clients that are ready within `batch_window_s` are run as ONE front-end + encoder
    launch instead of blocking the event loop once per client;
  * log-mel + encoder run in libttasr_b200.so; decoding stays with the host application through `decode_fn`
    (HF `generate(encoder_outputs=...)`, a CTranslate2 decoder, ...), which receives the hidden states on the GPU.

The result dictionary has the reference's fields (faster_whisper_asr.py:240-255).
"""
from __future__ import annotations

import asyncio
import time
from dataclasses import dataclass, field
from typing import Any, Callable

from . import _lib
from .pipeline import B200LogMelEncoder


@dataclass
class Utterance:
    """One client's scratch buffer on its way through the batcher."""
    pcm_i16: Any                      # torch.int16 CPU tensor [n]
    client: Any
    future: Any = None
    meta: dict = field(default_factory=dict)


def pcm_bytes_to_tensor(buf, samples_width: int = 2):
    """bytes / bytearray of little-endian int16 mono PCM (the WebSocket wire format, streaming_asr.py:223-226)."""
    import numpy as np
    import torch

    if samples_width != 2:
        raise ValueError(f"only 16-bit PCM is supported (samples_width={samples_width})")
    n = len(buf) // 2
    return torch.from_numpy(np.frombuffer(bytes(buf[: 2 * n]), dtype="<i2").copy())


class MicroBatcher:
    """Collects utterances for at most `window_s` (or until `max_batch`) and encodes them in one launch."""

    def __init__(self, pipeline: B200LogMelEncoder, window_s: float = 0.005, max_batch: int = 64,
                 use_graphs: bool = False):
        """use_graphs: replay one CUDA graph per batch-size bucket (pipeline.GraphedLogMelEncoder) instead of issuing
        the ~230 launches of a forward from the event loop."""
        self.pipeline = pipeline
        self.window_s = window_s
        self.max_batch = max_batch
        self.graphed = None
        if use_graphs:
            from .pipeline import GraphedLogMelEncoder

            # dense at the small end, where a padded row costs a visible share of the launch
            ladder = list(range(1, 9)) + list(range(10, 17, 2)) + list(range(20, 33, 4)) + list(range(40, 65, 8)) + \
                list(range(80, 257, 16))
            buckets = [b for b in ladder if b < max_batch] + [max_batch]
            self.graphed = GraphedLogMelEncoder(pipeline, buckets)
        self.n_samples = pipeline.feature_extractor.n_samples
        from .pipeline import PinnedStaging

        self._staging = PinnedStaging()
        self._pending: list[Utterance] = []
        self._flusher = None
        self.launches = 0
        self.encoded = 0

    def encode_batch(self, utterances: list[Utterance]):
        """Synchronous core: ragged int16 rows (each at most one 30 s window: B200ASR.transcribe splits longer buffers
        into several rows) -> hidden states [B, 1500, d] (CUDA)."""
        import torch

        B = len(utterances)
        for u in utterances:
            if int(u.pcm_i16.numel()) > self.n_samples:
                raise _lib.TtasrError(-2, f"utterance of {int(u.pcm_i16.numel())} samples exceeds one {self.n_samples}-sample "
                                          "window: split it into windows (B200ASR.transcribe does) instead of truncating")
        lens = [int(u.pcm_i16.numel()) for u in utterances]
        if self.graphed is not None:
            hidden = self.graphed.encode([u.pcm_i16 for u in utterances], lens)
            self.launches += 1
            self.encoded += B
            return hidden, lens
        width = max(max(lens), 1)
        width = (width + 7) // 8 * 8  # 16-byte rows keep the bulk-copy path
        host = self._staging.rows(B, width, torch.int16)
        for i, u in enumerate(utterances):
            host[i, : lens[i]] = u.pcm_i16[: lens[i]]
        dev = self.pipeline.device
        pcm = host.to(dev, non_blocking=True)
        self._staging.sent(dev)
        n_valid = torch.tensor(lens, dtype=torch.int32).to(dev, non_blocking=True)
        hidden = self.pipeline.encode_device(pcm, n_valid=n_valid)
        self.launches += 1
        self.encoded += B
        return hidden, lens

    async def submit(self, utt: Utterance):
        loop = asyncio.get_running_loop()
        utt.future = loop.create_future()
        self._pending.append(utt)
        if len(self._pending) >= self.max_batch:
            self._flush()
        elif self._flusher is None:
            self._flusher = loop.call_later(self.window_s, self._flush)
        return await utt.future

    def _flush(self):
        if self._flusher is not None:
            self._flusher.cancel()
            self._flusher = None
        batch, self._pending = self._pending[: self.max_batch], self._pending[self.max_batch:]
        if not batch:
            return
        try:
            hidden, lens = self.encode_batch(batch)
            for i, u in enumerate(batch):
                if not u.future.done():
                    u.future.set_result((hidden[i: i + 1], lens[i]))
        except Exception as e:  # surfaced to every waiter; the reference's wrappers log and return None
            for u in batch:
                if not u.future.done():
                    u.future.set_exception(e)
        if self._pending:
            self._flusher = asyncio.get_running_loop().call_later(self.window_s, self._flush)


class B200ASR:
    """ASRInterface implementation (duck-typed: `async transcribe(client)`, `warm_up()`).

    Constructible the way the reference's factory constructs its plugin — `ASRFactory.create_asr_pipeline(type,
    **kwargs)` -> `FasterWhisperASR(**kwargs)` with `model_size=...` (asr_factory.py:9-30, streaming_asr.py:116-121):

        B200ASR(model_size="large-v3-turbo")                    # resolves the model directory like the reference does
        B200ASR(model_size=..., decode_fn=my_decoder)           # bring your own decoder
        B200ASR(pipeline, decode_fn)                            # round-1 form: ready-made objects

    `from_kwargs` is the same thing spelled as a factory hook."""

    def __init__(self, pipeline: B200LogMelEncoder | None = None, decode_fn: Callable[[Any, dict], dict] | None = None,
                 batch_window_s: float = 0.005, max_batch: int = 64, language: str = "zh",
                 use_graphs: bool = False, **kwargs):
        """decode_fn(hidden [n_windows, 1500, d] CUDA bf16, info dict) -> {"text": str, "words": [...], "language": ...,
        "language_probability": ...}; it owns beam search and text post-processing exactly as the reference's
        decoder does today.  Keyword arguments the reference passes (`model_size`) or reads back (`device`,
        `compute_type`, `model_size`, `model_path`: health checks, faster_whisper_asr.py:110-114) are honoured:

          model_size   name or path; looked up as <root>/<model_size> under `model_root` (default: the reference's
                       project-root rule relative to `reference_api_dir` when given), $TTASR_MODEL_ROOT, the cwd
          residual     encoder residual-stream mode (see include/ttasr_abi.h); default = library default
          warm_up_wav  path of the warm-up recording (default: <reference stt_streaming dir>/warm_up.wav if known)
        """
        self.model_size = kwargs.get("model_size")
        self.model_path = None
        self.weights_format = None
        self._stt_dir = kwargs.get("stt_streaming_dir")
        self._warm_up_wav = kwargs.get("warm_up_wav")
        if pipeline is None:
            pipeline = self._pipeline_from_kwargs(kwargs)
        if decode_fn is None:
            decode_fn = self._default_decoder(kwargs)
        self.pipeline = pipeline
        self.decode_fn = decode_fn
        self.batcher = MicroBatcher(pipeline, batch_window_s, max_batch, use_graphs=use_graphs)
        self.language = language
        self.device = "cuda"
        self.compute_type = "bfloat16"
        self._executor = None
        self.default_transcribe_kwargs = {"beam_size": 5, "condition_on_previous_text": True,
                                          "initial_prompt": "繁體中文", "language": language}

    @classmethod
    def from_kwargs(cls, **kwargs) -> "B200ASR":
        """Factory hook: `if type == "b200": return B200ASR.from_kwargs(**kwargs)` in ASRFactory.create_asr_pipeline."""
        return cls(**kwargs)

    # ------------------------------------------------------------------ construction from the reference's kwargs
    def _pipeline_from_kwargs(self, kwargs) -> B200LogMelEncoder:
        import os

        from .encoder import B200WhisperEncoder
        from .feature_extractor import B200WhisperFeatureExtractor
        from .model_dir import load_encoder_weights, resolve_model_dir

        model_size = kwargs.get("model_size", "large-v3-turbo")    # the reference's default (faster_whisper_asr.py:21)
        self.model_size = model_size
        roots = []
        if kwargs.get("model_root"):
            roots.append(kwargs["model_root"])
        if self._stt_dir:                                          # <project root> = stt_streaming/../.. (:27-33)
            roots.append(os.path.dirname(os.path.dirname(os.path.abspath(self._stt_dir))))
        model_dir = resolve_model_dir(model_size, roots)
        if model_dir is None:
            raise FileNotFoundError(
                f"model '{model_size}' not found under {roots or ['$TTASR_MODEL_ROOT', os.getcwd()]}: the reference "
                "falls back to a hub download here (faster_whisper_asr.py:50-51), which an offline B200 box cannot do")
        cfg, sd, fmt = load_encoder_weights(model_dir)
        self.model_path, self.weights_format = model_dir, fmt
        enc = B200WhisperEncoder(cfg, sd, residual=kwargs.get("residual"))
        return B200LogMelEncoder(B200WhisperFeatureExtractor(feature_size=cfg.num_mel_bins), enc)

    def _default_decoder(self, kwargs):
        """Without a caller-supplied decode_fn: the Hugging Face decoder of the same model directory, fed through
        `generate(encoder_outputs=...)` (decode_handoff.hf_generate) — beam 5, language zh, as the reference's
        default_transcribe_kwargs (faster_whisper_asr.py:139-149).  Needs the HF layout + tokenizer in the directory."""
        import os

        if not self.model_path or self.weights_format != "hf":
            raise _lib.TtasrError(-1, "B200ASR needs decode_fn=...: no Hugging Face decoder can be built from "
                                      f"{self.model_path or 'the given pipeline'} (the CTranslate2 decoder lives in "
                                      "faster-whisper, which this package does not import)")
        try:
            import torch
            from transformers import WhisperForConditionalGeneration, WhisperProcessor
        except Exception as e:  # pragma: no cover
            raise _lib.TtasrError(-1, f"B200ASR needs decode_fn=...: transformers is not importable ({e})")
        from .decode_handoff import encoder_outputs

        model = WhisperForConditionalGeneration.from_pretrained(self.model_path, torch_dtype=torch.bfloat16)
        model = model.to("cuda").eval()
        proc = WhisperProcessor.from_pretrained(self.model_path) if os.path.exists(
            os.path.join(self.model_path, "tokenizer.json")) or os.path.exists(
            os.path.join(self.model_path, "vocab.json")) else None
        beams = int(kwargs.get("beam_size", 5))

        def decode(hidden, info):
            with torch.no_grad():
                ids = model.generate(encoder_outputs=encoder_outputs(hidden, torch.bfloat16), num_beams=beams,
                                     language=info.get("language"), task="transcribe")
            texts = proc.batch_decode(ids, skip_special_tokens=True) if proc is not None else [str(i.tolist()) for i in ids]
            return {"text": " ".join(t.strip() for t in texts).strip(), "words": [], "language": info.get("language")}

        return decode

    # ------------------------------------------------------------------ ASRInterface
    async def transcribe(self, client):
        try:
            pcm = pcm_bytes_to_tensor(client.scratch_buffer, getattr(client, "samples_width", 2))
            if pcm.numel() == 0:
                return None
            # buffers longer than one 30 s window (long speech without a VAD gap, large BUFFERING_CHUNK_LENGTH_SECONDS)
            # are split into consecutive windows, each its own row of the micro-batch — WhisperModel.transcribe windows
            # buffers of any length too (faster_whisper_asr.py:170); nothing is dropped
            n_win = self.batcher.n_samples
            windows = [pcm[i: i + n_win] for i in range(0, int(pcm.numel()), n_win)]
            parts = await asyncio.gather(*(self.batcher.submit(Utterance(w, client)) for w in windows))
            import torch

            hidden = parts[0][0] if len(parts) == 1 else torch.cat([p[0] for p in parts], dim=0)
            n = sum(p[1] for p in parts)
            info = {"language": self.language, "n_samples": n, "window_samples": [p[1] for p in parts],
                    "sampling_rate": self.pipeline.feature_extractor.sampling_rate}
            # decoding is host work of unbounded length (beam search): keep it off the event loop so other clients'
            # buffers keep flowing into the micro-batcher meanwhile
            loop = asyncio.get_running_loop()
            out = await loop.run_in_executor(self._decode_executor(), self.decode_fn, hidden, info)
            out = out or {}
            text = out.get("text")
            if not text:
                return None
            t0 = getattr(client, "last_start_time", 0) or 0
            words = out.get("words") or []
            duration = words[-1]["end"] if words else n / info["sampling_rate"]
            return {
                "language": out.get("language", self.language),
                "language_probability": out.get("language_probability"),
                "final": True,
                "text": text,
                "duration": duration,
                "words": [{"word": w.get("word", ""), "start": (w.get("start", 0) or 0) + t0,
                           "end": (w.get("end", 0) or 0) + t0, "probability": w.get("probability")} for w in words],
            }
        except _lib.TtasrError:
            raise  # no silent CPU fallback: a missing GPU / library is a deployment error
        except Exception:
            return None  # same contract as the reference wrapper (faster_whisper_asr.py:260-267)

    def _decode_executor(self):
        if self._executor is None:
            from concurrent.futures import ThreadPoolExecutor

            self._executor = ThreadPoolExecutor(max_workers=1, thread_name_prefix="ttasr-decode")
        return self._executor

    def warm_up(self):
        """Module load, tensor-map cache, allocator, CUDA graphs — on the reference's own warm-up recording when it can
        be found (`<stt_streaming>/warm_up.wav`, 44.1 kHz stereo; faster_whisper_asr.py:269-294), decoded with the
        standard library and brought to 16 kHz mono chunks by `B200AudioIngest`; else on one second of silence."""
        import os
        import torch

        t0 = time.time()
        wav = self._warm_up_wav or (os.path.join(self._stt_dir, "warm_up.wav") if self._stt_dir else None)
        info = {"audio": "1 s of silence"}
        if wav and os.path.exists(wav):
            from .ingest import B200AudioIngest, read_wav_pcm16

            frames, sr = read_wav_pcm16(wav)
            dev = self.pipeline.device
            chunks, n_valid = B200AudioIngest(sr).load(torch.from_numpy(frames).to(dev))
            hidden = self.pipeline.encode_device(chunks, n_valid=n_valid)
            info = {"audio": wav, "orig_sr": sr, "channels": int(frames.shape[1]), "chunks": int(chunks.shape[0]),
                    "seconds": float(n_valid.sum().item()) / 16000.0}
        utt = Utterance(torch.zeros(16000, dtype=torch.int16), None)
        hidden, _ = self.batcher.encode_batch([utt])
        graphs = self.batcher.graphed.build_all() if self.batcher.graphed is not None else 0
        torch.cuda.synchronize(self.pipeline.device)
        return {"warm_up_seconds": time.time() - t0, "hidden_shape": tuple(hidden.shape), "cuda_graphs": graphs, **info}
