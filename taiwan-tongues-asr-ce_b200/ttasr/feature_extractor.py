"""Drop-in for the `WhisperFeatureExtractor` object the reference obtains from
`AutoFeatureExtractor.from_pretrained` (train_asr.py:518-527) and calls per example at train_asr.py:607-616:

    inputs = feature_extractor(sample["array"], sampling_rate=sr, return_attention_mask=flag)
    batch["input_features"] = inputs.get("input_features")[0]

Same constructor arguments, `__call__` keywords, attributes (`sampling_rate`, `model_input_names`, `n_samples`,
`nb_max_frames`), `pad()` for the collator (train_asr.py:290-298), `save_pretrained()` (train_asr.py:683) and
errors (ValueError on a wrong sampling rate or multi-channel input; feature_extraction_whisper.py:261-276) as the
Hugging Face class.  The arithmetic runs in the fused sm_100a kernel behind `ttasr_frontend_run`; `extract()` is the
batched device-resident fast path.  No CPU implementation is bundled.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Any

import numpy as np

from . import _lib, mel

try:  # the reference returns a transformers.BatchFeature; use it when present so .get()/convert_to_tensors match
    from transformers.feature_extraction_utils import BatchFeature as _BatchFeature
except Exception:  # pragma: no cover - transformers is a dependency of the reference, not of the kernels

    class _BatchFeature(dict):
        def convert_to_tensors(self, tensor_type=None):
            if tensor_type in ("pt", "torch"):
                import torch

                for k, v in list(self.items()):
                    self[k] = torch.as_tensor(np.asarray(v))
            elif tensor_type in ("np", "numpy"):
                for k, v in list(self.items()):
                    self[k] = np.asarray(v)
            return self


class B200WhisperFeatureExtractor:
    model_input_names = ["input_features"]
    feature_extractor_type = "WhisperFeatureExtractor"

    def __init__(self, feature_size: int = 80, sampling_rate: int = 16000, hop_length: int = 160,
                 chunk_length: int = 30, n_fft: int = 400, padding_value: float = 0.0, dither: float = 0.0,
                 return_attention_mask: bool = False, device: str | int | None = None, **kwargs: Any):
        self.feature_size = feature_size
        self.sampling_rate = sampling_rate
        self.hop_length = hop_length
        self.chunk_length = chunk_length
        self.n_fft = n_fft
        self.padding_value = padding_value
        self.padding_side = kwargs.pop("padding_side", "right")
        self.dither = dither
        self.return_attention_mask = return_attention_mask
        self.n_samples = chunk_length * sampling_rate
        self.nb_max_frames = self.n_samples // hop_length
        self.mel_filters = mel.slaney_mel_filters(feature_size, n_fft, sampling_rate)
        self._device = device
        self._handles = {}   # CUDA device index -> native handle (tables and scratch live on that device)
        self._extra = kwargs

    # ------------------------------------------------------------------ native handle
    def _torch_device(self):
        import torch

        if not torch.cuda.is_available():
            raise _lib.TtasrError(-3, "no CUDA device: the B200 front end has no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device()) if self._device is None else torch.device(self._device)
        if dev.type != "cuda":
            raise _lib.TtasrError(-3, f"device {dev} is not a CUDA device: the B200 front end has no CPU fallback")
        return dev

    def _native(self, dev=None):
        """The native front end of `dev` (default: this object's device).  One handle per device: its tables and
        per-call scratch are device memory, and a handle must not be shared by concurrent streams (ttasr_abi.h) — this
        wrapper is single-stream per device."""
        import torch

        dev = self._torch_device() if dev is None else torch.device(dev)
        if dev.type != "cuda":
            raise _lib.TtasrError(-3, f"device {dev} is not a CUDA device: the B200 front end has no CPU fallback")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        h = self._handles.get(idx)
        if h is None:
            lib = _lib.lib()
            filt = np.ascontiguousarray(self.mel_filters, dtype=np.float32)
            win = np.ascontiguousarray(mel.periodic_hann(self.n_fft), dtype=np.float32)
            h = C.c_void_p()
            with torch.cuda.device(idx):
                _lib.check(lib.ttasr_frontend_create(self.feature_size, self.n_fft, self.hop_length, self.n_samples,
                                                     filt.ctypes.data, win.ctypes.data, C.byref(h)))
            self._handles[idx] = h
        return h

    def mel_mode(self, dev=None) -> int:
        """80: the native front end runs the mel projection compiled in for the 80-filter Whisper bank (chosen when
        `mel_filters` is bit-identical to the table baked into the kernel); 0: the generic program built from
        `mel_filters` (every other bank, the 128-filter one included).  Same features either way."""
        out = C.c_int()
        _lib.check(_lib.lib().ttasr_frontend_mel_mode(self._native(dev), C.byref(out)))
        return int(out.value)

    def __del__(self):
        handles, self._handles = getattr(self, "_handles", {}), {}
        for h in handles.values():
            try:
                _lib.lib().ttasr_frontend_destroy(h)
            except Exception:
                pass

    # ------------------------------------------------------------------ batched device fast path
    def extract(self, pcm, n_valid=None, return_time_major: bool = False, features: bool = True,
                clamp_decades: float = 8.0):
        """pcm: CUDA tensor [B, >= n_samples] float32 (in [-1, 1]) or int16 -> float32 [B, feature_size, 3000] (CUDA).

        n_valid: optional int32 CUDA tensor [B]; samples from n_valid[b] on count as zero padding and are not read
        (rows may then be shorter than 30 s).  return_time_major additionally returns the bf16 [B, 3000, ld] copy the
        encoder's conv stem consumes directly; with features=False (and return_time_major) the fp32 `input_features`
        are not written at all and (None, time_major) is returned — the PCM -> hidden-state pipeline's mode.
        clamp_decades: the `max(x, x.max() - 8)` range per chunk; float("inf") returns unclamped features."""
        import torch

        pcm = _lib.require_cuda_tensor(pcm, "pcm")
        h = self._native(pcm.device)
        if not features and not return_time_major:
            raise _lib.TtasrError(-1, "features=False needs return_time_major=True")
        if pcm.dim() == 1:
            pcm = pcm.unsqueeze(0)
        if pcm.dim() != 2:
            raise ValueError(f"Only mono-channel audio is supported for input to {self.__class__.__name__}")
        if pcm.dtype == torch.float32:
            dtype = _lib.PCM_F32
        elif pcm.dtype == torch.int16:
            dtype = _lib.PCM_I16
        else:
            raise _lib.TtasrError(-1, f"pcm dtype must be float32 or int16, got {pcm.dtype}")
        B, stride = pcm.shape
        nv_ptr = None
        if n_valid is not None:
            n_valid = _lib.require_cuda_tensor(n_valid, "n_valid").to(torch.int32)
            if n_valid.numel() != B:
                raise _lib.TtasrError(-2, "n_valid must hold one length per row")
            # (lengths beyond the row are clamped by the kernel; no host sync here, so the call is graph-capturable)
            nv_ptr = n_valid.data_ptr()
        elif stride < self.n_samples:
            raise _lib.TtasrError(-2, f"rows hold {stride} samples < {self.n_samples}; pass n_valid for ragged input")
        with torch.cuda.device(pcm.device):
            feats = torch.empty((B, self.feature_size, self.nb_max_frames), dtype=torch.float32,
                                device=pcm.device) if features else None
            tm, ld = None, 0
            if return_time_major:
                ld = (self.feature_size + 7) // 8 * 8
                tm = torch.empty((B, self.nb_max_frames, ld), dtype=torch.bfloat16, device=pcm.device)
            _lib.check(_lib.lib().ttasr_frontend_run_ex(
                h, pcm.data_ptr(), dtype, B, stride, nv_ptr, feats.data_ptr() if feats is not None else None,
                tm.data_ptr() if tm is not None else None,
                ld, float(clamp_decades), _lib.current_stream_ptr(pcm.device)))
        return (feats, tm) if return_time_major else feats

    # ------------------------------------------------------------------ the reference's call surface
    def __call__(self, raw_speech, truncation: bool = True, pad_to_multiple_of: int | None = None,
                 return_tensors: str | None = None, return_attention_mask: bool | None = None,
                 padding: str | None = "max_length", max_length: int | None = None,
                 sampling_rate: int | None = None, do_normalize: bool | None = None,
                 device: str | None = None, **kwargs):
        import torch

        if sampling_rate is not None and sampling_rate != self.sampling_rate:
            raise ValueError(
                f"The model corresponding to this feature extractor: {self.__class__.__name__} was trained using a"
                f" sampling rate of {self.sampling_rate}. Please make sure that the provided `raw_speech` input"
                f" was sampled with {self.sampling_rate} and not {sampling_rate}.")
        if self.dither != 0.0:
            raise NotImplementedError("dither != 0 is not implemented by the B200 front end")
        if return_attention_mask is None:
            return_attention_mask = self.return_attention_mask
        if padding not in ("max_length", True) or not truncation or pad_to_multiple_of is not None or (
                max_length not in (None, self.n_samples)):
            raise NotImplementedError(
                "the B200 front end implements the reference's configuration only: padding='max_length', "
                f"truncation=True, max_length={self.n_samples}")
        is_batched_numpy = isinstance(raw_speech, np.ndarray) and raw_speech.ndim > 1
        if is_batched_numpy and raw_speech.ndim > 2:
            raise ValueError(f"Only mono-channel audio is supported for input to {self}")
        is_batched = is_batched_numpy or (
            isinstance(raw_speech, (list, tuple)) and len(raw_speech) > 0
            and isinstance(raw_speech[0], (np.ndarray, tuple, list)))
        rows = [np.asarray(r, dtype=np.float32).reshape(-1) for r in raw_speech] if is_batched else [
            np.asarray(raw_speech, dtype=np.float32).reshape(-1)]
        B = len(rows)
        lengths = np.array([min(len(r), self.n_samples) for r in rows], dtype=np.int32)
        host = np.full((B, self.n_samples), self.padding_value, dtype=np.float32)
        for i, r in enumerate(rows):
            host[i, : lengths[i]] = r[: lengths[i]]
        mask = None
        if return_attention_mask or do_normalize:
            mask = (np.arange(self.n_samples)[None, :] < lengths[:, None]).astype(np.int32)
        if do_normalize:  # zero-mean / unit-variance over the unpadded part (feature_extraction_whisper.py:166-187)
            for i in range(B):
                seg = host[i, : lengths[i]]
                host[i, : lengths[i]] = (seg - seg.mean()) / np.sqrt(seg.var() + 1e-7)
                host[i, lengths[i]:] = self.padding_value
        dev = self._torch_device() if device in (None, "cpu") else torch.device(device)
        pcm = torch.from_numpy(host).to(dev, non_blocking=False)
        # right padding with 0.0 is what n_valid means to the kernel: padded tiles skip the transform altogether
        n_valid = torch.from_numpy(lengths).to(dev) if (self.padding_value == 0.0 and not do_normalize) else None
        feats = self.extract(pcm, n_valid=n_valid)
        out = _BatchFeature({"input_features": feats.cpu().numpy()})
        if return_attention_mask:
            out["attention_mask"] = mask[:, :: self.hop_length]
        if return_tensors is not None:
            out = out.convert_to_tensors(return_tensors)
        return out

    # ------------------------------------------------------------------ collator / persistence helpers
    def pad(self, processed_features, padding=True, max_length=None, truncation=False, pad_to_multiple_of=None,
            return_attention_mask=None, return_tensors=None):
        """Collate already-extracted features (all [feature_size, 3000]): the use at train_asr.py:296-298."""
        if isinstance(processed_features, (list, tuple)):
            keys = processed_features[0].keys()
            processed_features = {k: [f[k] for f in processed_features] for k in keys}
        feats = [np.asarray(f, dtype=np.float32) for f in processed_features[self.model_input_names[0]]]
        shapes = {f.shape for f in feats}
        if len(shapes) != 1:
            raise ValueError(f"input_features of different shapes cannot be collated: {sorted(shapes)}")
        out = _BatchFeature({self.model_input_names[0]: np.stack(feats, axis=0)})
        if "attention_mask" in processed_features:
            out["attention_mask"] = np.stack([np.asarray(m) for m in processed_features["attention_mask"]], axis=0)
        if return_tensors is not None:
            out = out.convert_to_tensors(return_tensors)
        return out

    def to_dict(self) -> dict:
        return {
            "feature_extractor_type": self.feature_extractor_type, "feature_size": self.feature_size,
            "sampling_rate": self.sampling_rate, "hop_length": self.hop_length, "chunk_length": self.chunk_length,
            "n_fft": self.n_fft, "padding_value": self.padding_value, "padding_side": self.padding_side,
            "dither": self.dither, "return_attention_mask": self.return_attention_mask,
            "n_samples": self.n_samples, "nb_max_frames": self.nb_max_frames, "processor_class": "WhisperProcessor",
        }

    def save_pretrained(self, save_directory: str, **kwargs) -> list[str]:
        os.makedirs(save_directory, exist_ok=True)
        path = os.path.join(save_directory, "preprocessor_config.json")
        with open(path, "w", encoding="utf-8") as f:
            json.dump(self.to_dict(), f, indent=2, sort_keys=True)
            f.write("\n")
        return [path]

    @classmethod
    def from_pretrained(cls, path: str, **kwargs):
        cfg_path = os.path.join(path, "preprocessor_config.json") if os.path.isdir(path) else path
        with open(cfg_path, encoding="utf-8") as f:
            cfg = json.load(f)
        for k in ("feature_extractor_type", "processor_class", "n_samples", "nb_max_frames"):
            cfg.pop(k, None)
        cfg.update(kwargs)
        return cls(**cfg)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__} {json.dumps(self.to_dict(), indent=2, sort_keys=True)}"
