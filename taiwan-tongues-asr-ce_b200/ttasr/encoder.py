"""Drop-in for the encoder forward the reference reaches through `WhisperForConditionalGeneration`
(train_asr.py:539-545,697-716,736-740 -> WhisperEncoder.forward, modeling_whisper.py:593-647) and through
`faster_whisper.WhisperModel.encode` (asr_core.py:159-167 et al.):

    B200WhisperEncoder(config, state_dict).encode(input_features)  -> [B, 1500, d] bf16 on the GPU
    B200WhisperEncoder.forward(input_features, **_)                -> BaseModelOutput(last_hidden_state=...)

`state_dict` uses the Hugging Face encoder names (`conv1.weight`, `layers.0.self_attn.q_proj.weight`, ...; a
`model.encoder.` / `encoder.` prefix is accepted).  All arithmetic runs in the sm_100a kernels behind
`ttasr_encoder_forward`; there is no PyTorch or CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib


@dataclass(frozen=True)
class EncoderConfig:
    d_model: int
    encoder_layers: int
    encoder_attention_heads: int
    encoder_ffn_dim: int
    num_mel_bins: int
    max_source_positions: int = 1500

    @classmethod
    def from_any(cls, cfg) -> "EncoderConfig":
        if isinstance(cfg, cls):
            return cls(**cfg.__dict__)
        get = (lambda k, d=None: cfg.get(k, d)) if isinstance(cfg, dict) else (lambda k, d=None: getattr(cfg, k, d))
        return cls(get("d_model"), get("encoder_layers"), get("encoder_attention_heads"), get("encoder_ffn_dim"),
                   get("num_mel_bins"), get("max_source_positions", 1500))

    @classmethod
    def named(cls, name: str) -> "EncoderConfig":
        table = {
            "tiny": (384, 4, 6, 1536, 80), "base": (512, 6, 8, 2048, 80), "small": (768, 12, 12, 3072, 80),
            "medium": (1024, 24, 16, 4096, 80), "large-v2": (1280, 32, 20, 5120, 80),
            "large-v3": (1280, 32, 20, 5120, 128), "large-v3-turbo": (1280, 32, 20, 5120, 128),
        }
        return cls(*table[name])

    def flops_per_chunk(self) -> int:
        d, f, L, T = self.d_model, self.encoder_ffn_dim, self.encoder_layers, self.max_source_positions
        return (2 * 2 * T * 3 * self.num_mel_bins * d + 2 * T * 3 * d * d
                + L * (8 * T * d * d + 4 * T * T * d + 4 * T * d * f))


class _Output(dict):
    """Minimal BaseModelOutput stand-in (attribute + index + key access) when transformers is unavailable."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __getitem__(self, k):
        if isinstance(k, int):
            return list(self.values())[k]
        return dict.__getitem__(self, k)


def _strip_prefix(sd: dict) -> dict:
    for prefix in ("model.encoder.", "encoder.", ""):
        if prefix + "conv1.weight" in sd:
            return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    raise KeyError("state_dict has no conv1.weight (expected Hugging Face WhisperEncoder names)")


class B200WhisperEncoder:
    main_input_name = "input_features"

    def __init__(self, config, state_dict: dict, device=None, residual: str | None = None):
        """`residual`: how the residual stream is held between the blocks — "split" (library default: bf16 hi + lo
        pair, LayerNorms folded into the QKV / fc1 GEMMs), "f32" (fp32 stream + LayerNorm kernels), "bf16" (hi only),
        or None = the TTASR_RESIDUAL environment variable, else the default (include/ttasr_abi.h)."""
        import torch

        if residual not in _lib.RESIDUAL_MODES:
            raise _lib.TtasrError(-1, f"residual must be one of 'f32', 'split', 'bf16' or None (got {residual!r})")
        self.residual = residual

        self.config = EncoderConfig.from_any(config)
        c = self.config
        if not torch.cuda.is_available():
            raise _lib.TtasrError(-3, "no CUDA device: the B200 encoder has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        sd = _strip_prefix(state_dict)
        keep = []  # device copies handed to ttasr_encoder_create (which packs its own; freed right after)

        def mat(name):
            t = torch.as_tensor(sd[name]).detach().to(self.device, torch.bfloat16).contiguous()
            keep.append(t)
            return t.data_ptr()

        def vec(name, shape=None):
            t = torch.as_tensor(sd[name]).detach().to(self.device, torch.float32).contiguous()
            if shape is not None and tuple(t.shape) != tuple(shape):
                raise _lib.TtasrError(-2, f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
            keep.append(t)
            return t.data_ptr()

        d, f = c.d_model, c.encoder_ffn_dim
        if tuple(sd["conv1.weight"].shape) != (d, c.num_mel_bins, 3):
            raise _lib.TtasrError(-2, f"conv1.weight has shape {tuple(sd['conv1.weight'].shape)}, "
                                      f"expected {(d, c.num_mel_bins, 3)}")
        layers = (_lib.LayerWeights * c.encoder_layers)()
        for i in range(c.encoder_layers):
            p = f"layers.{i}."
            lw = layers[i]
            lw.ln1_g, lw.ln1_b = vec(p + "self_attn_layer_norm.weight", (d,)), vec(p + "self_attn_layer_norm.bias", (d,))
            lw.wq, lw.bq = mat(p + "self_attn.q_proj.weight"), vec(p + "self_attn.q_proj.bias", (d,))
            lw.wk = mat(p + "self_attn.k_proj.weight")
            lw.wv, lw.bv = mat(p + "self_attn.v_proj.weight"), vec(p + "self_attn.v_proj.bias", (d,))
            lw.wo, lw.bo = mat(p + "self_attn.out_proj.weight"), vec(p + "self_attn.out_proj.bias", (d,))
            lw.ln2_g, lw.ln2_b = vec(p + "final_layer_norm.weight", (d,)), vec(p + "final_layer_norm.bias", (d,))
            lw.w1, lw.b1 = mat(p + "fc1.weight"), vec(p + "fc1.bias", (f,))
            lw.w2, lw.b2 = mat(p + "fc2.weight"), vec(p + "fc2.bias", (d,))
        w = _lib.Weights()
        w.conv1_w, w.conv1_b = mat("conv1.weight"), vec("conv1.bias", (d,))
        w.conv2_w, w.conv2_b = mat("conv2.weight"), vec("conv2.bias", (d,))
        w.pos = vec("embed_positions.weight", (c.max_source_positions, d))
        w.ln_post_g, w.ln_post_b = vec("layer_norm.weight", (d,)), vec("layer_norm.bias", (d,))
        w.layers = layers
        cfg = _lib.EncoderCfg(d, c.encoder_layers, c.encoder_attention_heads, f, c.num_mel_bins,
                              c.max_source_positions)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            _lib.check(_lib.lib().ttasr_encoder_create_ex(C.byref(cfg), C.byref(w), _lib.RESIDUAL_MODES[residual],
                                                          C.byref(h)))
        self._handle = h
        del keep
        self._ws = None
        n = C.c_int64()
        _lib.check(_lib.lib().ttasr_encoder_launch_count(self._handle, C.byref(n)))
        self.launches_per_forward = int(n.value)

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h is not None:
            try:
                _lib.lib().ttasr_encoder_destroy(h)
            except Exception:
                pass

    @classmethod
    def from_hf(cls, model, device=None, residual=None) -> "B200WhisperEncoder":
        """Build from a Hugging Face WhisperEncoder / WhisperModel / WhisperForConditionalGeneration."""
        enc = model.get_encoder() if hasattr(model, "get_encoder") else model
        return cls(EncoderConfig.from_any(enc.config), enc.state_dict(), device=device, residual=residual)

    # ------------------------------------------------------------------ per-stage timing (bench / profiling)
    def profile(self, on: bool = True) -> None:
        _lib.check(_lib.lib().ttasr_encoder_profile_enable(self._handle, 1 if on else 0))

    def profile_read(self, reset: bool = True) -> dict:
        """{stage name: (total ms, launches)} accumulated since the last reset; waits for the recorded events."""
        ms = (C.c_double * _lib.PROFILE_KINDS)()
        n = (C.c_int64 * _lib.PROFILE_KINDS)()
        _lib.check(_lib.lib().ttasr_encoder_profile_read(self._handle, ms, n, 1 if reset else 0))
        names = [_lib.lib().ttasr_encoder_profile_kind_name(k).decode() for k in range(_lib.PROFILE_KINDS)]
        return {names[k]: (float(ms[k]), int(n[k])) for k in range(_lib.PROFILE_KINDS)}

    # ------------------------------------------------------------------ forward
    def workspace_bytes(self, batch: int) -> int:
        n = C.c_size_t()
        _lib.check(_lib.lib().ttasr_encoder_workspace_bytes(self._handle, batch, C.byref(n)))
        return int(n.value)

    def _workspace(self, batch: int):
        import torch

        need = self.workspace_bytes(batch)
        if self._ws is None or self._ws.numel() < need + 1024:
            # growing replaces the buffer: anything that baked the old address in (a captured CUDA graph) is stale
            # from here on — `workspace_generation` lets such holders notice; reserve_workspace() avoids it up front
            self._ws = None
            self._ws = torch.empty(need + 1024, dtype=torch.uint8, device=self.device)
            self.workspace_generation = getattr(self, "workspace_generation", 0) + 1
        off = (-self._ws.data_ptr()) % 1024
        return self._ws.data_ptr() + off, need

    def reserve_workspace(self, max_batch: int) -> int:
        """Size the cached workspace for `max_batch` chunks now, so that later calls up to that batch never reallocate
        it (required before capturing CUDA graphs of encode()).  Returns the generation counter of the buffer."""
        self._workspace(max_batch)
        return getattr(self, "workspace_generation", 0)

    def encode(self, input_features, out_dtype=None, time_major_ld: int | None = None):
        """[B, n_mels, 3000] float32 (numpy / torch, host or device) -> [B, 1500, d] (bf16 by default) on the GPU.

        With `time_major_ld`, `input_features` is instead the bf16 [B, 3000, ld] tensor produced by
        `B200WhisperFeatureExtractor.extract(..., return_time_major=True)`."""
        import torch

        c = self.config
        out_dtype = torch.bfloat16 if out_dtype is None else out_dtype
        if out_dtype not in (torch.bfloat16, torch.float32):
            raise _lib.TtasrError(-1, "out_dtype must be torch.bfloat16 or torch.float32")
        t_in = 2 * c.max_source_positions
        if time_major_ld is None:
            x = torch.as_tensor(np.asarray(input_features) if not isinstance(input_features, torch.Tensor)
                                else input_features)
            if x.dim() == 2:
                x = x.unsqueeze(0)
            if x.dim() != 3 or x.shape[1] != c.num_mel_bins:
                raise _lib.TtasrError(-2, f"input_features must be [B, {c.num_mel_bins}, {t_in}], got {tuple(x.shape)}")
            if x.shape[-1] != t_in:
                raise ValueError(
                    f"Whisper expects the mel input features to be of length {t_in}, but found {x.shape[-1]}. "
                    f"Make sure to pad the input mel features to {t_in}.")
            x = x.to(self.device, torch.float32).contiguous()
            layout, ld = _lib.FEATS_F32_MEL_MAJOR, 0
        else:
            x = _lib.require_cuda_tensor(input_features, "input_features")
            if x.dtype != torch.bfloat16 or x.dim() != 3 or x.shape[1] != t_in or x.shape[2] != time_major_ld:
                raise _lib.TtasrError(-2, f"time-major features must be bf16 [B, {t_in}, {time_major_ld}]")
            if x.device != self.device:
                raise _lib.TtasrError(-1, f"time-major features live on {x.device} but this encoder's weights and "
                                          f"workspace are on {self.device} (one encoder object per GPU)")
            layout, ld = _lib.FEATS_BF16_TIME_MAJOR, time_major_ld
        B = x.shape[0]
        with torch.cuda.device(self.device):
            out = torch.empty((B, c.max_source_positions, c.d_model), dtype=out_dtype, device=self.device)
            if B == 0:
                return out
            ws_ptr, ws_bytes = self._workspace(B)
            _lib.check(_lib.lib().ttasr_encoder_forward(
                self._handle, x.data_ptr(), layout, ld, B, ws_ptr, ws_bytes, out.data_ptr(),
                _lib.OUT_F32 if out_dtype == torch.float32 else _lib.OUT_BF16, _lib.current_stream_ptr(self.device)))
        return out

    def forward(self, input_features, attention_mask=None, **kwargs):
        """Signature of WhisperEncoder.forward; `attention_mask` is accepted and ignored exactly as upstream does
        (modeling_whisper.py:608-611)."""
        hidden = self.encode(input_features, out_dtype=kwargs.pop("out_dtype", None))
        try:
            from transformers.modeling_outputs import BaseModelOutput

            return BaseModelOutput(last_hidden_state=hidden)
        except Exception:  # pragma: no cover
            return _Output(last_hidden_state=hidden)

    __call__ = forward
