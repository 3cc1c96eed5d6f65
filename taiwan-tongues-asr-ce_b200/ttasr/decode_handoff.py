"""Hand-off of the encoder output to the decoder that already exists in the host application, without leaving the
GPU.

  * Hugging Face (train_asr.py evaluation with predict_with_generate, train_asr.py:697-716,736-740):
    `WhisperForConditionalGeneration.generate(encoder_outputs=BaseModelOutput(last_hidden_state=h))` skips the model's
    own encoder (generation_whisper.py:1327-1335,1613-1654) — `hf_generate` builds that call from PCM or features.
  * CTranslate2 / faster-whisper (asr_core.py:159-167, api/file_asr.py:457-465, faster_whisper_asr.py:170-172):
    `WhisperModel.encode` must return a `ctranslate2.StorageView`; `to_storage_view` wraps the CUDA tensor through
    `__cuda_array_interface__` (fp16, the dtype CT2 decodes in on CUDA) — pass it as `wrap=` to
    `compat_faster_whisper.patch_model`.

Decoding itself (beam search, timestamps, text post-processing) stays with the host application."""
from __future__ import annotations


def encoder_outputs(hidden, dtype=None):
    """[B, 1500, d] CUDA tensor -> transformers BaseModelOutput in the decoder's dtype."""
    from transformers.modeling_outputs import BaseModelOutput

    return BaseModelOutput(last_hidden_state=hidden if dtype is None else hidden.to(dtype))


def hf_generate(model, pipeline, pcm=None, input_features=None, n_valid=None, **generate_kwargs):
    """model: a transformers WhisperForConditionalGeneration on the same CUDA device as `pipeline`
    (ttasr.B200LogMelEncoder).  Give either `pcm` ([B, 480000] CUDA, float32 / int16) or `input_features`
    ([B, n_mels, 3000]).  Returns what `model.generate` returns."""
    if (pcm is None) == (input_features is None):
        raise ValueError("give exactly one of pcm / input_features")
    if pcm is not None:
        hidden = pipeline.encode_device(pcm, n_valid=n_valid)
    else:
        hidden = pipeline.encoder.encode(input_features)
    dtype = next(model.parameters()).dtype
    return model.generate(encoder_outputs=encoder_outputs(hidden, dtype), **generate_kwargs)


def to_storage_view(hidden):
    """CUDA tensor [B, 1500, d] -> ctranslate2.StorageView sharing the memory (fp16 copy when the tensor is bf16:
    CTranslate2 has no bf16 StorageView constructor from the array interface).  Raises ImportError when ctranslate2
    is not installed — there is no substitute."""
    import ctranslate2
    import torch

    if hidden.dtype == torch.bfloat16:
        hidden = hidden.to(torch.float16)
    return ctranslate2.StorageView.from_array(hidden.contiguous())
