"""Encoder weights from the model directory the reference deploys (REF api/stt_streaming/src/asr/faster_whisper_asr.py:
15-60 resolves `<project root>/<model_size>/{model.bin, config.json, tokenizer.json}` and hands it to
`faster_whisper.WhisperModel`; REF train_asr.py:539-545 loads the Hugging Face layout of the same model).

Two on-disk layouts are read, without importing ctranslate2 / faster-whisper / transformers:

* Hugging Face: `model.safetensors` (or sharded `model-0000x-of-0000y.safetensors` + index, or `pytorch_model.bin`)
  with `model.encoder.*` names — passed to `B200WhisperEncoder` as is.
* CTranslate2: `model.bin`, the converter's flat binary (ctranslate2/specs/model_spec.py `_serialize`, read back by
  ctranslate2/src/models/model.cc), little-endian:

      u32 binary_version (>= 4 here; current 6) | str spec_name | u32 spec_revision | u32 n_variables
      n_variables x { str name | u8 rank | rank x u32 dims | u8 dtype_id | u32 n_bytes | raw data }
      u32 n_aliases | n_aliases x { str alias | str variable_name }
      str = u16 length INCLUDING the terminating NUL | bytes | NUL
      dtype_id (ctranslate2/include/ctranslate2/types.h): 0 float32, 1 int8, 2 int16, 3 int32, 4 float16, 5 bfloat16

  Whisper encoder variables (ctranslate2/specs/whisper_spec.py, transformer_spec.py): `encoder/conv{1,2}/{weight,bias}`,
  `encoder/position_encodings/encodings`, `encoder/layer_norm/{gamma,beta}`, and per block
  `encoder/layer_<i>/self_attention/{layer_norm/{gamma,beta}, linear_0/{weight,bias} (fused q|k|v, zero k bias),
  linear_1/{weight,bias}}`, `encoder/layer_<i>/ffn/{layer_norm/{gamma,beta}, linear_0/{weight,bias},
  linear_1/{weight,bias}}`.  int8 / int16 weights carry a per-row `weight_scale` (stored value = weight * scale).

ctranslate2 is not installable offline, so this reader is checked against a synthetic `model.bin` written by the tests
from the layout above (tests/test_model_dir.py), not against the converter itself: parity with real converter output is
unpinned and said so in DESIGN.md.
"""
from __future__ import annotations

import json
import os
import struct

import numpy as np

from .encoder import EncoderConfig

CT2_DTYPES = {0: np.dtype("<f4"), 1: np.dtype("i1"), 2: np.dtype("<i2"), 3: np.dtype("<i4"), 4: np.dtype("<f2"),
              5: "bfloat16"}


class ModelDirError(ValueError):
    pass


# ------------------------------------------------------------------------------------------------ CTranslate2
def read_ct2_model_bin(path: str, prefix: str = "encoder/"):
    """-> (spec_name, spec_revision, {name: numpy array}, {alias: name}); only variables under `prefix` are
    materialised (the decoder's are skipped with a seek).  bfloat16 arrays come back as uint16 bit patterns; their names
    are listed in the extra entry variables["__bf16__"] (a set)."""
    variables, aliases, bf16 = {}, {}, set()
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        def take(fmt):
            n = struct.calcsize(fmt)
            buf = f.read(n)
            if len(buf) != n:
                raise ModelDirError(f"{path}: truncated file")
            return struct.unpack("<" + fmt, buf)

        def string():
            (n,) = take("H")
            raw = f.read(n)
            if len(raw) != n or n == 0 or raw[-1] != 0:
                raise ModelDirError(f"{path}: malformed string field")
            return raw[:-1].decode("utf-8")

        (version,) = take("I")
        if not 4 <= version <= 16:
            raise ModelDirError(f"{path}: unsupported CTranslate2 binary version {version} (need >= 4)")
        spec = string()
        (revision,) = take("I")
        (n_vars,) = take("I")
        for _ in range(n_vars):
            name = string()
            (rank,) = take("B")
            dims = take(f"{rank}I") if rank else ()
            (dtype_id,) = take("B")
            (n_bytes,) = take("I")
            if dtype_id not in CT2_DTYPES:
                raise ModelDirError(f"{path}: variable {name} has unknown dtype id {dtype_id}")
            if f.tell() + n_bytes > size:
                raise ModelDirError(f"{path}: variable {name} runs past the end of the file")
            if not name.startswith(prefix):
                f.seek(n_bytes, os.SEEK_CUR)
                continue
            dt = CT2_DTYPES[dtype_id]
            if dt == "bfloat16":
                arr = np.frombuffer(f.read(n_bytes), dtype="<u2")
                bf16.add(name)
            else:
                arr = np.frombuffer(f.read(n_bytes), dtype=dt)
            count = int(np.prod(dims)) if rank else 1
            if arr.size != count:
                raise ModelDirError(f"{path}: variable {name}: {arr.size} elements for shape {dims}")
            variables[name] = arr.reshape(dims)
        if f.tell() < size:
            (n_alias,) = take("I")
            for _ in range(n_alias):
                alias = string()
                aliases[alias] = string()
    variables["__bf16__"] = bf16
    return spec, revision, variables, aliases


def ct2_encoder_state(variables: dict, aliases: dict | None = None):
    """CTranslate2 Whisper variables -> (EncoderConfig, Hugging Face encoder state dict of torch tensors)."""
    import torch

    bf16 = variables.get("__bf16__", set())
    aliases = aliases or {}

    def raw(name):
        name = aliases.get(name, name)
        if name not in variables:
            raise ModelDirError(f"CTranslate2 model has no variable {name}")
        return name, variables[name]

    def tensor(name):
        name, a = raw(name)
        if name in bf16:
            return torch.from_numpy((a.astype(np.uint32) << 16).view(np.float32).copy())
        return torch.from_numpy(np.ascontiguousarray(a)).to(torch.float32)

    def weight(prefix):
        name, a = raw(prefix + "/weight")
        w = tensor(prefix + "/weight")
        if a.dtype.kind == "i" and name not in bf16:       # quantised: stored = round(weight * scale), scale per row
            scale = tensor(prefix + "/weight_scale").reshape(-1, *([1] * (w.dim() - 1)))
            w = w / scale
        return w

    conv1 = weight("encoder/conv1")
    d, n_mels = int(conv1.shape[0]), int(conv1.shape[1])
    n_layers = 0
    while f"encoder/layer_{n_layers}/self_attention/linear_0/weight" in variables:
        n_layers += 1
    if n_layers == 0:
        raise ModelDirError("CTranslate2 model has no encoder/layer_0 (not a Whisper model?)")
    pos = tensor("encoder/position_encodings/encodings")
    sd = {"conv1.weight": conv1, "conv1.bias": tensor("encoder/conv1/bias"),
          "conv2.weight": weight("encoder/conv2"), "conv2.bias": tensor("encoder/conv2/bias"),
          "embed_positions.weight": pos,
          "layer_norm.weight": tensor("encoder/layer_norm/gamma"), "layer_norm.bias": tensor("encoder/layer_norm/beta")}
    ffn = 0
    for i in range(n_layers):
        src, dst = f"encoder/layer_{i}/", f"layers.{i}."
        qkv_w = weight(src + "self_attention/linear_0")
        qkv_b = tensor(src + "self_attention/linear_0/bias")
        if tuple(qkv_w.shape) != (3 * d, d):
            raise ModelDirError(f"layer {i}: fused in-projection has shape {tuple(qkv_w.shape)}, expected {(3 * d, d)}")
        for j, p in enumerate(("q_proj", "k_proj", "v_proj")):
            sd[dst + f"self_attn.{p}.weight"] = qkv_w[j * d:(j + 1) * d].contiguous()
            if p != "k_proj":
                sd[dst + f"self_attn.{p}.bias"] = qkv_b[j * d:(j + 1) * d].contiguous()
        sd[dst + "self_attn.out_proj.weight"] = weight(src + "self_attention/linear_1")
        sd[dst + "self_attn.out_proj.bias"] = tensor(src + "self_attention/linear_1/bias")
        sd[dst + "self_attn_layer_norm.weight"] = tensor(src + "self_attention/layer_norm/gamma")
        sd[dst + "self_attn_layer_norm.bias"] = tensor(src + "self_attention/layer_norm/beta")
        sd[dst + "final_layer_norm.weight"] = tensor(src + "ffn/layer_norm/gamma")
        sd[dst + "final_layer_norm.bias"] = tensor(src + "ffn/layer_norm/beta")
        sd[dst + "fc1.weight"] = weight(src + "ffn/linear_0")
        sd[dst + "fc1.bias"] = tensor(src + "ffn/linear_0/bias")
        sd[dst + "fc2.weight"] = weight(src + "ffn/linear_1")
        sd[dst + "fc2.bias"] = tensor(src + "ffn/linear_1/bias")
        ffn = int(sd[dst + "fc1.weight"].shape[0])
    if d % 64 != 0:
        raise ModelDirError(f"d_model {d} is not a multiple of the head dimension 64")
    cfg = EncoderConfig(d_model=d, encoder_layers=n_layers, encoder_attention_heads=d // 64, encoder_ffn_dim=ffn,
                        num_mel_bins=n_mels, max_source_positions=int(pos.shape[0]))
    return cfg, sd


# ------------------------------------------------------------------------------------------------ Hugging Face
def _hf_state(model_dir: str):
    from safetensors import safe_open

    names = sorted(n for n in os.listdir(model_dir) if n.endswith(".safetensors"))
    if names:
        sd = {}
        for n in names:
            with safe_open(os.path.join(model_dir, n), framework="pt", device="cpu") as f:
                for k in f.keys():
                    if k.startswith(("model.encoder.", "encoder.")):
                        sd[k] = f.get_tensor(k)
        return sd
    path = os.path.join(model_dir, "pytorch_model.bin")
    if os.path.exists(path):
        import torch

        full = torch.load(path, map_location="cpu", weights_only=True)
        return {k: v for k, v in full.items() if k.startswith(("model.encoder.", "encoder."))}
    return None


def load_encoder_weights(model_dir: str):
    """-> (EncoderConfig, state dict with Hugging Face encoder names, "hf" | "ct2").  Prefers the Hugging Face files
    when both layouts are present (they hold the fine-tuned weights before conversion)."""
    if not os.path.isdir(model_dir):
        raise FileNotFoundError(f"model directory {model_dir} does not exist")
    cfg_path = os.path.join(model_dir, "config.json")
    hf = _hf_state(model_dir)
    if hf:
        if not os.path.exists(cfg_path):
            raise FileNotFoundError(f"{cfg_path} is missing")
        with open(cfg_path) as f:
            return EncoderConfig.from_any(json.load(f)), hf, "hf"
    bin_path = os.path.join(model_dir, "model.bin")
    if not os.path.exists(bin_path):
        raise FileNotFoundError(f"{model_dir} holds neither *.safetensors / pytorch_model.bin nor a CTranslate2 model.bin")
    _, _, variables, aliases = read_ct2_model_bin(bin_path)
    cfg, sd = ct2_encoder_state(variables, aliases)
    return cfg, sd, "ct2"


def resolve_model_dir(model_size: str, search_roots=()):
    """The reference's rule (faster_whisper_asr.py:24-49): `<project root>/<model_size>` if it exists, where the project
    root is two levels above `api/stt_streaming`; here the caller passes the candidate roots (plus $TTASR_MODEL_ROOT and
    the working directory).  An existing path is taken as is.  Returns None when nothing is found — the reference would
    then let faster-whisper download `model_size` from the hub, which an offline B200 deployment cannot do."""
    if os.path.isdir(model_size):
        return os.path.abspath(model_size)
    roots = list(search_roots)
    if os.environ.get("TTASR_MODEL_ROOT"):
        roots.append(os.environ["TTASR_MODEL_ROOT"])
    roots.append(os.getcwd())
    for r in roots:
        cand = os.path.join(r, model_size)
        if os.path.isdir(cand):
            return os.path.abspath(cand)
    return None
