"""Weights from the model directory the reference deploys (ttasr/model_dir.py): the CTranslate2 `model.bin` reader is
checked against a synthetic file written here from the documented layout (ctranslate2 itself is not installable
offline — parity with the real converter is unpinned), the Hugging Face layout through safetensors."""
import json
import os
import struct

import numpy as np
import pytest

from oracle import encoder as OE


def _wstr(f, s):
    b = s.encode("utf-8")
    f.write(struct.pack("<H", len(b) + 1))
    f.write(b)
    f.write(b"\0")


def write_ct2_model_bin(path, variables, aliases=(), spec="WhisperSpec", revision=3, version=6):
    """ctranslate2/specs/model_spec.py `_serialize`, restated: see the layout in ttasr/model_dir.py's docstring."""
    ids = {np.dtype("float32"): 0, np.dtype("int8"): 1, np.dtype("int16"): 2, np.dtype("int32"): 3,
           np.dtype("float16"): 4}
    with open(path, "wb") as f:
        f.write(struct.pack("<I", version))
        _wstr(f, spec)
        f.write(struct.pack("<I", revision))
        f.write(struct.pack("<I", len(variables)))
        for name, (arr, dtype_id) in variables.items():
            _wstr(f, name)
            f.write(struct.pack("<B", arr.ndim))
            for d in arr.shape:
                f.write(struct.pack("<I", d))
            f.write(struct.pack("<B", ids[arr.dtype] if dtype_id is None else dtype_id))
            raw = np.ascontiguousarray(arr).tobytes()
            f.write(struct.pack("<I", len(raw)))
            f.write(raw)
        f.write(struct.pack("<I", len(aliases)))
        for a, b in aliases:
            _wstr(f, a)
            _wstr(f, b)


def ct2_variables_from_hf(w, arch, storage="float16"):
    """HF encoder state dict -> CTranslate2 Whisper variable names (whisper_spec.py / transformer_spec.py)."""
    import torch

    out = {}

    def put(name, t, matrix=False):
        a = t.detach().numpy().astype(np.float32)
        if matrix and storage == "int8":
            scale = 127.0 / np.abs(a).reshape(a.shape[0], -1).max(axis=1)
            q = np.round(a * scale.reshape(-1, *([1] * (a.ndim - 1)))).astype(np.int8)
            out[name] = (q, None)
            out[name + "_scale"] = (scale.astype(np.float32), None)
        elif storage == "bfloat16":
            bits = (torch.from_numpy(a).to(torch.bfloat16).view(torch.int16).numpy().astype(np.uint16))
            out[name] = (bits.view(np.int16), 5)
        elif storage == "float16":
            out[name] = (a.astype(np.float16), None)
        else:
            out[name] = (a, None)

    put("encoder/conv1/weight", w["conv1.weight"], True)
    put("encoder/conv1/bias", w["conv1.bias"])
    put("encoder/conv2/weight", w["conv2.weight"], True)
    put("encoder/conv2/bias", w["conv2.bias"])
    put("encoder/position_encodings/encodings", w["embed_positions.weight"])
    put("encoder/layer_norm/gamma", w["layer_norm.weight"])
    put("encoder/layer_norm/beta", w["layer_norm.bias"])
    d = arch.d_model
    for i in range(arch.layers):
        p, q = f"layers.{i}.", f"encoder/layer_{i}/"
        qkv = torch.cat([w[p + f"self_attn.{n}.weight"] for n in ("q_proj", "k_proj", "v_proj")])
        bias = torch.cat([w[p + "self_attn.q_proj.bias"], torch.zeros(d), w[p + "self_attn.v_proj.bias"]])
        put(q + "self_attention/linear_0/weight", qkv, True)
        put(q + "self_attention/linear_0/bias", bias)
        put(q + "self_attention/linear_1/weight", w[p + "self_attn.out_proj.weight"], True)
        put(q + "self_attention/linear_1/bias", w[p + "self_attn.out_proj.bias"])
        put(q + "self_attention/layer_norm/gamma", w[p + "self_attn_layer_norm.weight"])
        put(q + "self_attention/layer_norm/beta", w[p + "self_attn_layer_norm.bias"])
        put(q + "ffn/layer_norm/gamma", w[p + "final_layer_norm.weight"])
        put(q + "ffn/layer_norm/beta", w[p + "final_layer_norm.bias"])
        put(q + "ffn/linear_0/weight", w[p + "fc1.weight"], True)
        put(q + "ffn/linear_0/bias", w[p + "fc1.bias"])
        put(q + "ffn/linear_1/weight", w[p + "fc2.weight"], True)
        put(q + "ffn/linear_1/bias", w[p + "fc2.bias"])
    # something that is not the encoder's: must be skipped without being materialised
    out["decoder/embeddings/weight"] = (np.zeros((7, d), np.float16), None)
    return out


@pytest.mark.parametrize("storage,tol", [("float32", 0.0), ("float16", 1e-3), ("bfloat16", 8e-3), ("int8", 1e-2)])
def test_ct2_model_bin_round_trip(tmp_path, storage, tol):
    from ttasr.model_dir import ct2_encoder_state, load_encoder_weights, read_ct2_model_bin

    arch = OE.ARCHS["micro"]
    w = OE.init_weights(arch, seed=2, ln_jitter=0.05)
    path = tmp_path / "model.bin"
    write_ct2_model_bin(str(path), ct2_variables_from_hf(w, arch, storage),
                        aliases=[("encoder/alias_of_final_norm", "encoder/layer_norm/gamma")])
    spec, rev, variables, aliases = read_ct2_model_bin(str(path))
    assert spec == "WhisperSpec" and rev == 3 and aliases == {"encoder/alias_of_final_norm": "encoder/layer_norm/gamma"}
    assert not any(k.startswith("decoder/") for k in variables)
    cfg, sd = ct2_encoder_state(variables, aliases)
    assert (cfg.d_model, cfg.encoder_layers, cfg.encoder_attention_heads, cfg.encoder_ffn_dim, cfg.num_mel_bins,
            cfg.max_source_positions) == (arch.d_model, arch.layers, arch.heads, arch.ffn, arch.n_mels, arch.n_ctx)
    assert set(sd) == set(w)
    for k in w:
        scale = float(w[k].abs().max()) + 1e-6
        err = float((sd[k] - w[k]).abs().max()) / scale
        assert err <= tol, (k, err)
    cfg2, sd2, fmt = load_encoder_weights(str(tmp_path))
    assert fmt == "ct2" and cfg2 == cfg


def test_hf_directory_is_preferred_and_parsed(tmp_path):
    import torch
    from safetensors.torch import save_file
    from ttasr.model_dir import load_encoder_weights, resolve_model_dir

    arch = OE.ARCHS["micro"]
    w = OE.init_weights(arch, seed=4)
    sd = {"model.encoder." + k: v.contiguous() for k, v in w.items()}
    sd["model.decoder.embed_tokens.weight"] = torch.zeros(3, arch.d_model)
    mdir = tmp_path / "my-finetune"
    mdir.mkdir()
    save_file(sd, str(mdir / "model.safetensors"))
    (mdir / "config.json").write_text(json.dumps(dict(
        d_model=arch.d_model, encoder_layers=arch.layers, encoder_attention_heads=arch.heads,
        encoder_ffn_dim=arch.ffn, num_mel_bins=arch.n_mels, max_source_positions=arch.n_ctx)))
    write_ct2_model_bin(str(mdir / "model.bin"), {})           # a (broken) CT2 file next to it must not be touched
    cfg, got, fmt = load_encoder_weights(str(mdir))
    assert fmt == "hf" and cfg.d_model == arch.d_model
    assert set(got) == {"model.encoder." + k for k in w}
    assert resolve_model_dir("my-finetune", [str(tmp_path)]) == str(mdir)
    assert resolve_model_dir(str(mdir)) == str(mdir)
    assert resolve_model_dir("large-v3-turbo", [str(tmp_path)]) is None


def test_bad_files_are_rejected(tmp_path):
    from ttasr.model_dir import ModelDirError, load_encoder_weights, read_ct2_model_bin

    p = tmp_path / "model.bin"
    p.write_bytes(struct.pack("<I", 2) + b"\0" * 16)
    with pytest.raises(ModelDirError):
        read_ct2_model_bin(str(p))
    arch = OE.ARCHS["micro"]
    good = tmp_path / "good.bin"
    write_ct2_model_bin(str(good), ct2_variables_from_hf(OE.init_weights(arch, seed=1), arch, "float16"))
    raw = good.read_bytes()
    p.write_bytes(raw[: len(raw) // 2])
    with pytest.raises(ModelDirError):
        read_ct2_model_bin(str(p))
    with pytest.raises(FileNotFoundError):
        load_encoder_weights(str(tmp_path / "nope"))
    os.remove(p)
    os.remove(good)
    with pytest.raises(FileNotFoundError):
        load_encoder_weights(str(tmp_path))


def test_pause_aligned_cuts():
    from ttasr.ingest import pause_aligned_cuts

    rng = np.random.default_rng(0)
    db = np.full(9000, -20.0) + rng.standard_normal(9000)          # 90 s of "speech" at 10 ms hops
    for c in (2500, 2950, 5800, 5950, 8000):
        db[c - 15: c + 15] = -70.0                                  # 300 ms pauses
    db[4000:4005] = -70.0                                           # 50 ms dip: not a pause
    cuts = pause_aligned_cuts(db)
    assert cuts == [(0, 2950), (2950, 5950), (5950, 8000), (8000, 9000)]
    assert all(b - a <= 3000 for a, b in cuts) and cuts[0][0] == 0 and cuts[-1][1] == 9000
    assert all(cuts[i][1] == cuts[i + 1][0] for i in range(len(cuts) - 1))      # tiles the recording
    # no pause at all -> hard cuts every 30 s, as round 1 always did
    flat = np.full(7000, -20.0)
    assert pause_aligned_cuts(flat) == [(0, 3000), (3000, 6000), (6000, 7000)]
    assert pause_aligned_cuts(np.zeros(0)) == []
    assert pause_aligned_cuts(np.full(100, -30.0)) == [(0, 100)]
