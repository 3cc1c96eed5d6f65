"""Encoder parity through the drop-in Python surface -> C ABI -> sm_100a kernels, against the fp32 oracle on the same
bf16-rounded weights.

Gates, frozen in round 2 at about twice the error measured on the B200 for the default (split residual stream) path with
bf16 output (profiles/r2_parity_report.json; max-abs, a single-element statistic, gets 2.5x): per architecture, since the
error grows with depth.  For scale, Hugging Face's own encoder run in bf16 on the same GPU (the control run SURVEY 8d asks
for) sits at 0.146 / 0.0109 / 0.99990 for large-v3 — five times further from the fp32 oracle than this library."""
import os

import numpy as np
import pytest

from oracle import encoder as OE
from oracle import frontend as OF

pytestmark = pytest.mark.gpu
GATES = {  # measured (split, bf16 out): micro ~tiny; tiny 0.013 / 0.0017 / 0.9999975; small 0.022 / 0.0030 / 0.9999930;
    #           large-v3 0.029 / 0.0040 / 0.9999875
    "micro": (0.035, 0.0035, 0.999995), "tiny": (0.035, 0.0035, 0.999995), "small": (0.055, 0.006, 0.999985),
    "large-v3": (0.07, 0.008, 0.999975),
    # synthetic stress streams (massive-activation channels, token-mean offset): the round-1 provisional gate
    "stress": (0.10, 0.01, 0.9999),
}
MAX_ABS, MEAN_ABS, COS = GATES["large-v3"]   # the loosest: for checks that do not name their architecture


def _cfg(arch):
    return dict(d_model=arch.d_model, encoder_layers=arch.layers, encoder_attention_heads=arch.heads,
                encoder_ffn_dim=arch.ffn, num_mel_bins=arch.n_mels, max_source_positions=arch.n_ctx)


def _build(arch_name, seed=0):
    from ttasr import B200WhisperEncoder

    arch = OE.ARCHS[arch_name]
    w = OE.round_weights_bf16(OE.init_weights(arch, seed=seed, ln_jitter=0.02))
    return arch, w, B200WhisperEncoder(_cfg(arch), w)


def _check(got, ref, arch="large-v3"):
    s = OE.parity_stats(got, ref)
    max_abs, mean_abs, cos = GATES[arch]
    assert s["max_abs"] <= max_abs and s["mean_abs"] <= mean_abs and s["cosine"] >= cos, (arch, s)
    return s


@pytest.mark.parametrize("arch_name", ["micro", "tiny"])
def test_encoder_matches_fp32_oracle(cuda_device, arch_name):
    import torch

    arch, w, enc = _build(arch_name)
    feats = np.stack([OF.log_mel(OF.synth_noise(), arch.n_mels), OF.log_mel(OF.synth_tones(), arch.n_mels),
                      OF.log_mel(OF.synth_short(), arch.n_mels)])
    ref = OE.encoder_forward(torch.from_numpy(feats), w, arch).numpy()
    got = enc.encode(feats, out_dtype=torch.float32).cpu().numpy()
    assert got.shape == (3, 1500, arch.d_model)
    _check(got, ref, arch_name)
    got16 = enc.encode(feats)
    assert got16.dtype == torch.bfloat16
    _check(got16.float().cpu().numpy(), ref, arch_name)
    # same chunk, other batch position -> identical bits (SURVEY.md section 8e determinism)
    again = enc.encode(feats[[2, 0]], out_dtype=torch.float32).cpu().numpy()
    assert np.array_equal(again[1], got[0]) and np.array_equal(again[0], got[2])


def test_golden_fixture_tiny(cuda_device):
    import torch

    arch, w, enc = _build("tiny")
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "encoder_hf.npz"))
    feats = OF.log_mel(OF.synth_noise(), arch.n_mels)[None]
    got = enc.encode(feats, out_dtype=torch.float32)[0].cpu().numpy()
    _check(got[::25], g["enc_tiny_sub"], "tiny")


def test_forward_is_a_drop_in_for_hf_encoder(cuda_device):
    """encoder(input_features) -> .last_hidden_state, usable as encoder_outputs by the reference's decode code."""
    import torch

    arch, w, enc = _build("micro")
    feats = torch.from_numpy(OF.log_mel(OF.synth_noise(), arch.n_mels)[None])
    out = enc(feats, attention_mask=None)
    assert out.last_hidden_state.shape == (1, 1500, arch.d_model) and out[0] is out.last_hidden_state
    with pytest.raises(ValueError):
        enc(feats[..., :2000])


def test_pcm_to_hidden_pipeline(cuda_device):
    import torch
    from ttasr import B200LogMelEncoder, B200WhisperFeatureExtractor

    arch, w, enc = _build("tiny")
    pipe = B200LogMelEncoder(B200WhisperFeatureExtractor(feature_size=arch.n_mels), enc)
    clips = np.stack([OF.pad_or_trim(OF.synth_noise(3)), OF.pad_or_trim(OF.synth_short(4))])
    got = pipe.encode_device(torch.from_numpy(clips).to(cuda_device), out_dtype=torch.float32).cpu().numpy()
    feats = np.stack([OF.log_mel(c, arch.n_mels) for c in clips])
    ref = OE.encoder_forward(torch.from_numpy(feats), w, arch).numpy()
    _check(got, ref, "tiny")
    host = torch.from_numpy(clips).pin_memory()
    out_host = torch.empty((2, 1500, arch.d_model), dtype=torch.bfloat16).pin_memory()
    pipe.encode_host(host, out_host)
    torch.cuda.synchronize()
    _check(out_host.float().numpy(), ref, "tiny")


@pytest.mark.parametrize("arch_name", ["small"])
def test_encoder_small_config2_sample(cuda_device, arch_name):
    """Config 2 architecture (whisper-small, 80-bin) on two chunks; the oracle needs ~2 s per chunk on CPU."""
    import torch

    arch, w, enc = _build(arch_name)
    feats = np.stack([OF.log_mel(OF.synth_noise(11), arch.n_mels), OF.log_mel(OF.synth_tones(12), arch.n_mels)])
    ref = OE.encoder_forward(torch.from_numpy(feats), w, arch).numpy()
    got = enc.encode(feats, out_dtype=torch.float32).cpu().numpy()
    _check(got, ref, arch_name)


def test_pipelined_host_stream_matches_single_calls(cuda_device):
    """stream_host overlaps copy-in / kernels / copy-out over a batch sequence; results equal the unpipelined path."""
    import torch
    from ttasr import B200LogMelEncoder, B200WhisperFeatureExtractor

    arch, w, enc = _build("micro")
    pipe = B200LogMelEncoder(B200WhisperFeatureExtractor(feature_size=arch.n_mels), enc)
    batches = [torch.from_numpy(np.stack([OF.pad_or_trim(OF.synth_noise(20 + 2 * k)),
                                          OF.pad_or_trim(OF.synth_tones(21 + 2 * k))])).pin_memory() for k in range(5)]
    outs = [torch.empty((2, 1500, arch.d_model), dtype=torch.bfloat16).pin_memory() for _ in batches]
    pipe.stream_host(batches, outs)
    torch.cuda.synchronize()
    for b, o in zip(batches, outs):
        ref = pipe.encode_device(b.to(cuda_device)).cpu()
        assert torch.equal(o, ref)


@pytest.mark.parametrize("use_graphs", [False, True])
def test_streaming_plugin_end_to_end(cuda_device, use_graphs):
    """B200ASR (the reference's ASRInterface): int16 scratch buffers of several clients -> one batched launch ->
    hidden states identical to encoding each utterance alone; decode stays with the host callback."""
    import asyncio
    import types
    import torch
    from ttasr import B200LogMelEncoder, B200WhisperFeatureExtractor
    from ttasr.asr_plugin import B200ASR

    arch, w, enc = _build("micro")
    pipe = B200LogMelEncoder(B200WhisperFeatureExtractor(feature_size=arch.n_mels), enc)
    seen = {}

    def decode(hidden, info):
        seen[info["n_samples"]] = hidden.float().cpu()
        return {"text": "測試", "words": []}

    asr = B200ASR(pipe, decode, batch_window_s=0.02, max_batch=8, use_graphs=use_graphs)
    assert asr.warm_up()["hidden_shape"] == (1, 1500, arch.d_model)
    rng = np.random.default_rng(8)
    pcm = [(rng.standard_normal(n) * 2000).astype("<i2") for n in (16000, 52345, 80000)]
    clients = [types.SimpleNamespace(scratch_buffer=bytearray(p.tobytes()), samples_width=2, last_start_time=0.0,
                                     client_id=i) for i, p in enumerate(pcm)]

    async def run():
        return await asyncio.gather(*(asr.transcribe(c) for c in clients))

    launches0 = asr.batcher.launches
    results = asyncio.run(run())
    assert all(r and r["text"] == "測試" and r["final"] for r in results)
    assert asr.batcher.launches == launches0 + 1  # three clients, one launch
    for p in pcm:
        feats = OF.log_mel(p.astype(np.float32) / 32768.0, arch.n_mels)[None]
        ref = OE.encoder_forward(torch.from_numpy(feats), w, arch)
        _check(seen[len(p)].numpy(), ref.numpy(), "micro")


def test_faster_whisper_seams(cuda_device):
    """compat_faster_whisper: whole-file log-mel with the global clamp + encode() of a padded window."""
    import types
    import torch
    from ttasr import B200WhisperFeatureExtractor
    from ttasr.compat_faster_whisper import patch_model

    arch, w, enc = _build("micro")
    model = patch_model(types.SimpleNamespace(), enc, B200WhisperFeatureExtractor(feature_size=arch.n_mels))
    rng = np.random.default_rng(4)
    wave = (0.1 * rng.standard_normal(16000 * 41 + 77)).astype(np.float32)
    feats = model.feature_extractor(wave, padding=160)
    padded = np.concatenate([wave, np.zeros(160, np.float32)])
    ref = OF.log_mel_unclamped(padded, arch.n_mels)[:, :-1]
    ref = (np.maximum(ref, ref.max() - 8.0) + 4.0) / 4.0
    assert feats.shape == ref.shape and np.abs(feats - ref).max() <= 1e-4
    window = feats[:, 3000:]  # second 30 s window is short: encode() pads it in feature space like faster-whisper
    hidden = model.encode(window)
    padded_window = np.concatenate([window, np.zeros((arch.n_mels, 3000 - window.shape[1]), np.float32)], axis=1)
    ref_h = OE.encoder_forward(torch.from_numpy(padded_window[None]), w, arch)
    _check(hidden.float().cpu().numpy(), ref_h.numpy(), "micro")


def test_large_v3_single_chunk_against_oracle(cuda_device):
    """The headline architecture (config 3 shape: d=1280, 32 layers, 20 heads, 128 mels) on one chunk; the fp32 oracle
    takes a few seconds on the host.  Gate as stated in DESIGN.md section 2."""
    import torch

    arch, w, enc = _build("large-v3")
    feats = OF.log_mel(OF.synth_tones(31), arch.n_mels)[None]
    ref = OE.encoder_forward(torch.from_numpy(feats), w, arch).numpy()
    got = enc.encode(feats, out_dtype=torch.float32).cpu().numpy()
    s = _check(got, ref, "large-v3")
    print("large-v3 parity:", s)


def test_forward_can_be_captured_in_a_cuda_graph(cuda_device):
    """Everything behind ttasr_frontend_run / ttasr_encoder_forward is stream-ordered with no host synchronisation,
    so the latency-bound streaming shape (one utterance, config 5) can be replayed as a CUDA graph."""
    import torch
    from ttasr import B200LogMelEncoder, B200WhisperFeatureExtractor

    arch, w, enc = _build("tiny")
    pipe = B200LogMelEncoder(B200WhisperFeatureExtractor(feature_size=arch.n_mels), enc)
    pcm = torch.from_numpy(OF.pad_or_trim(OF.synth_noise(41))[None]).to(cuda_device)
    eager = pipe.encode_device(pcm).clone()
    static_in = pcm.clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            pipe.encode_device(static_in)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static_out = pipe.encode_device(static_in)
    static_in.copy_(torch.from_numpy(OF.pad_or_trim(OF.synth_tones(42))[None]).to(cuda_device))
    graph.replay()
    torch.cuda.synchronize()
    other = pipe.encode_device(static_in)
    assert torch.equal(static_out, other) and not torch.equal(static_out, eager)
    static_in.copy_(pcm)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(static_out, eager)


def test_full_size_properties_large_v3_b256(cuda_device):
    """BASELINE.json configs[2] at full size (large-v3, 256 x 30 s): the oracle cannot run here, so the path is checked
    through size-independent properties — determinism (two runs bit-identical), batch-position invariance (a chunk's
    hidden states do not depend on where it sits in the batch or on the batch size), finiteness, and the
    all-padding chunk equal to the all-zero chunk."""
    import torch
    import bench
    import ttasr

    dev = cuda_device
    cfg = ttasr.EncoderConfig.named("large-v3")
    fe = ttasr.B200WhisperFeatureExtractor(feature_size=cfg.num_mel_bins)
    enc = ttasr.B200WhisperEncoder(cfg, bench.make_gpu_weights(cfg, dev))
    pipe = ttasr.B200LogMelEncoder(fe, enc)
    B = 256
    g = torch.Generator(device=dev).manual_seed(1234)
    pcm = (0.1 * torch.randn((B, 480000), device=dev, generator=g)).clamp_(-1, 1)
    pcm[7] = 0.0                                   # an all-zero chunk ...
    nv = torch.full((B,), 480000, dtype=torch.int32, device=dev)
    nv[9] = 0                                      # ... and an all-padding one
    h1 = pipe.encode_device(pcm, n_valid=nv)
    h2 = pipe.encode_device(pcm, n_valid=nv)
    assert h1.shape == (B, 1500, cfg.d_model) and h1.dtype == torch.bfloat16
    assert bool(torch.isfinite(h1.float()).all())
    assert torch.equal(h1, h2), "two runs of the same batch differ"
    assert torch.equal(h1[7], h1[9]), "all-zero chunk and all-padding chunk differ"
    rows = [0, 100, 255, 7]
    sub = pipe.encode_device(pcm[rows].contiguous(), n_valid=nv[rows].contiguous())
    for k, r in enumerate(rows):
        assert torch.equal(sub[k], h1[r]), f"chunk {r}: result depends on batch position / batch size"
    # a permutation of the batch permutes the output
    perm = torch.randperm(B, device=dev, generator=g)
    hp = pipe.encode_device(pcm[perm].contiguous(), n_valid=nv[perm].contiguous())
    assert torch.equal(hp, h1[perm])


def test_graph_buckets_match_eager_for_ragged_micro_batches(cuda_device):
    """Serving path: replaying one CUDA graph per batch-size bucket gives exactly the eager result for ragged int16
    utterances, across buckets, repeated replays and padded (empty) rows."""
    import torch
    from ttasr import B200LogMelEncoder, B200WhisperFeatureExtractor, GraphedLogMelEncoder

    arch, w, enc = _build("micro")
    pipe = B200LogMelEncoder(B200WhisperFeatureExtractor(feature_size=arch.n_mels), enc)
    graphed = GraphedLogMelEncoder(pipe, buckets=(1, 2, 4))
    rng = np.random.default_rng(3)
    for lens in ([16000], [80000, 24000, 50001], [48000, 16000], [1], [33333, 480000, 7, 160]):
        rows = [torch.from_numpy((rng.standard_normal(n) * 3000).astype(np.int16)) for n in lens]
        got = graphed.encode(rows)
        width = (max(lens) + 7) // 8 * 8
        host = torch.zeros((len(lens), width), dtype=torch.int16)
        for i, r in enumerate(rows):
            host[i, : lens[i]] = r
        ref = pipe.encode_device(host.to(cuda_device), n_valid=torch.tensor(lens, dtype=torch.int32, device=cuda_device))
        assert got.shape == ref.shape and torch.equal(got, ref), lens
    assert graphed.replays == 5 and sorted(graphed._graphs) == [1, 2, 4]
    with pytest.raises(Exception):
        graphed.encode([torch.zeros(10, dtype=torch.int16)] * 5)


def test_graph_buckets_survive_workspace_growth(cuda_device):
    """A larger eager batch after capture replaces the encoder's cached workspace; the graph holder must notice and
    re-capture instead of replaying graphs that point at the freed buffer."""
    import torch
    from ttasr import B200LogMelEncoder, B200WhisperFeatureExtractor, GraphedLogMelEncoder

    arch, w, enc = _build("micro")
    pipe = B200LogMelEncoder(B200WhisperFeatureExtractor(feature_size=arch.n_mels), enc)
    graphed = GraphedLogMelEncoder(pipe, buckets=(1, 2))
    rng = np.random.default_rng(4)
    row = torch.from_numpy((rng.standard_normal(40000) * 3000).astype(np.int16))
    first = graphed.encode([row])
    gen = enc.workspace_generation
    big = torch.from_numpy((rng.standard_normal((6, 480000)) * 3000).astype(np.int16)).to(cuda_device)
    pipe.encode_device(big)                                  # grows the workspace past the reservation
    assert enc.workspace_generation > gen
    torch.empty(64 << 20, dtype=torch.uint8, device=cuda_device).fill_(0xAB)   # scribble over whatever was freed
    again = graphed.encode([row])
    assert torch.equal(first, again)


@pytest.mark.parametrize("arch_name", ["micro", "tiny"])
def test_residual_modes_match_oracle_and_each_other(cuda_device, arch_name):
    """The three representations of the residual stream (include/ttasr_abi.h: split = default, f32, bf16).  `split`
    folds the per-layer LayerNorms into the QKV / fc1 GEMMs (statistics emitted by the residual GEMMs' epilogues), so it
    launches two kernels fewer per layer; it and `f32` meet the same gate and agree with each other far inside it."""
    import torch
    from ttasr import B200WhisperEncoder

    arch = OE.ARCHS[arch_name]
    w = OE.round_weights_bf16(OE.init_weights(arch, seed=0, ln_jitter=0.02))
    enc = {m: B200WhisperEncoder(_cfg(arch), w, residual=m) for m in ("split", "f32", "bf16")}
    assert enc["split"].launches_per_forward == enc["f32"].launches_per_forward - 2 * arch.layers
    assert enc["bf16"].launches_per_forward == enc["split"].launches_per_forward
    feats = np.stack([OF.log_mel(OF.synth_noise(31), arch.n_mels), OF.log_mel(OF.synth_tones(32), arch.n_mels),
                      OF.log_mel(OF.pad_or_trim(OF.synth_short()), arch.n_mels)])
    ref = OE.encoder_forward(torch.from_numpy(feats), w, arch).numpy()
    out = {m: e.encode(feats, out_dtype=torch.float32) for m, e in enc.items()}
    for m in ("split", "f32"):
        print(arch_name, m, _check(out[m].cpu().numpy(), ref, arch_name))
    print(arch_name, "bf16", OE.parity_stats(out["bf16"].cpu(), ref))
    s = OE.parity_stats(out["split"].cpu(), out["f32"].cpu())
    assert s["max_abs"] <= 0.05 and s["cosine"] >= 0.99995, s
    for m, e in enc.items():
        assert torch.equal(out[m], e.encode(feats, out_dtype=torch.float32)), m             # deterministic
        assert torch.equal(out[m][1], e.encode(feats[1:2], out_dtype=torch.float32)[0]), m  # batch-invariant


@pytest.mark.parametrize("residual", ["f32", "split"])
def test_residual_stream_with_outlier_channels_and_offset(cuda_device, residual):
    """Pre-trained Whisper encoders carry a few massive-activation channels and non-trivial LayerNorm gains; random
    init has neither.  Emulate them: two positional channels pinned at +40 / -25, a +1.5 offset on every channel (a
    token mean larger than the token's ordinary spread), LayerNorm gains in [0.5, 2] and biases ~0.3.  Both the
    fp32 stream and the default split stream (which rounds x to bf16 BEFORE normalising) must stay inside the stated
    tolerance against the fp32 oracle."""
    import torch
    from ttasr import B200WhisperEncoder

    arch = OE.ARCHS["tiny"]
    w = OE.init_weights(arch, seed=3, std=0.05, ln_jitter=0.02)
    g = torch.Generator().manual_seed(9)
    pos = w["embed_positions.weight"].clone()
    pos += 1.5
    pos[:, 5] = 40.0
    pos[:, 77] = -25.0
    w["embed_positions.weight"] = pos
    for k in list(w):
        if k.endswith("layer_norm.weight"):
            w[k] = 0.5 + 1.5 * torch.rand(arch.d_model, generator=g)
        elif k.endswith("layer_norm.bias"):
            w[k] = 0.3 * torch.randn(arch.d_model, generator=g)
    w = OE.round_weights_bf16(w)
    enc = B200WhisperEncoder(_cfg(arch), w, residual=residual)
    feats = np.stack([OF.log_mel(OF.synth_noise(41), arch.n_mels), OF.log_mel(OF.synth_tones(42), arch.n_mels)])
    ref = OE.encoder_forward(torch.from_numpy(feats), w, arch).numpy()
    got = enc.encode(feats, out_dtype=torch.float32).cpu().numpy()
    s = _check(got, ref, "stress")
    print(f"outlier/offset stream, residual={residual}: {s}")


def test_plugin_is_constructible_from_the_factory_kwargs(cuda_device, tmp_path):
    """`ASRFactory.create_asr_pipeline(type, model_size=...)` -> `Plugin(**kwargs)` (asr_factory.py:9-30,
    streaming_asr.py:116-121): B200ASR resolves `<root>/<model_size>` like faster_whisper_asr.py:24-49, loads the
    encoder from the CTranslate2 `model.bin` found there, warms up on a 44.1 kHz stereo WAV like :269-294, splits a
    buffer longer than 30 s into windows instead of truncating it, and runs decode_fn off the event loop."""
    import asyncio
    import threading
    import types
    import wave
    import torch
    from test_model_dir import ct2_variables_from_hf, write_ct2_model_bin
    from ttasr.asr_plugin import B200ASR

    arch = OE.ARCHS["micro"]
    w = OE.round_weights_bf16(OE.init_weights(arch, seed=0, ln_jitter=0.02))
    mdir = tmp_path / "my-whisper-ct2"
    mdir.mkdir()
    write_ct2_model_bin(str(mdir / "model.bin"), ct2_variables_from_hf(w, arch, "float16"))
    (mdir / "config.json").write_text("{}")
    (mdir / "tokenizer.json").write_text("{}")
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "warm_up_excerpt.npz"))
    wav = tmp_path / "warm_up.wav"
    with wave.open(str(wav), "wb") as f:
        f.setnchannels(2); f.setsampwidth(2); f.setframerate(int(g["sampling_rate"]))
        f.writeframes(np.stack([g["pcm_ch0"]] * 2, axis=1).astype("<i2").tobytes())

    loop_thread = threading.get_ident()
    seen = []

    def decode(hidden, info):
        seen.append((tuple(hidden.shape), info["n_samples"], threading.get_ident()))
        return {"text": "好", "words": []}

    with pytest.raises(FileNotFoundError):
        B200ASR(model_size="no-such-model", model_root=str(tmp_path), decode_fn=decode)
    asr = B200ASR.from_kwargs(model_size="my-whisper-ct2", model_root=str(tmp_path), decode_fn=decode,
                              warm_up_wav=str(wav), batch_window_s=0.01)
    assert asr.weights_format == "ct2" and asr.model_path == str(mdir) and asr.model_size == "my-whisper-ct2"
    info = asr.warm_up()
    assert info["orig_sr"] == 44100 and info["channels"] == 2 and info["chunks"] == 1 and 3.9 < info["seconds"] < 4.1
    # fp16 storage of bf16-representable weights is exact: the plugin's encoder equals one built from the HF names
    rng = np.random.default_rng(12)
    long_pcm = (rng.standard_normal(16000 * 41) * 2500).astype("<i2")     # 41 s: two windows
    client = types.SimpleNamespace(scratch_buffer=bytearray(long_pcm.tobytes()), samples_width=2, last_start_time=1.5,
                                   client_id=0)
    out = asyncio.run(asr.transcribe(client))
    assert out["text"] == "好" and out["final"] and abs(out["duration"] - 41.0) < 1e-6
    shape, n, tid = seen[-1]
    assert shape == (2, 1500, arch.d_model) and n == len(long_pcm) and tid != loop_thread
    feats = np.stack([OF.log_mel(long_pcm[:480000].astype(np.float32) / 32768.0, arch.n_mels),
                      OF.log_mel(OF.pad_or_trim(long_pcm[480000:].astype(np.float32) / 32768.0), arch.n_mels)])
    ref = OE.encoder_forward(torch.from_numpy(feats), w, arch)
    hidden = asr.pipeline.encode_device(
        torch.from_numpy(np.stack([long_pcm[:480000], np.pad(long_pcm[480000:], (0, 480000 - (len(long_pcm) - 480000)))])
                         ).to(cuda_device),
        n_valid=torch.tensor([480000, len(long_pcm) - 480000], dtype=torch.int32, device=cuda_device))
    _check(hidden.float().cpu().numpy(), ref.numpy(), "micro")
    # without a decoder and without a Hugging Face directory there is nothing to decode with: say so at construction
    with pytest.raises(Exception):
        B200ASR(model_size="my-whisper-ct2", model_root=str(tmp_path))


def test_faster_whisper_padding_schemes(cuda_device):
    """Both upstream padding schemes (>= 1.1.0: 160 zero samples; <= 1.0.3: 30 s of zero audio) and padding=0, whose
    last frames reach into the whole-file reflection — all against the oracle's whole-file log-mel with the global clamp."""
    from ttasr import B200WhisperFeatureExtractor
    from ttasr.compat_faster_whisper import FileFeatureExtractor

    ffe = FileFeatureExtractor(B200WhisperFeatureExtractor(feature_size=80))
    assert ffe.chunk_length == 30 and ffe.nb_max_frames == 3000
    rng = np.random.default_rng(6)
    wave_ = (0.1 * rng.standard_normal(16000 * 33 + 91)).astype(np.float32)
    wave_[-400:] *= 8.0                                    # loud file end: the reflection matters for padding < 40
    for padding, n_pad in ((160, 160), (True, 480000), (0, 0), (17, 17)):
        got = ffe(wave_, padding=padding)
        padded = np.concatenate([wave_, np.zeros(n_pad, np.float32)])
        ref = OF.log_mel_unclamped(padded, 80)[:, :-1]
        ref = (np.maximum(ref, ref.max() - 8.0) + 4.0) / 4.0
        assert got.shape == ref.shape, (padding, got.shape, ref.shape)
        assert np.abs(got - ref).max() <= 1e-4, padding
    with pytest.raises(NotImplementedError):
        ffe(wave_, chunk_length=20)
    assert ffe(np.zeros(100, np.float32), padding=0).shape == (80, 0)
