"""Host-side logic that needs no GPU: constants, the reference-facing argument handling and error behaviour of the
drop-in classes, persistence, the data-parallel sharder (incl. a world_size-2 gloo run), the host FFT check."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import frontend as OF

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_mel_filters_and_window_equal_the_oracle():
    from ttasr import mel

    for n in (80, 128):
        assert np.array_equal(mel.slaney_mel_filters(n), OF.mel_filter_bank(n))
        fb = mel.slaney_mel_filters(n)
        assert (fb != 0).sum(axis=1).max() <= 2 and not fb[0].any() and not fb[200].any()
    assert np.array_equal(mel.periodic_hann(400), OF.hann_window(400))


def test_feature_extractor_attributes_and_errors():
    from ttasr import B200WhisperFeatureExtractor

    fe = B200WhisperFeatureExtractor(feature_size=128)
    assert fe.sampling_rate == 16000 and fe.n_samples == 480000 and fe.nb_max_frames == 3000
    assert fe.model_input_names[0] == "input_features" and fe.mel_filters.shape == (201, 128)
    x = np.zeros(1600, np.float32)
    with pytest.raises(ValueError, match="sampling rate"):
        fe(x, sampling_rate=22050)
    with pytest.raises(ValueError, match="mono"):
        fe(np.zeros((2, 2, 100), np.float32), sampling_rate=16000)
    with pytest.raises(NotImplementedError):
        fe(x, sampling_rate=16000, padding="longest")


def test_feature_extractor_pad_and_persistence(tmp_path):
    from ttasr import B200WhisperFeatureExtractor

    fe = B200WhisperFeatureExtractor(feature_size=80)
    feats = [{"input_features": np.full((80, 3000), i, np.float32)} for i in range(3)]
    batch = fe.pad(feats, return_tensors="pt")  # the collator call of train_asr.py:296-298
    assert tuple(batch["input_features"].shape) == (3, 80, 3000) and float(batch["input_features"][2, 0, 0]) == 2.0
    with pytest.raises(ValueError):
        fe.pad([{"input_features": np.zeros((80, 3000))}, {"input_features": np.zeros((80, 2999))}])
    fe.save_pretrained(str(tmp_path))
    cfg = json.load(open(tmp_path / "preprocessor_config.json"))
    assert cfg["feature_size"] == 80 and cfg["feature_extractor_type"] == "WhisperFeatureExtractor"
    again = B200WhisperFeatureExtractor.from_pretrained(str(tmp_path))
    assert again.to_dict() == fe.to_dict()
    tf = pytest.importorskip("transformers")
    hf = tf.WhisperFeatureExtractor.from_pretrained(str(tmp_path))  # the saved config round-trips through HF
    assert hf.feature_size == 80 and hf.n_fft == 400 and hf.hop_length == 160


def test_encoder_config_helpers():
    from ttasr import EncoderConfig

    c = EncoderConfig.named("large-v3")
    assert (c.d_model, c.encoder_layers, c.num_mel_bins) == (1280, 32, 128)
    assert c.flops_per_chunk() == 2_273_771_520_000
    assert EncoderConfig.from_any({"d_model": 384, "encoder_layers": 4, "encoder_attention_heads": 6,
                                   "encoder_ffn_dim": 1536, "num_mel_bins": 80}).max_source_positions == 1500


def test_shard_bounds_cover_everything_once():
    from ttasr.dp import shard_bounds

    for n, world, keep in [(4096, 8, 1), (4096, 8, 10), (7, 4, 1), (0, 2, 1), (410, 3, 10), (5, 8, 1)]:
        spans = [shard_bounds(n, r, world, keep) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all(lo % keep == 0 for lo, _ in spans if lo < n)
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) < 2 * keep  # one group of imbalance + a partial last group
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


GLOO_SCRIPT = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from ttasr.dp import shard_bounds, gather_host, all_max
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
r = dist.get_rank()
lo, hi = shard_bounds(4096, r, 2, keep_together=10)
got = gather_host({"rank": r, "span": (lo, hi), "checksum": sum(range(lo, hi))})
m = all_max(1.0 + r)
if r == 0:
    assert [g["rank"] for g in got] == [0, 1]
    assert got[0]["span"][1] == got[1]["span"][0] and got[1]["span"][1] == 4096
    assert sum(g["checksum"] for g in got) == sum(range(4096))
else:
    assert got is None
assert m == 2.0
dist.destroy_process_group()
print("ok", r)
"""


def test_two_rank_gloo_shard_and_gather(tmp_path):
    script = tmp_path / "gloo_dp.py"
    script.write_text(GLOO_SCRIPT)
    port = str(29000 + os.getpid() % 2000)
    pkg = os.path.join(ROOT, "taiwan-tongues-asr-ce_b200")
    procs = [subprocess.Popen([sys.executable, str(script), pkg, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "ok 0" in outs[0] and "ok 1" in outs[1]


def test_fft400_factorisation_on_the_host(tmp_path):
    """csrc/fft400.cuh compiled for the host: 20 x 20 PFA passes + Hermitian split vs a naive fp64 DFT."""
    exe = tmp_path / "fft400_check"
    subprocess.run(["g++", "-O2", "-I", os.path.join(ROOT, "taiwan-tongues-asr-ce_b200", "csrc"),
                    os.path.join(ROOT, "tests", "host", "fft400_check.cpp"), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert float(r.stdout.split()[1]) < 2e-3


def test_file_level_features_match_whole_file_stft(monkeypatch):
    """faster-whisper semantics (whole-file log-mel, global clamp) assembled from 30 s chunk calls: the chunking /
    halo / re-clamp logic of ttasr.compat_faster_whisper, with the CUDA extractor replaced by the oracle."""
    import torch
    from ttasr import B200WhisperFeatureExtractor
    from ttasr.compat_faster_whisper import FileFeatureExtractor

    fe = B200WhisperFeatureExtractor(feature_size=80)
    monkeypatch.setattr(fe, "_torch_device", lambda: torch.device("cpu"))
    monkeypatch.setattr(fe, "extract", lambda pcm, **kw: torch.from_numpy(OF.log_mel_batch(pcm.numpy(), 80)))
    rng = np.random.default_rng(3)
    wave = (0.05 * rng.standard_normal(16000 * 75 + 1234)).astype(np.float32)  # 75 s -> 3 chunks
    wave[16000 * 40:] *= 20.0                                                   # the file max sits in a later chunk
    got = FileFeatureExtractor(fe)(wave, padding=160)
    padded = np.concatenate([wave, np.zeros(160, np.float32)])
    ref = OF.log_mel_unclamped(padded, 80)[:, :-1]
    ref = (np.maximum(ref, ref.max() - 8.0) + 4.0) / 4.0
    assert got.shape == ref.shape == (80, padded.shape[0] // 160)
    assert np.abs(got - ref).max() < 2e-6


def test_asr_plugin_microbatcher_batches_concurrent_clients():
    """B200ASR (ASRInterface): three clients ready within one window -> ONE encode launch; reference result fields."""
    import asyncio
    import types
    import torch
    from ttasr.asr_plugin import B200ASR, pcm_bytes_to_tensor

    calls = []

    class FakePipe:  # stands in for the CUDA pipeline: records what would be launched
        device = torch.device("cpu")
        feature_extractor = types.SimpleNamespace(n_samples=480000, sampling_rate=16000)

        def encode_device(self, pcm, n_valid=None):
            calls.append((tuple(pcm.shape), n_valid.tolist()))
            return torch.zeros((pcm.shape[0], 1500, 8))

    monkey_pin = torch.Tensor.pin_memory
    torch.Tensor.pin_memory = lambda self, *a, **k: self  # no CUDA on the CPU box
    try:
        asr = B200ASR(FakePipe(), lambda hidden, info: {"text": f"n={info['n_samples']}", "words": []},
                      batch_window_s=0.01)
        clients = [types.SimpleNamespace(scratch_buffer=bytearray(np.full(n, 7, "<i2").tobytes()), samples_width=2,
                                         last_start_time=1.5, client_id=i) for i, n in enumerate((16000, 40000, 24001))]

        async def run():
            return await asyncio.gather(*(asr.transcribe(c) for c in clients))

        results = asyncio.run(run())
    finally:
        torch.Tensor.pin_memory = monkey_pin
    assert len(calls) == 1 and calls[0][0][0] == 3 and calls[0][1] == [16000, 40000, 24001]
    assert [r["text"] for r in results] == ["n=16000", "n=40000", "n=24001"]
    assert all(r["final"] and r["language"] == "zh" and set(r) == {"language", "language_probability", "final",
                                                                    "text", "duration", "words"} for r in results)
    assert pcm_bytes_to_tensor(bytearray(b"\x01\x00\xff\xff\x05")).tolist() == [1, -1]


def test_generated_mel_code_is_current():
    """csrc/mel_baked.inc (committed) is what csrc/gen_mel_baked.py generates from ttasr/mel.py today, and its hashes are
    those of the tables the Python extractor hands to ttasr_frontend_create."""
    import importlib.util
    import os

    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "taiwan-tongues-asr-ce_b200", "csrc")
    spec = importlib.util.spec_from_file_location("gen_mel_baked", os.path.join(csrc, "gen_mel_baked.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    with open(os.path.join(csrc, "mel_baked.inc")) as f:
        assert f.read() == gen.generate()
    from ttasr import B200WhisperFeatureExtractor

    for n_mels in gen.BANKS:
        table = np.ascontiguousarray(B200WhisperFeatureExtractor(feature_size=n_mels).mel_filters, dtype=np.float32)
        assert f"0x{gen.fnv1a64(table.tobytes()):016x}ull" in gen.gen_bank(n_mels)
