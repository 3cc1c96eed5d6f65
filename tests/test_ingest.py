"""Ingest step.  CPU: the numpy filter design against scipy's, the oracle against the golden fixture.
GPU: ttasr_ingest_run against the oracle (scipy.signal.resample_poly = librosa's res_type="polyphase")."""
import math
import os

import numpy as np
import pytest

from oracle import ingest as OI

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ingest_scipy.npz")
TOL = 2e-6  # fp32 accumulation order over <= 61 taps on |x| <= 0.4


@pytest.mark.parametrize("orig_sr", [48000, 44100, 8000, 22050, 32000])
def test_filter_design_matches_scipy_firwin(orig_sr):
    import scipy.signal
    from ttasr.ingest import resample_poly_filter

    g = math.gcd(orig_sr, 16000)
    up, down = 16000 // g, orig_sr // g
    mr = max(up, down)
    want = (scipy.signal.firwin(2 * 10 * mr + 1, 1.0 / mr, window=("kaiser", 5.0)).astype(np.float32) * np.float32(up))
    got = resample_poly_filter(up, down)
    assert got.dtype == np.float32 and got.shape == want.shape
    assert np.abs(got - want).max() <= 2e-7 * np.abs(want).max()


def test_oracle_matches_golden_fixture():
    g = np.load(GOLDEN)
    for key in ("48000_2_int16", "44100_2_int16", "44100_1_float32", "8000_1_int16"):
        sr, ch, dt = key.split("_")
        frames = OI.synth_frames(int(sr), 0.25, int(ch), np.int16 if dt == "int16" else np.float32, seed=5)
        y = OI.load_like_librosa(frames, int(sr))
        assert y.shape == g[key].shape
        assert np.abs(y - g[key]).max() <= 1e-7


def test_oracle_chunking():
    y = np.arange(1, 480000 * 2 + 11, dtype=np.float32)
    rows, nv = OI.chunk(y)
    assert rows.shape == (3, 480000) and list(nv) == [480000, 480000, 10]
    assert rows[2, 9] == y[-1] and rows[2, 10] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("orig_sr,channels,dtype,seconds", [
    (48000, 2, np.int16, 1.37), (44100, 2, np.int16, 0.81), (44100, 1, np.float32, 0.5), (8000, 1, np.int16, 2.0),
    (16000, 2, np.int16, 0.7), (22050, 3, np.float32, 0.33), (48000, 1, np.float32, 0.001)])
def test_ingest_matches_oracle(cuda_device, orig_sr, channels, dtype, seconds):
    import torch
    from ttasr import B200AudioIngest

    frames = OI.synth_frames(orig_sr, seconds, channels, dtype, seed=3)
    want = OI.load_like_librosa(frames, orig_sr)
    ing = B200AudioIngest(orig_sr)
    flat = ing.load(torch.from_numpy(frames).to(cuda_device), pad_to_chunks=False).cpu().numpy()
    assert flat.shape == want.shape == (ing.out_len(frames.shape[0]),)
    assert np.abs(flat - want).max() <= TOL
    rows, nv = ing.load(torch.from_numpy(frames).to(cuda_device))
    ref_rows, ref_nv = OI.chunk(want)
    assert rows.shape == ref_rows.shape and np.array_equal(nv.cpu().numpy(), ref_nv)
    assert np.abs(rows.cpu().numpy() - ref_rows).max() <= TOL
    assert np.all(rows.cpu().numpy().reshape(-1)[len(want):] == 0)


@pytest.mark.gpu
def test_ingest_long_file_chunks_feed_the_front_end(cuda_device):
    """70 s of 44.1 kHz stereo int16 -> 3 chunks whose log-mel equals the oracle's on the oracle's resampled audio."""
    import torch
    from oracle import frontend as OF
    from ttasr import B200AudioIngest, B200WhisperFeatureExtractor

    frames = OI.synth_frames(44100, 70.0, 2, np.int16, seed=9)
    want_rows, want_nv = OI.chunk(OI.load_like_librosa(frames, 44100))
    rows, nv = B200AudioIngest(44100).load(torch.from_numpy(frames).to(cuda_device))
    assert rows.shape == (3, 480000) and np.array_equal(nv.cpu().numpy(), want_nv)
    feats = B200WhisperFeatureExtractor(feature_size=80).extract(rows, n_valid=nv).cpu().numpy()
    for i in range(3):
        assert np.abs(feats[i] - OF.log_mel(want_rows[i], 80)).max() <= 1e-4


@pytest.mark.gpu
def test_ingest_bad_arguments(cuda_device):
    import torch
    from ttasr import B200AudioIngest, TtasrError

    ing = B200AudioIngest(48000)
    with pytest.raises(TtasrError):
        ing.load(torch.zeros((10, 2), dtype=torch.float64, device=cuda_device))
    with pytest.raises(TtasrError):
        ing.load(torch.zeros((10, 9), dtype=torch.int16, device=cuda_device))
    with pytest.raises(ValueError):
        B200AudioIngest(44100.5)


# ------------------------------------------------------------------ real audio + pause-aligned chunking (round 2)
def _warm_up_excerpt():
    import os

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "warm_up_excerpt.npz"))
    ch0 = g["pcm_ch0"]
    frames = np.stack([ch0, ch0], axis=1)           # the reference file's two channels are sample-identical
    return g, frames, int(g["sampling_rate"])


@pytest.mark.gpu
def test_real_audio_warm_up_wav_through_ingest_and_front_end(cuda_device):
    """The reference's only real recording (api/stt_streaming/warm_up.wav, 44.1 kHz stereo; faster_whisper_asr.py:
    279-294) -> B200AudioIngest (mono mix, 160/441 polyphase resampling, chunking) -> log-mel, against scipy
    resample_poly and the HF numpy extractor: live when they import, and against the committed golden samples."""
    import torch
    from ttasr import B200AudioIngest, B200WhisperFeatureExtractor

    g, frames, sr = _warm_up_excerpt()
    chunks, n_valid = B200AudioIngest(sr).load(torch.from_numpy(frames).to(cuda_device))
    n = int(n_valid[0])
    assert chunks.shape == (1, 480000) and n == int(np.ceil(frames.shape[0] * 16000 / sr))
    y = chunks[0, :n].cpu().numpy()
    assert np.abs(y[::16] - g["mono16k_sub"]).max() <= 2e-6
    assert abs(float(y.astype(np.float64).sum()) - float(g["mono16k_stats"][0])) <= 1e-3
    assert float(chunks[0, n:].abs().max()) == 0.0
    for n_mels in (80, 128):
        feats = B200WhisperFeatureExtractor(feature_size=n_mels).extract(chunks, n_valid=n_valid)[0].cpu().numpy()
        assert np.abs(feats[:, :420:3] - g[f"logmel{n_mels}_sub"]).max() <= 1e-4
        assert abs(float(feats.astype(np.float64).sum()) - float(g[f"logmel{n_mels}_stats"][0])) <= 1e-4 * feats.size
    # live against the oracle chain (numpy log-mel of the scipy-resampled signal)
    from oracle import frontend as OF
    from oracle import ingest as OI

    ref_y = OI.load_like_librosa(frames, sr)
    assert np.abs(y - ref_y).max() <= 2e-6
    ref = OF.log_mel(OF.pad_or_trim(ref_y), 128)
    feats = B200WhisperFeatureExtractor(feature_size=128).extract(chunks, n_valid=n_valid)[0].cpu().numpy()
    assert np.abs(feats - ref).max() <= 1e-4


@pytest.mark.gpu
def test_pause_aligned_chunks_on_real_speech(cuda_device):
    """Long-form ingest (BASELINE.json configs[3]): the excerpt tiled to ~100 s; every cut of `load_aligned` must fall
    into one of the recording's pauses (or be a 30 s hard cut when a window holds none), chunks tile the signal, and
    each row is the corresponding slice of the flat 16 kHz signal, zero padded."""
    import torch
    from ttasr import B200AudioIngest

    _, frames, sr = _warm_up_excerpt()
    long = np.concatenate([frames] * 25)              # 100 s, a 0.5 s pause every 4 s
    ing = B200AudioIngest(sr)
    dev_frames = torch.from_numpy(long).to(cuda_device)
    flat = ing.load(dev_frames, pad_to_chunks=False)
    chunks, n_valid, starts = ing.load_aligned(dev_frames)
    n = int(flat.shape[0])
    lens = n_valid.cpu().numpy().astype(np.int64)
    st = starts.numpy()
    assert st[0] == 0 and (st[1:] == (st[:-1] + lens[:-1])).all() and st[-1] + lens[-1] == n
    assert lens.max() <= 480000 and len(lens) == 4    # 100 s in <= 30 s pieces cut at pauses
    flat_np = flat.cpu().numpy()
    rows = chunks.cpu().numpy()
    for r in range(len(lens)):
        assert np.array_equal(rows[r, : lens[r]], flat_np[st[r]: st[r] + lens[r]])
        assert not rows[r, lens[r]:].any()
    # interior cuts sit in quiet audio: the 100 ms around each one is at least 25 dB under the recording's loud parts
    loud = 10 * np.log10(np.percentile(flat_np ** 2, 99) + 1e-12)
    for c in st[1:]:
        seg = flat_np[c - 800: c + 800]
        assert 10 * np.log10((seg ** 2).mean() + 1e-12) <= loud - 25.0, c
    # caller-supplied cut list (e.g. from the host's own VAD) and its validation
    chunks2, nv2, st2 = ing.load_aligned(dev_frames, cut_points=[(0, 160000), (160000, 400000)])
    assert chunks2.shape == (2, 480000) and nv2.tolist() == [160000, 240000]
    assert np.array_equal(chunks2[1, :240000].cpu().numpy(), flat_np[160000:400000])
    with pytest.raises(ValueError):
        ing.load_aligned(dev_frames, cut_points=[(0, 480001)])
