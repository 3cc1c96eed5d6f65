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
