"""Parity of the fused CUDA log-mel front end (through the C ABI) against the numpy oracle: |diff| <= 1e-4 on the
final (x + 4) / 4 features, plus the analytic cases of SURVEY.md section 8c."""
import numpy as np
import pytest

from oracle import frontend as OF

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _fe(n_mels):
    from ttasr import B200WhisperFeatureExtractor

    return B200WhisperFeatureExtractor(feature_size=n_mels)


@pytest.mark.parametrize("n_mels", [80, 128])
def test_logmel_matches_oracle_on_config1_clips(cuda_device, n_mels):
    import torch

    clips = [OF.pad_or_trim(f()) for f in (OF.synth_noise, OF.synth_tones, OF.synth_short)]
    clips.append(np.zeros(OF.N_SAMPLES, np.float32))
    pcm = torch.from_numpy(np.stack(clips)).to(cuda_device)
    got = _fe(n_mels).extract(pcm).cpu().numpy()
    assert got.shape == (4, n_mels, 3000) and got.dtype == np.float32
    for i, c in enumerate(clips):
        ref = OF.log_mel(c, n_mels)
        err = np.abs(got[i] - ref).max()
        assert err <= TOL, f"clip {i}: max abs err {err}"
    # K1: all-zero PCM -> exactly -1.5 everywhere
    assert np.all(got[3] == -1.5)


def test_golden_fixture(cuda_device):
    import os
    import torch

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "frontend_hf.npz"))
    for n_mels in (80, 128):
        for name, fn in (("noise", OF.synth_noise), ("tones", OF.synth_tones), ("short", OF.synth_short)):
            pcm = torch.from_numpy(OF.pad_or_trim(fn())[None]).to(cuda_device)
            got = _fe(n_mels).extract(pcm)[0].cpu().numpy()
            key = f"logmel{n_mels}_{name}"
            assert np.abs(got[:, ::37] - g[key + "_sub"]).max() <= TOL
            assert np.abs(got[:, :8] - g[key + "_head"]).max() <= TOL
            assert np.abs(got[:, -8:] - g[key + "_tail"]).max() <= TOL
            assert abs(got.astype(np.float64).sum() - g[key + "_stats"][0]) <= TOL * got.size


def test_int16_input_and_ragged_rows(cuda_device):
    import torch

    rng = np.random.default_rng(5)
    lens = [480000, 48000, 16000, 123457, 0]
    rows16 = [(rng.standard_normal(n) * 3000).astype(np.int16) for n in lens]
    maxlen = max(lens)
    host = np.zeros((len(lens), maxlen), np.int16)
    for i, r in enumerate(rows16):
        host[i, : len(r)] = r
    fe = _fe(80)
    pcm = torch.from_numpy(host).to(cuda_device)
    nv = torch.tensor(lens, dtype=torch.int32, device=cuda_device)
    got = fe.extract(pcm, n_valid=nv).cpu().numpy()
    for i, r in enumerate(rows16):
        ref = OF.log_mel(r.astype(np.float32) / 32768.0, 80)
        assert np.abs(got[i] - ref).max() <= TOL, f"row {i} (len {lens[i]})"
    # K2: frames wholly inside the zero padding sit at the clamp floor == the row minimum
    assert got[2, :, 2000:].max() == got[2].min()
    # poison the padding: with n_valid the kernel must not read it
    host2 = host.copy()
    for i, n in enumerate(lens):
        host2[i, n:] = 12345
    got2 = fe.extract(torch.from_numpy(host2).to(cuda_device), n_valid=nv).cpu().numpy()
    assert np.array_equal(got, got2)


def test_batch_invariance_and_time_major_copy(cuda_device):
    import torch

    g = torch.Generator(device=cuda_device).manual_seed(1234)
    pcm = (0.1 * torch.randn((9, OF.N_SAMPLES), generator=g, device=cuda_device)).clamp_(-1, 1)
    fe = _fe(128)
    feats, tm = fe.extract(pcm, return_time_major=True)
    single = fe.extract(pcm[4:5])
    assert torch.equal(feats[4], single[0])  # K4: bit-identical regardless of batch position
    assert tm.shape == (9, 3000, 128) and tm.dtype == torch.bfloat16
    assert torch.equal(tm.transpose(1, 2).float(), feats.to(torch.bfloat16).float())
    ref = OF.log_mel(pcm[7].cpu().numpy(), 128)
    assert np.abs(feats[7].cpu().numpy() - ref).max() <= TOL


def test_reference_call_surface(cuda_device):
    """The call train_asr.py:610-616 makes, on the CUDA extractor."""
    fe = _fe(80)
    x = OF.synth_short()
    out = fe(x, sampling_rate=16000, return_attention_mask=True)
    f = out.get("input_features")[0]
    assert f.shape == (80, 3000)
    assert np.abs(f - OF.log_mel(x, 80)).max() <= TOL
    assert out.get("attention_mask")[0].shape == (3000,) and out["attention_mask"][0].sum() == 300
    with pytest.raises(ValueError):
        fe(x, sampling_rate=8000)


@pytest.mark.parametrize("n_mels", [80, 128])
def test_clamp_pass_paths_f32_and_time_major(cuda_device, n_mels):
    """The second kernel only touches tiles below max - 8 (in-data silence, large dynamic range) and the padding tiles
    the first kernel skipped (n_valid): both the fp32 features and the bf16 time-major copy must carry the clamp."""
    import torch

    rng = np.random.default_rng(11)
    loud = (0.3 * rng.standard_normal(OF.N_SAMPLES)).astype(np.float32)
    a = loud.copy()
    a[100_000:300_000] = 0.0                       # digital silence inside the clip (stays a "written" tile)
    b = loud.copy()
    b[200_000:] *= 1e-5                            # 100 dB quieter tail: values below the clamp, not constant
    c = OF.pad_or_trim(loud[:50_000])              # short clip, padding given as n_valid
    clips = [a, b, c]
    pcm = torch.from_numpy(np.stack(clips)).to(cuda_device)
    nv = torch.tensor([OF.N_SAMPLES, OF.N_SAMPLES, 50_000], dtype=torch.int32, device=cuda_device)
    fe = _fe(n_mels)
    feats, tm = fe.extract(pcm, n_valid=nv, return_time_major=True)
    feats, tm = feats.cpu().numpy(), tm.float().cpu().numpy()
    for i, clip in enumerate(clips):
        ref = OF.log_mel(clip, n_mels)
        assert np.abs(feats[i] - ref).max() <= TOL, f"clip {i}"
        assert ref.min() == pytest.approx(ref.max() - 2.0, abs=1e-5)   # the clamp is active in every clip
        # bf16 copy = rounding of the final fp32 features, channels zero-padded
        want = torch.from_numpy(feats[i].T.copy()).to(torch.bfloat16).float().numpy()
        assert np.array_equal(tm[i, :, :n_mels], want), f"clip {i}: time-major copy differs from the fp32 features"
        assert np.all(tm[i, :, n_mels:] == 0)


def test_padding_tiles_after_real_tiles_do_not_race(cuda_device):
    """Regression: a CTA that walks from a real tile straight into an all-padding tile (no block barrier on that path)
    used to overwrite the per-warp tile minima the previous tile was still reducing, so the clamp kernel took the real
    tile for an unwritten one.  Ragged batch larger than one wave of CTAs, repeated with the allocator churned."""
    import torch

    rng = np.random.default_rng(5)
    lens = [480000, 48000, 16000, 123457, 0, 480000, 300000]
    rows = [(0.1 * rng.standard_normal(n)).astype(np.float32) for n in lens]
    host = np.zeros((len(lens), max(lens)), np.float32)
    for i, r in enumerate(rows):
        host[i, : len(r)] = r
    fe = _fe(80)
    pcm = torch.from_numpy(host).to(cuda_device)
    nv = torch.tensor(lens, dtype=torch.int32, device=cuda_device)
    refs = [OF.log_mel(r, 80) for r in rows]
    for it in range(6):
        torch.empty((9, 80, 3000), device=cuda_device).fill_(float(it))
        got = fe.extract(pcm, n_valid=nv).cpu().numpy()
        for i, ref in enumerate(refs):
            assert np.abs(got[i] - ref).max() <= TOL, f"iteration {it}, row {i} (len {lens[i]})"


def test_full_size_properties_b256_128mel(cuda_device):
    """BASELINE.json configs[2] front end at full size (256 x 30 s, 128 bins): per-row properties the domain offers —
    every row spans exactly [max - 2, max] after the clamp + affine map, a row's features do not depend on the batch
    around it, the bf16 time-major copy is the rounding of the fp32 features, and 8 sampled rows match the oracle."""
    import torch

    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(1234)
    pcm = (0.1 * torch.randn((256, 480000), device=dev, generator=g)).clamp_(-1, 1)
    pcm[3, 100000:] *= 1e-6                       # one row with a clamped tail
    fe = _fe(128)
    feats, tm = fe.extract(pcm, return_time_major=True)
    assert feats.shape == (256, 128, 3000) and tm.shape[:2] == (256, 3000)
    mx = feats.amax(dim=(1, 2))
    mn = feats.amin(dim=(1, 2))
    assert bool((mn >= mx - 2.0 - 1e-6).all())
    assert float(mn[3]) == pytest.approx(float(mx[3]) - 2.0, abs=1e-6)
    assert torch.equal(tm[..., :128], feats.transpose(1, 2).to(torch.bfloat16))
    for r in (0, 3, 255):
        assert torch.equal(fe.extract(pcm[r: r + 1].contiguous())[0], feats[r])
    host = pcm.cpu().numpy()
    for r in (0, 3, 37, 64, 128, 200, 254, 255):
        assert np.abs(feats[r].cpu().numpy() - OF.log_mel(host[r], 128)).max() <= TOL, f"row {r}"


@pytest.mark.parametrize("n_mels", [80, 128])
def test_matches_the_torch_extractor_the_reference_dispatches_to(cuda_device, n_mels):
    """with torch installed `WhisperFeatureExtractor.__call__` runs `_torch_extract_fbank_features`
    (feature_extraction_whisper.py:135-164,317-320), fp32 torch.stft — the variant the reference's `fe(...)` call really
    executes.  It differs from the numpy oracle (a3, what this build is pinned to) by <= ~1e-5 (K8), so the CUDA path
    must agree with it within 1.5e-4, on the CPU and on the GPU variant of a4 alike."""
    import torch

    transformers = pytest.importorskip("transformers")
    hf = transformers.WhisperFeatureExtractor(feature_size=n_mels)
    clips = np.stack([OF.pad_or_trim(f()) for f in (OF.synth_noise, OF.synth_tones, OF.synth_short)])
    got = _fe(n_mels).extract(torch.from_numpy(clips).to(cuda_device)).cpu().numpy()
    for device in ("cpu", "cuda"):
        ref = np.asarray(hf._torch_extract_fbank_features(clips, device))
        assert ref.shape == got.shape
        err = float(np.abs(got - ref).max())
        print(f"a4 ({device}), {n_mels} mel bins: max abs diff {err:.2e}")
        assert err <= 1.5e-4
    # and through the public call surface, the way train_asr.py:607-616 uses it (batched per-row max, a4 semantics)
    out = hf(list(clips), sampling_rate=16000, return_tensors="np")["input_features"]
    assert float(np.abs(got - out).max()) <= 1.5e-4


def test_compiled_in_mel_projection_is_selected_and_bit_equal_to_the_generic_program(cuda_device, monkeypatch):
    """The 80-filter Whisper bank runs straight-line mel code generated at build time (csrc/mel_baked.inc); the generic
    per-bin program (any bank; forced with TTASR_FRONTEND_MEL=generic) accumulates in the same order, so both must give
    the same bits: fp32 features and the bf16 time-major copy, f32 and int16 PCM, full tiles, the partial last tile and
    n_valid padding."""
    import torch

    clips = np.stack([OF.pad_or_trim(f()) for f in (OF.synth_noise, OF.synth_tones, OF.synth_short)])
    pcm = torch.from_numpy(clips).to(cuda_device)
    pcm_i16 = (pcm * 20000).to(torch.int16)
    nv = torch.tensor([480000, 300001, 48000], dtype=torch.int32, device=cuda_device)
    n_mels = 80
    assert _fe(128).mel_mode() == 0   # the 128-filter bank stays on the program (instruction-cache footprint)
    baked = _fe(n_mels)
    assert baked.mel_mode() == n_mels, "the standard bank did not select the compiled-in projection"
    monkeypatch.setenv("TTASR_FRONTEND_MEL", "generic")
    generic = _fe(n_mels)
    assert generic.mel_mode() == 0
    for x, kw in ((pcm, {}), (pcm_i16, {}), (pcm, {"n_valid": nv})):
        fa, ta = baked.extract(x, return_time_major=True, **kw)
        fb, tb = generic.extract(x, return_time_major=True, **kw)
        assert torch.equal(fa, fb) and torch.equal(ta, tb)
        _, tc = baked.extract(x, return_time_major=True, features=False, **kw)
        assert torch.equal(tc, ta)


def test_generic_mel_program_on_a_non_whisper_bank(cuda_device):
    """Any other bank of overlapping triangles (here 64 filters) runs the host-built program."""
    import torch

    clips = [OF.pad_or_trim(f()) for f in (OF.synth_noise, OF.synth_tones, OF.synth_short)]
    pcm = torch.from_numpy(np.stack(clips)).to(cuda_device)
    fe = _fe(64)
    assert fe.mel_mode() == 0
    got = fe.extract(pcm).cpu().numpy()
    for i, c in enumerate(clips):
        err = np.abs(got[i] - OF.log_mel(c, 64)).max()
        assert err <= TOL, f"clip {i}: max abs err {err}"
