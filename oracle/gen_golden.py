"""Generate tests/golden/*.npz from the installed Hugging Face implementation (the arithmetic the
reference calls).  Run in the build container:  python -m oracle.gen_golden

The fixtures are deliberately small: strided samples of each output plus whole-array statistics, so the
oracle (and through it the CUDA path) is pinned without committing megabytes.
Inputs are regenerated from seeds (oracle.frontend.synth_*), weights from oracle.encoder.init_weights.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from . import encoder as E
from . import frontend as FE

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
FRAME_STRIDE = 37
POS_STRIDE = 25

CLIPS = {"noise": FE.synth_noise, "tones": FE.synth_tones, "short": FE.synth_short}


def stats(a: np.ndarray) -> np.ndarray:
    a64 = a.astype(np.float64)
    return np.array([a64.sum(), a64.min(), a64.max(), np.abs(a64).sum(), (a64 * a64).sum()])


def hf_encoder(arch: E.Arch, w: dict):
    from transformers import WhisperConfig
    from transformers.models.whisper.modeling_whisper import WhisperEncoder

    cfg = WhisperConfig(
        d_model=arch.d_model, encoder_layers=arch.layers, encoder_attention_heads=arch.heads,
        encoder_ffn_dim=arch.ffn, num_mel_bins=arch.n_mels, decoder_layers=1,
        decoder_attention_heads=arch.heads, decoder_ffn_dim=64, max_source_positions=arch.n_ctx,
    )
    enc = WhisperEncoder(cfg).eval().float()
    missing, unexpected = enc.load_state_dict(w, strict=True)
    assert not missing and not unexpected
    return enc


def main() -> None:
    from transformers import WhisperFeatureExtractor

    os.makedirs(GOLDEN_DIR, exist_ok=True)
    out = {}
    for n_mels in (80, 128):
        fe = WhisperFeatureExtractor(feature_size=n_mels)
        for name, fn in CLIPS.items():
            pcm = FE.pad_or_trim(fn())
            ref = fe._np_extract_fbank_features(pcm[None], "cpu")[0]
            assert ref.shape == (n_mels, 3000) and ref.dtype == np.float32
            key = f"logmel{n_mels}_{name}"
            out[key + "_sub"] = ref[:, ::FRAME_STRIDE].copy()
            out[key + "_head"] = ref[:, :8].copy()
            out[key + "_tail"] = ref[:, -8:].copy()
            out[key + "_stats"] = stats(ref)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "frontend_hf.npz"), **out)

    out = {}
    torch.manual_seed(0)
    for arch_name in ("micro", "tiny"):
        arch = E.ARCHS[arch_name]
        w = E.round_weights_bf16(E.init_weights(arch, seed=0, ln_jitter=0.02))
        feats = torch.from_numpy(FE.log_mel(FE.synth_noise(), arch.n_mels))[None]
        with torch.no_grad():
            ref = hf_encoder(arch, w)(feats).last_hidden_state[0].numpy()
        assert ref.shape == (1500, arch.d_model)
        out[f"enc_{arch_name}_sub"] = ref[::POS_STRIDE].copy()
        out[f"enc_{arch_name}_stats"] = stats(ref)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "encoder_hf.npz"), **out)
    for f in ("frontend_hf.npz", "encoder_hf.npz"):
        print(f, os.path.getsize(os.path.join(GOLDEN_DIR, f)), "bytes")


def gen_ingest():
    """tests/golden/ingest_scipy.npz: scipy.signal.resample_poly (the function librosa's res_type="polyphase" calls)
    run directly on deterministic frames, independent of oracle/ingest.py's own wrapper."""
    import math

    import scipy.signal

    from oracle import ingest as OI

    out = {}
    for sr, ch, dt in ((48000, 2, np.int16), (44100, 2, np.int16), (44100, 1, np.float32), (8000, 1, np.int16)):
        frames = OI.synth_frames(sr, 0.25, ch, dt, seed=5)
        x = frames.astype(np.float32) / np.float32(32768.0) if dt == np.int16 else frames.astype(np.float32)
        if ch > 1:
            x = x.mean(axis=1, dtype=np.float32)
        g = math.gcd(sr, 16000)
        y = scipy.signal.resample_poly(x.astype(np.float32), 16000 // g, sr // g)
        n = int(math.ceil(len(x) * 16000 / sr))
        out[f"{sr}_{ch}_{np.dtype(dt).name}"] = np.asarray(y[:n], dtype=np.float32)
    path = os.path.join(GOLDEN_DIR, "ingest_scipy.npz")
    np.savez_compressed(path, **out)
    print("ingest_scipy.npz", os.path.getsize(path), "bytes")


def gen_warmup(ref_wav="/root/reference/api/stt_streaming/warm_up.wav"):
    """tests/golden/warm_up_excerpt.npz: 4 s (0.3 s .. 4.3 s: a spoken phrase, a 0.5 s pause, the start of the next
    phrase) of the reference's only real audio fixture — api/stt_streaming/warm_up.wav, 44.1 kHz 16-bit stereo, the
    file `FasterWhisperASR.warm_up` transcribes (faster_whisper_asr.py:269-294).  Its two channels are sample-identical,
    so one is stored and the test rebuilds the interleaved stereo frames.  Plus the expected 16 kHz mono signal from
    scipy.signal.resample_poly and the HF numpy log-mel (80 / 128 bins) of that signal, strided."""
    import math
    import wave

    import scipy.signal
    from transformers import WhisperFeatureExtractor

    with wave.open(ref_wav, "rb") as w:
        sr, ch = w.getframerate(), w.getnchannels()
        frames = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").reshape(-1, ch)
    assert (sr, ch) == (44100, 2) and np.array_equal(frames[:, 0], frames[:, 1])
    a, b = int(0.3 * sr), int(4.3 * sr)
    exc = frames[a:b]
    x = exc.astype(np.float32) / np.float32(32768.0)
    x = x.mean(axis=1, dtype=np.float32)
    g = math.gcd(sr, 16000)
    y = scipy.signal.resample_poly(x, 16000 // g, sr // g)[: int(math.ceil(len(x) * 16000 / sr))].astype(np.float32)
    out = {"pcm_ch0": exc[:, 0].copy(), "sampling_rate": np.int32(sr), "channels": np.int32(ch),
           "mono16k_sub": y[::16].copy(), "mono16k_stats": stats(y)}
    for n_mels in (80, 128):
        ref = WhisperFeatureExtractor(feature_size=n_mels)._np_extract_fbank_features(FE.pad_or_trim(y)[None], "cpu")[0]
        out[f"logmel{n_mels}_sub"] = ref[:, :420:3].copy()       # the 4 s of audio are frames 0 .. 399
        out[f"logmel{n_mels}_stats"] = stats(ref)
    path = os.path.join(GOLDEN_DIR, "warm_up_excerpt.npz")
    np.savez_compressed(path, **out)
    print("warm_up_excerpt.npz", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    if "--ingest-only" in sys.argv:
        gen_ingest()
    elif "--warmup-only" in sys.argv:
        gen_warmup()
    else:
        main()
        gen_ingest()
        gen_warmup()
