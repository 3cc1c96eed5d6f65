#!/usr/bin/env python
"""Numerics of running the front end's two 20-point DFT stages on the tensor cores (tcgen05 kind::tf32), emulated on the
CPU: operands rounded to TF32 (10 explicit mantissa bits, round-to-nearest-even), products accumulated in fp32 — exactly
what the MMA does — for (a) one TF32 pass and (b) the 3xTF32 split (a = a_hi + a_lo, three MMAs, the lo*lo term dropped).
Everything else (window, twiddles between the stages, power, mel, log, clamp) as in csrc/frontend_logmel.cu, in fp32.
Reports max |log-mel - oracle| on the config-1 clips against the 1e-4 gate.

    python tools/tf32_dft_emulation.py > profiles/r2_tf32_dft_emulation.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import frontend as OF  # noqa: E402


def tf32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    u = (u + 0xFFF + ((u >> 13) & 1)) & ~np.uint64(0x1FFF)       # round to nearest even at bit 13
    return u.astype(np.uint32).view(np.float32)


def mma(a, b, mode):
    """a [.., m, k] x b [k, n], fp32 accumulate (emulated in fp64, then rounded once: the accumulation error of a
    20- or 40-term fp32 dot product is far below TF32's operand rounding)."""
    if mode == "fp32":
        return (a.astype(np.float64) @ b.astype(np.float64)).astype(np.float32)
    ah, bh = tf32(a), tf32(b)
    if mode == "tf32":
        return (ah.astype(np.float64) @ bh.astype(np.float64)).astype(np.float32)
    al, bl = tf32(a - ah), tf32(b - bh)
    acc = ah.astype(np.float64) @ bh.astype(np.float64) + ah.astype(np.float64) @ bl.astype(np.float64) \
        + al.astype(np.float64) @ bh.astype(np.float64)
    return acc.astype(np.float32)


def logmel(pcm, n_mels, mode):
    x = np.pad(OF.pad_or_trim(pcm), (200, 200), mode="reflect").astype(np.float32)
    frames = np.lib.stride_tricks.as_strided(x, shape=(3000, 400), strides=(x.strides[0] * 160, x.strides[0]))
    fr = (frames * OF.hann_window().astype(np.float32)[None, :]).astype(np.float32)      # [3000, 400]
    n = np.arange(20)
    w20 = np.exp(-2j * np.pi * np.outer(n, n) / 20)
    wr, wi = w20.real.astype(np.float32), w20.imag.astype(np.float32)
    a = fr.reshape(3000, 20, 20).transpose(0, 2, 1)               # [f, n2, n1]:  x[20 n1 + n2]
    x1r, x1i = mma(a, wr, mode), mma(a, wi, mode)                 # [f, n2, k1]
    tw = np.exp(-2j * np.pi * np.outer(n, n) / 400)               # W400^(n2 k1)
    twr, twi = tw.real.astype(np.float32), tw.imag.astype(np.float32)
    yr = x1r * twr[None] - x1i * twi[None]                        # fp32 on the CUDA cores
    yi = x1r * twi[None] + x1i * twr[None]
    yr, yi = yr.transpose(0, 2, 1), yi.transpose(0, 2, 1)         # [f, k1, n2]
    zr = mma(yr, wr, mode) - mma(yi, wi, mode)                    # [f, k1, k2] -> bin k1 + 20 k2
    zi = mma(yr, wi, mode) + mma(yi, wr, mode)
    spec_r = zr.transpose(0, 2, 1).reshape(3000, 400)[:, :201]
    spec_i = zi.transpose(0, 2, 1).reshape(3000, 400)[:, :201]
    power = spec_r.astype(np.float32) ** 2 + spec_i.astype(np.float32) ** 2
    mel = np.maximum(1e-10, power @ OF.mel_filter_bank(n_mels).astype(np.float32))
    lg = np.log10(mel).T.astype(np.float32)
    lg = np.maximum(lg, lg.max() - 8.0)
    return (lg + 4.0) / 4.0


def main():
    out = {"gate": 1e-4, "what": __doc__.split("\n\n")[0].replace("\n", " ")}
    for name, fn in (("noise", OF.synth_noise), ("tones", OF.synth_tones), ("short", OF.synth_short)):
        pcm = fn()
        for n_mels in (80, 128):
            ref = OF.log_mel(pcm, n_mels)
            out[f"{name}_{n_mels}"] = {m: float(np.abs(logmel(pcm, n_mels, m) - ref).max()) for m in ("fp32", "tf32", "3xtf32")}
    worst = {m: max(v[m] for k, v in out.items() if isinstance(v, dict)) for m in ("fp32", "tf32", "3xtf32")}
    out["worst"] = worst
    out["verdict"] = {m: ("passes" if worst[m] <= 1e-4 else "FAILS") + " the 1e-4 gate" for m in worst}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
