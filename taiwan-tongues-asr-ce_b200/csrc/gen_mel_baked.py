#!/usr/bin/env python
"""Generates csrc/mel_baked.inc: the mel projection of the 80-filter Whisper bank (slaney triangles over the 201 bins
of a 400-point FFT at 16 kHz — HF WhisperFeatureExtractor.mel_filters, feature_extraction_whisper.py:90-98) as
straight-line code per warp of the front-end kernel, weights as immediates.

Why: the generic mel pass of frontend_logmel.cu interprets a host-built program (one op per frequency bin: constant-bank
loads, a flag test and a branch per op — 14 issue slots per op for 2 useful FFMAs, 37 % of the kernel's instructions).
For a bank known at compile time the same walk is LDS + 2 FFMA per bin with immediate weights and offsets.  The kernel
uses the baked code only when the bank passed to ttasr_frontend_create is bit-identical to the one baked here (FNV-1a
hash of the fp32 table); any other bank runs the generic program.

Measured on the B200 (64 chunks, fp32 out): 80 filters 0.173 -> 0.154 ms.  The 128-filter bank is NOT baked: its code
(1900 instructions per output mode, ten different streams fetched at once) no longer fits the SM's instruction cache —
22 % fewer instructions executed, the same 0.60 ms per 256 chunks, the mel pass stalled on instruction fetch
(profiles/r2_ncu_full_summaries.md) — so BANKS below stays (80,); generate(banks=(80, 128)) still works for experiments.

    python gen_mel_baked.py            # rewrites mel_baked.inc next to this file
    python gen_mel_baked.py --stdout   # prints it (tests/test_host_logic.py checks the committed file is current)
"""
from __future__ import annotations

import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from ttasr import mel as M  # noqa: E402

N_WARPS = 10       # kMelWarps
N_FREQ = 201
EMIT_COST = 2.5    # issue slots of completing one filter, in units of one bin walked (3 slots)


def fnv1a64(data: bytes) -> int:
    h = 0xCBF29CE484222325
    for b in data:
        h ^= b
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def hexfloat(x: np.float32) -> str:
    """exact C literal of an fp32 value"""
    bits = struct.unpack("<I", struct.pack("<f", float(x)))[0]
    return f"__uint_as_float(0x{bits:08x}u)"


def partition(first, last, n_mels):
    """contiguous filter ranges starting at even filters, balanced by bins walked + completions"""
    cost = []
    for m in range(n_mels):
        prev_last = last[m - 1] if m > 0 else -1
        cost.append(EMIT_COST + max(0, last[m] - max(prev_last, first[m] - 1)))
    total = sum(cost)
    bounds = [0]
    acc = 0.0
    for m in range(n_mels):
        if len(bounds) < N_WARPS and m % 2 == 0 and m > bounds[-1] and acc >= total * len(bounds) / N_WARPS:
            bounds.append(m)
        acc += cost[m]
    while len(bounds) <= N_WARPS:
        bounds.append(n_mels)
    return bounds


def gen_bank(n_mels: int) -> str:
    fb = M.slaney_mel_filters(n_mels).astype(np.float32)  # [201, n_mels]
    assert fb.shape == (N_FREQ, n_mels) and n_mels % 2 == 0
    nz = fb != 0
    first = [int(np.argmax(nz[:, m])) for m in range(n_mels)]
    last = [int(N_FREQ - 1 - np.argmax(nz[::-1, m])) for m in range(n_mels)]
    for m in range(n_mels):
        assert nz[:, m].any(), f"empty filter {m}"
        assert nz[first[m]:last[m] + 1, m].all(), f"filter {m} is not contiguous"
    for k in range(N_FREQ):
        ms = np.nonzero(nz[k])[0]
        assert len(ms) <= 2 and (len(ms) < 2 or ms[1] == ms[0] + 1), f"bin {k} feeds {ms}"
    bounds = partition(first, last, n_mels)
    out = []
    out.append(f"// ---- {n_mels} filters; warp ranges {bounds}")
    out.append(f"template <> struct MelBaked<{n_mels}> {{")
    out.append(f"  static constexpr unsigned long long kHash = 0x{fnv1a64(fb.tobytes()):016x}ull;  // FNV-1a 64 of the fp32 [201, {n_mels}] table")
    out.append("  template <typename Out>")
    out.append("  static __device__ __forceinline__ void run(int wrp, const float* pw, Out& o) {")
    out.append("    float p;")
    out.append("    switch (wrp) {")
    for w in range(N_WARPS):
        m0, m1 = bounds[w], bounds[w + 1]
        if m0 >= m1:
            continue
        out.append(f"      case {w}: {{  // filters [{m0}, {m1})")
        k0 = min(first[m] for m in range(m0, m1))
        k1 = max(last[m] for m in range(m0, m1))
        started = set()
        done = set()
        body = []

        def finish_upto(k):
            for m in range(m0, m1):
                if m not in done and last[m] < k:
                    done.add(m)
                    body.append(f"const float y{m} = o.y(f{m});")
                    if m % 2 == 1:
                        body.append(f"o.emit2({m - 1}, y{m - 1}, y{m});")

        for k in range(k0, k1 + 1):
            finish_upto(k)
            ms = [m for m in range(m0, m1) if nz[k, m]]
            if not ms:
                continue
            body.append(f"p = pw[{k} * kPwPitch];")
            for m in ms:
                wv = hexfloat(fb[k, m])
                if m in started:
                    body.append(f"f{m} = fmaf({wv}, p, f{m});")
                else:
                    started.add(m)
                    body.append(f"float f{m} = {wv} * p;")
        finish_upto(N_FREQ + 1)
        assert done == set(range(m0, m1))
        for line in body:
            out.append("        " + line)
        out.append("      } break;")
    out.append("      default: break;")
    out.append("    }")
    out.append("  }")
    out.append("};")
    return "\n".join(out)


BANKS = (80,)


def generate(banks=BANKS) -> str:
    head = ("// GENERATED by csrc/gen_mel_baked.py — do not edit.  Mel projection of the Whisper filter banks as straight-line\n"
            "// code per warp (lane = frame, pw = the lane's column of the [bin][frame] power-spectrum buffer).\n"
            "// Included by frontend_logmel.cu inside namespace ttasr::{anonymous}; kPwPitch is defined there.\n"
            "template <int N_MELS> struct MelBaked;\n")
    return head + "\n".join(gen_bank(n) for n in banks) + "\n"


if __name__ == "__main__":
    text = generate()
    if "--stdout" in sys.argv:
        sys.stdout.write(text)
    else:
        with open(os.path.join(HERE, "mel_baked.inc"), "w") as f:
            f.write(text)
        print(f"wrote mel_baked.inc ({len(text.splitlines())} lines)")
