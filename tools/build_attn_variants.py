#!/usr/bin/env python
"""Build A/B variants of the library that differ only in the attention kernel's compile-time switches
(csrc/attention_sm100.cu, attention4_sm100.cu: TTASR_ATTN*), for tools/attn_ab.py.  Only those two are recompiled per variant; the
other objects are taken from the default build.   python tools/build_attn_variants.py  ->  lib/variants/attn_<name>.so"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "taiwan-tongues-asr-ce_b200")
sys.path.insert(0, PKG)
import build as B  # noqa: E402

V4 = "TTASR_ATTN_DEFAULT_VARIANT=4"
LM = "TTASR_ATTN_LATEMAX=1"
VARIANTS = {
    "v2": [],                                   # two softmax warpgroups, 128-key tiles (attention_sm100.cu), as shipped
    "base": ["TTASR_ATTN_LATEMAX=0"],           # ... with the whole row max taken before the sweep (the round-1 chain)
    "abl_max": ["TTASR_ATTN_ABLATE=1"],         # diagnostics (wrong results): one element of the chain removed at a time
    "abl_st": ["TTASR_ATTN_ABLATE=2"],
    "abl_ld": ["TTASR_ATTN_ABLATE=4"],
    "abl_pack": ["TTASR_ATTN_ABLATE=8"],
    "abl_sum": ["TTASR_ATTN_ABLATE=16"],
    "abl_all": ["TTASR_ATTN_ABLATE=15", "TTASR_ATTN_FAKE_EXP=8"],
    "poll": ["TTASR_ATTN_POLL=1"],              # mbarrier waits poll (test_wait) instead of suspending (try_wait)
    "poll_mma2": ["TTASR_ATTN_POLL=1", "TTASR_ATTN_TWO_MMA=1"],
    "poll_fake8": ["TTASR_ATTN_POLL=1", "TTASR_ATTN_FAKE_EXP=8"],
    "mma2": ["TTASR_ATTN_TWO_MMA=1"],           # one MMA-issuing warp per query tile
    "mma2_p0": ["TTASR_ATTN_TWO_MMA=1", "TTASR_ATTN_PRETOKEN=0"],
    "mma2_st": ["TTASR_ATTN_TWO_MMA=1", "TTASR_ATTN_SMSP_TOKEN=1"],
    "mma2_fake8": ["TTASR_ATTN_TWO_MMA=1", "TTASR_ATTN_FAKE_EXP=8"],
    "fake2": ["TTASR_ATTN_FAKE_EXP=2"],         # diagnostic (wrong results): 2 / 4 / 8 of every 8 exponentials are free
    "fake4": ["TTASR_ATTN_FAKE_EXP=4"],
    "fake8": ["TTASR_ATTN_FAKE_EXP=8"],
    "er": ["TTASR_ATTN_EARLY_RELEASE=1"],       # token handed over before the last quarter's exponentials
    "er_p0": ["TTASR_ATTN_EARLY_RELEASE=1", "TTASR_ATTN_PRETOKEN=0"],
    "er_st": ["TTASR_ATTN_EARLY_RELEASE=1", "TTASR_ATTN_SMSP_TOKEN=1"],
    "er_st_p0": ["TTASR_ATTN_EARLY_RELEASE=1", "TTASR_ATTN_SMSP_TOKEN=1", "TTASR_ATTN_PRETOKEN=0"],
    "st": ["TTASR_ATTN_SMSP_TOKEN=1"],          # token passed per scheduler (64-thread barriers) instead of per warpgroup
    "st_p2": ["TTASR_ATTN_SMSP_TOKEN=1", "TTASR_ATTN_PRETOKEN=2"],
    "st_p0": ["TTASR_ATTN_SMSP_TOKEN=1", "TTASR_ATTN_PRETOKEN=0"],
    "st_q2": ["TTASR_ATTN_SMSP_TOKEN=1", "TTASR_ATTN_POLY_Q0=2"],
    "st_lm": ["TTASR_ATTN_SMSP_TOKEN=1", LM],
    "lm": [LM],                                 # quarter-0 max first, the rest beside quarter 0's exponentials
    "lm_n": [LM, "TTASR_ATTN_SETMAXNREG=1"],    # ... softmax warps at 224 registers
    "lm_p2": [LM, "TTASR_ATTN_PRETOKEN=2"],     # ... two quarters outside the token
    "lm_p2n": [LM, "TTASR_ATTN_PRETOKEN=2", "TTASR_ATTN_SETMAXNREG=1"],
    "lm_q2": [LM, "TTASR_ATTN_POLY_Q0=2"],      # ... a quarter of quarter 0's exponentials on the FMA pipe
    "lm_p2q2": [LM, "TTASR_ATTN_PRETOKEN=2", "TTASR_ATTN_POLY_Q0=2", "TTASR_ATTN_POLY_Q1=2"],
    "lm_q4": [LM, "TTASR_ATTN_POLY_Q0=4"],      # ... half of quarter 0
    "v4": [V4],                                 # four softmax warpgroups, 64-key steps (attention4_sm100.cu)
    "v4p2": [V4, "TTASR_ATTN4_POLY8=2"],        # ... a quarter of the exponentials on the FMA pipe
    "v4p4": [V4, "TTASR_ATTN4_POLY8=4"],        # ... half
}
ATTN_SOURCES = ["attention_sm100.cu", "attention4_sm100.cu"]


def main():
    only = set(sys.argv[1:])
    B.build(verbose=False)
    outdir = os.path.join(PKG, "lib", "variants")
    os.makedirs(outdir, exist_ok=True)
    others = [os.path.join(B.OBJDIR, s.replace(".cu", ".o")) for s in B.SOURCES if s not in ATTN_SOURCES]
    for name, defs in VARIANTS.items():
        if only and name not in only:
            continue
        objs = []
        for src in ATTN_SOURCES:
            obj = os.path.join(outdir, f"attn_{name}_{src.replace('.cu', '.o')}")
            cmd = [B._nvcc(), *B.NVCC_FLAGS, *[f"-D{d}" for d in defs], "-I", B.CSRC, "-c",
                   os.path.join(B.CSRC, src), "-o", obj]
            subprocess.run(cmd, check=True)
            objs.append(obj)
        lib = os.path.join(outdir, f"attn_{name}.so")
        subprocess.run([B._nvcc(), "-shared", "-o", lib, *objs, *others, "-gencode", "arch=compute_100a,code=sm_100a",
                        "-Xcompiler", "-fPIC", "-lpthread", "-ldl", "-lrt"], check=True)
        for obj in objs:
            os.remove(obj)
        print("built", lib, defs)


if __name__ == "__main__":
    main()
