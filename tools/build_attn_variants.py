#!/usr/bin/env python
"""Build A/B variants of the library that differ only in the attention kernel's compile-time switches
(csrc/attention_sm100.cu: TTASR_ATTN_*), for tools/attn_ab.py.  Only attention_sm100.cu is recompiled per variant; the
other objects are taken from the default build.   python tools/build_attn_variants.py  ->  lib/variants/attn_<name>.so"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "taiwan-tongues-asr-ce_b200")
sys.path.insert(0, PKG)
import build as B  # noqa: E402

N = "TTASR_ATTN_SETMAXNREG=1"
VARIANTS = {
    "base": [],
    "nreg": [N],
    "q0p8": [N, "TTASR_ATTN_POLY_Q0=8"],
    "all2": [N] + [f"TTASR_ATTN_POLY_Q{q}=2" for q in range(4)],
    "all4": [N] + [f"TTASR_ATTN_POLY_Q{q}=4" for q in range(4)],
    "q01p8_pre2": [N, "TTASR_ATTN_POLY_Q0=8", "TTASR_ATTN_POLY_Q1=8", "TTASR_ATTN_PRETOKEN=2"],
    "q0p8_rest2": [N, "TTASR_ATTN_POLY_Q0=8"] + [f"TTASR_ATTN_POLY_Q{q}=2" for q in (1, 2, 3)],
    "all4_pre4": [N, "TTASR_ATTN_PRETOKEN=4"] + [f"TTASR_ATTN_POLY_Q{q}=4" for q in range(4)],
    "all2_pre4": [N, "TTASR_ATTN_PRETOKEN=4"] + [f"TTASR_ATTN_POLY_Q{q}=2" for q in range(4)],
    "q0p8_nonreg": ["TTASR_ATTN_POLY_Q0=8"],
}


def main():
    only = set(sys.argv[1:])
    B.build(verbose=False)
    outdir = os.path.join(PKG, "lib", "variants")
    os.makedirs(outdir, exist_ok=True)
    others = [os.path.join(B.OBJDIR, s.replace(".cu", ".o")) for s in B.SOURCES if s != "attention_sm100.cu"]
    for name, defs in VARIANTS.items():
        if only and name not in only:
            continue
        obj = os.path.join(outdir, f"attn_{name}.o")
        cmd = [B._nvcc(), *B.NVCC_FLAGS, *[f"-D{d}" for d in defs], "-I", B.CSRC, "-c",
               os.path.join(B.CSRC, "attention_sm100.cu"), "-o", obj]
        subprocess.run(cmd, check=True)
        lib = os.path.join(outdir, f"attn_{name}.so")
        subprocess.run([B._nvcc(), "-shared", "-o", lib, obj, *others, "-gencode", "arch=compute_100a,code=sm_100a",
                        "-Xcompiler", "-fPIC", "-lpthread", "-ldl", "-lrt"], check=True)
        os.remove(obj)
        print("built", lib, defs)


if __name__ == "__main__":
    main()
