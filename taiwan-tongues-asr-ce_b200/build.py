"""Builds lib/libttasr_b200.so (sm_100a only) from csrc/*.cu with nvcc.  Usage: python build.py [--force]

The library is built in-tree so that it travels with the repository snapshot to the GPU box; it is git-ignored.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libttasr_b200.so")
SOURCES = ["frontend_logmel.cu", "ingest_resample.cu", "gemm_sm100.cu", "attention_sm100.cu", "attention4_sm100.cu", "layernorm.cu", "ttasr_abi.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
    "-diag-suppress", "177,550",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h", ".inc")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode())
                    h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = True, defines: tuple = (), out: str | None = None) -> str:
    """defines / out: build an experiment variant (-D...) into another file without touching the default library."""
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = OBJDIR if not out else OBJDIR + "_" + os.path.basename(out)
    os.makedirs(objdir, exist_ok=True)
    lib = out or LIB
    stamp = os.path.join(objdir, "digest.txt")
    digest = _digest() + "|" + " ".join(defines)
    if not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read() == digest:
        return lib
    nvcc = _nvcc()

    def compile_one(src: str) -> str:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-I", CSRC, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC", "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    if verbose:
        print(f"built {lib} ({os.path.getsize(lib)} bytes)")
    return lib


if __name__ == "__main__":
    _defs = tuple(a[2:] for a in sys.argv[1:] if a.startswith("-D"))
    _out = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")), None)
    build(force="--force" in sys.argv, defines=_defs, out=_out)
