"""Run GPU test groups each in its own process (a trapped kernel poisons only its own CUDA context), with a
timeout, and write a summary + per-group logs under gpurun_out/.   python tools/gpu_probe.py [--only name,...]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
GROUPS = [
    ("frontend", ["tests/test_gpu_frontend.py"]),
    ("layernorm", ["tests/test_gpu_ops.py", "-k", "layernorm"]),
    ("gemm_cg1_small", ["tests/test_gpu_ops.py", "-k", "test_gemm and cg1 and m128n128k64"]),
    ("gemm_cg1", ["tests/test_gpu_ops.py", "-k", "test_gemm and cg1"]),
    ("gemm_cg2_small", ["tests/test_gpu_ops.py", "-k", "test_gemm and cg2 and m128n128k64"]),
    ("gemm_cg2", ["tests/test_gpu_ops.py", "-k", "test_gemm and cg2"]),
    ("attention_small", ["tests/test_gpu_ops.py", "-k", "test_attention and b1t128h1"]),
    ("attention", ["tests/test_gpu_ops.py", "-k", "attention"]),
    ("bad_args", ["tests/test_gpu_ops.py", "-k", "bad_arguments"]),
    ("encoder", ["tests/test_gpu_encoder.py"]),
    ("ingest", ["tests/test_ingest.py"]),
    ("decode_handoff", ["tests/test_decode_handoff.py"]),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    only = None
    if "--only" in sys.argv:
        only = set(sys.argv[sys.argv.index("--only") + 1].split(","))
    summary = {}
    for name, args in GROUPS:
        if only and name not in only:
            continue
        t0 = time.time()
        cmd = [sys.executable, "-m", "pytest", "-m", "gpu", "-q", "--no-header", "-p", "no:cacheprovider", *args]
        try:
            r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=420)
            rc, text = r.returncode, r.stdout + r.stderr
        except subprocess.TimeoutExpired as e:
            rc, text = -999, (e.stdout or b"").decode(errors="replace") + "\nTIMEOUT"
        with open(os.path.join(OUT, f"probe_{name}.log"), "w") as f:
            f.write(text)
        tail = [l for l in text.strip().splitlines() if l.strip()][-1:] or [""]
        summary[name] = {"rc": rc, "seconds": round(time.time() - t0, 1), "tail": tail[0][:200]}
        print(f"[probe] {name:18s} rc={rc:5d} {summary[name]['seconds']:7.1f}s  {tail[0][:150]}", flush=True)
    with open(os.path.join(OUT, "probe_summary.json"), "w") as f:
        json.dump(summary, f, indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
