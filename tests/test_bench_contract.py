"""bench.py's output contract (the driver parses ONE JSON line per arm): every required key present, on a small workload
so the check is fast.  The CPU arm runs anywhere; the GPU arm runs the full default path — including the cpu_baseline
leg, the front-end roofline passes and the library comparison — exactly as the driver invokes it."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
          "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(args, timeout):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                       timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    line = _run(["--impl", "reference", "--workload", "tiny", "--steps", "2", "--warmup", "1"], 600)
    assert COMMON <= set(line) and line["impl"] == "reference" and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert 2 <= line["reference_step"]["chunks_per_step"] <= 8
    assert str(line["reference_step"]["chunks_per_step"]) in line["cpu_baseline"]["sample"]


@pytest.mark.gpu
def test_default_arm_line(cuda_device):
    line = _run(["--workload", "tiny", "--batch", "8", "--steps", "3", "--warmup", "3"], 900)
    need = COMMON | {"gpu_launches", "clocks", "roofline", "gemm_roofline", "frontend_roofline", "kernels", "rank_probe",
                     "per_rank", "gpu_library_baseline"}
    assert need <= set(line), sorted(need - set(line))
    assert line["value"] > 0 and line["e2e"]["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 8 * 480000 * 4
    assert line["gpu_launches"] > 0 and line["dtype"] == "bf16" and line["n_gpus"] == 1
    rf = line["roofline"]
    assert rf["bound"] == "tensor" and rf["unit"] == "TFLOP/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    fr = line["frontend_roofline"]
    assert fr["bound"] == "hbm" and 0 < fr["frac"] < 1.5 and {"alone", "production_in_step"} <= set(fr)
    cb = line["cpu_baseline"]
    assert cb["value"] > 0 and cb["cores"] >= 1 and cb["kind"] in ("reference", "port")
    assert line["rank_probe"]["identical_on_all_ranks"] is True and len(line["per_rank"]) == 1
