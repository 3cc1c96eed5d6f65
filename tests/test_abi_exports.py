"""The C-ABI library builds for sm_100a, loads without a GPU, and exports every symbol include/ttasr_abi.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ttasr_abi.h")


def _declared():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"TTASR_API\s+[\w\s\*]+?\b(ttasr_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = _declared()
    for must in ("ttasr_frontend_create", "ttasr_frontend_run", "ttasr_encoder_create", "ttasr_encoder_forward",
                 "ttasr_encoder_workspace_bytes", "ttasr_last_error", "ttasr_op_gemm", "ttasr_op_attention"):
        assert must in names
    assert len(names) >= 15


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in ttasr_abi.h but not exported"
    lib.ttasr_abi_version.restype = ctypes.c_int
    assert lib.ttasr_abi_version() == 1


def test_python_prototypes_cover_the_header(built_lib):
    from ttasr import _lib

    assert sorted(_lib.PROTOTYPES) == _declared()
    assert _lib.abi_version() == 1
    assert os.path.samefile(_lib.library_path(), built_lib)


def test_library_is_sm100a_only_and_uses_tensor_memory(built_lib):
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    elf = subprocess.run([cuobjdump, "-lelf", built_lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", elf))
    assert archs == {"sm_100a"}, archs
    sass = subprocess.run([cuobjdump, "-sass", built_lib], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UBLKCP"):  # tcgen05.mma, TMA load/store, tcgen05.ld, bulk copy
        assert mnemonic in sass, f"{mnemonic} missing from SASS"
    assert "HMMA.16816" not in sass  # no legacy mma.sync path


def test_no_gpu_means_a_loud_error_not_a_fallback(built_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU-only box")
    from ttasr import B200WhisperFeatureExtractor, TtasrError, _lib

    rc = _lib.lib().ttasr_device_check(0)
    assert rc != 0 and _lib.lib().ttasr_last_error()
    with pytest.raises(TtasrError):
        B200WhisperFeatureExtractor(80)([0.0] * 16000, sampling_rate=16000)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "taiwan-tongues-asr-ce_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f"{f} imports the oracle"
