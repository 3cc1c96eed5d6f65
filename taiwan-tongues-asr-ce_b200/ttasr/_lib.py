"""ctypes binding of include/ttasr_abi.h.  The product path fails loudly when the CUDA library is missing."""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# TTASR_LIB_PATH selects an experiment build of the same ABI (tools/, profiling); the default is the in-tree library
_LIB_PATH = os.environ.get("TTASR_LIB_PATH") or os.path.join(os.path.dirname(_HERE), "lib", "libttasr_b200.so")

TTASR_OK = 0
PCM_F32, PCM_I16 = 0, 1
FEATS_F32_MEL_MAJOR, FEATS_BF16_TIME_MAJOR = 0, 1
OUT_BF16, OUT_F32 = 0, 1
RESIDUAL_AUTO, RESIDUAL_F32, RESIDUAL_SPLIT, RESIDUAL_BF16 = -1, 0, 1, 2
RESIDUAL_MODES = {None: RESIDUAL_AUTO, "auto": RESIDUAL_AUTO, "f32": RESIDUAL_F32, "split": RESIDUAL_SPLIT,
                  "bf16": RESIDUAL_BF16}
PROFILE_KINDS = 9
ERROR_NAMES = {-1: "TTASR_E_ARG", -2: "TTASR_E_SHAPE", -3: "TTASR_E_ARCH", -4: "TTASR_E_CUDA", -5: "TTASR_E_NOMEM"}


class TtasrError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {message}")
        self.code = code


class EncoderCfg(C.Structure):
    _fields_ = [("d_model", C.c_int), ("n_layers", C.c_int), ("n_heads", C.c_int), ("ffn_dim", C.c_int),
                ("n_mels", C.c_int), ("n_ctx", C.c_int)]


class LayerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "ln1_g", "ln1_b", "wq", "bq", "wk", "wv", "bv", "wo", "bo", "ln2_g", "ln2_b", "w1", "b1", "w2", "b2")]


class Weights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "conv1_w", "conv1_b", "conv2_w", "conv2_b", "pos", "ln_post_g", "ln_post_b")] + [
        ("layers", C.POINTER(LayerWeights))]


# name -> (restype, argtypes); must list every symbol include/ttasr_abi.h declares (tests check this)
PROTOTYPES = {
    "ttasr_abi_version": (C.c_int, []),
    "ttasr_last_error": (C.c_char_p, []),
    "ttasr_device_check": (C.c_int, [C.c_int]),
    "ttasr_frontend_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                        C.POINTER(C.c_void_p)]),
    "ttasr_frontend_run": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int, C.c_void_p]),
    "ttasr_frontend_run_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int, C.c_float, C.c_void_p]),
    "ttasr_frontend_max_batch": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "ttasr_frontend_mel_mode": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "ttasr_frontend_destroy": (None, [C.c_void_p]),
    "ttasr_ingest_create": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "ttasr_ingest_out_len": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "ttasr_ingest_run": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_int64,
                                   C.c_void_p]),
    "ttasr_ingest_frame_energy": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
    "ttasr_ingest_gather_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "ttasr_ingest_destroy": (None, [C.c_void_p]),
    "ttasr_encoder_create": (C.c_int, [C.POINTER(EncoderCfg), C.POINTER(Weights), C.POINTER(C.c_void_p)]),
    "ttasr_encoder_create_ex": (C.c_int, [C.POINTER(EncoderCfg), C.POINTER(Weights), C.c_int, C.POINTER(C.c_void_p)]),
    "ttasr_encoder_workspace_bytes": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_size_t)]),
    "ttasr_encoder_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_size_t,
                                        C.c_void_p, C.c_int, C.c_void_p]),
    "ttasr_encoder_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "ttasr_encoder_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "ttasr_encoder_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int]),
    "ttasr_encoder_profile_kind_name": (C.c_char_p, [C.c_int]),
    "ttasr_encoder_destroy": (None, [C.c_void_p]),
    "ttasr_op_gemm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ttasr_op_gemm_split": (C.c_int, [C.c_void_p] * 8 + [C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "ttasr_op_gemm_lnfold": (C.c_int, [C.c_void_p] * 5 + [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int,
                                       C.c_float, C.c_int, C.c_void_p]),
    "ttasr_op_conv_stem": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int] + [C.c_void_p] * 8),
    "ttasr_op_layernorm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                     C.c_void_p]),
    "ttasr_op_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
}

_lib = None
_lock = threading.Lock()


def library_path() -> str:
    return _LIB_PATH


def lib() -> C.CDLL:
    """Load libttasr_b200.so (once).  Raises if it has not been built: there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(_LIB_PATH):
                raise TtasrError(-4, f"{_LIB_PATH} is missing: build it with "
                                     f"`python {os.path.join(os.path.dirname(_HERE), 'build.py')}` "
                                     "(needs nvcc; the library is sm_100a-only and has no CPU fallback)")
            handle = C.CDLL(_LIB_PATH)
            for name, (res, args) in PROTOTYPES.items():
                fn = getattr(handle, name)
                fn.restype = res
                fn.argtypes = args
            _lib = handle
    return _lib


def abi_version() -> int:
    return int(lib().ttasr_abi_version())


def check(rc: int) -> None:
    if rc != TTASR_OK:
        msg = lib().ttasr_last_error()
        raise TtasrError(rc, msg.decode("utf-8", "replace") if msg else "")


def require_cuda_tensor(t, name: str):
    import torch

    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TtasrError(-1, f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise TtasrError(-1, f"{name} must be contiguous")
    return t


def current_stream_ptr(device=None) -> int:
    import torch

    return int(torch.cuda.current_stream(device).cuda_stream)
