"""ttasr — B200-native drop-in for the log-mel front end + Whisper encoder hot path of Taiwan-Tongues-ASR-CE.

Host side of the C ABI in include/ttasr_abi.h (lib/libttasr_b200.so, hand-written sm_100a CUDA).  There is no CPU
or PyTorch fallback: importing is cheap, but any compute call without the built library or without a B200 raises.
"""
from ._lib import TtasrError, abi_version, library_path  # noqa: F401
from . import decode_handoff  # noqa: F401
from .encoder import B200WhisperEncoder, EncoderConfig  # noqa: F401
from .feature_extractor import B200WhisperFeatureExtractor  # noqa: F401
from .ingest import B200AudioIngest, resample_poly_filter  # noqa: F401
from .pipeline import B200LogMelEncoder, GraphedLogMelEncoder  # noqa: F401

__all__ = [
    "B200WhisperFeatureExtractor", "B200WhisperEncoder", "EncoderConfig", "B200LogMelEncoder", "GraphedLogMelEncoder", "B200AudioIngest", "resample_poly_filter",
    "TtasrError", "abi_version", "library_path",
]
