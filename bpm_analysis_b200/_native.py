"""ctypes binding of libbpm_b200.so (include/bpm_b200.h).

There is no CPU fallback: if the shared library is missing or cannot be loaded
every entry point raises ``NativeLibraryError``.  PyTorch is used only to own
device memory and streams; the signatures below take raw device pointers.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BPM_B200_LIB") or os.path.join(HERE, "libbpm_b200.so")   # override: diagnostic builds
ABI_VERSION = 2
DESIGN_HEADER_WORDS = 312

PCM_DTYPES = {np.dtype(np.int16): 0, np.dtype(np.int32): 1, np.dtype(np.uint8): 2,
              np.dtype(np.float32): 3, np.dtype(np.float64): 4}

ITEM_DTYPE = np.dtype([("in_off", np.int64), ("n_in", np.int64), ("m_off", np.int64), ("m", np.int64)])


class NativeLibraryError(RuntimeError):
    pass


class BpmError(RuntimeError):
    def __init__(self, code: int, what: str):
        super().__init__(f"libbpm_b200: {what} (code {code})")
        self.code = code


class StageAConfig(C.Structure):
    _fields_ = [("stride", C.c_int64), ("block", C.c_int64), ("pcm_dtype", C.c_int32), ("channels", C.c_int32),
                ("env_window", C.c_int32), ("distance", C.c_int32), ("noise_window", C.c_int32),
                ("want_debug_wav", C.c_int32), ("trough_prom_q", C.c_double), ("peak_prom_q", C.c_double),
                ("floor_q", C.c_double), ("rejection_multiplier", C.c_double), ("smoothing_factor", C.c_double)]


class StageAOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("filtered", "envelope", "absmax", "debug_wav", "floor", "troughs",
                                          "trough_count", "peaks", "peak_count", "strength", "deviation",
                                          "smoothed_dev", "trough_total", "floor_mode")]


_P, _I, _L, _D, _Z = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_size_t

_SIGNATURES = {
    "bpm_abi_version": (C.c_int, []),
    "bpm_error_string": (C.c_char_p, [_I]),
    "bpm_launch_count": (_L, []),
    "bpm_profile_begin": (_I, [_P]),
    "bpm_profile_end": (_I, [C.c_char_p, _Z]),
    "bpm_frontend_workspace_bytes": (_Z, [_L, _I]),
    "bpm_frontend": (_I, [_P, _I, _I, _P, _P, _I, _L, _P, _P, _L, _I, _P, _P, _P, _P, _Z, _P]),
    "bpm_debug_wav": (_I, [_P, _P, _P, _P, _I, _P, _P]),
    "bpm_gather_frames": (_I, [_P, _I, _I, _P, _P, _I, _L, _P, _P]),
    "bpm_copy_frames": (_I, [_P, _I, _I, _P, _I, _L, _P, _P]),
    "bpm_quantile_workspace_bytes": (_Z, [_I]),
    "bpm_quantile": (_I, [_P, _P, _P, _I, _D, _P, _P, _Z, _P]),
    "bpm_find_peaks_workspace_bytes": (_Z, [_L, _I]),
    "bpm_find_peaks": (_I, [_P, _I, _P, _P, _I, _P, _P, _I, _P, _P, _P, _Z, _P]),
    "bpm_rolling_floor_workspace_bytes": (_Z, [_L, _I]),
    "bpm_rolling_floor": (_I, [_P, _P, _P, _P, _P, _I, _I, _D, _P, _P, _Z, _P]),
    "bpm_noise_floor_workspace_bytes": (_Z, [_L, _I]),
    "bpm_noise_floor": (_I, [_P, _P, _P, _I, _I, _D, _D, _I, _D, _P, _P, _P, _P, _P, _P, _Z, _P]),
    "bpm_sanitize_troughs_workspace_bytes": (_Z, [_L, _I]),
    "bpm_sanitize_troughs": (_I, [_P, _P, _P, _P, _P, _P, _I, _D, _P, _P, _P, _Z, _P]),
    "bpm_raw_peaks_workspace_bytes": (_Z, [_L, _I]),
    "bpm_raw_peaks": (_I, [_P, _P, _P, _P, _I, _I, _D, _P, _P, _P, _Z, _P]),
    "bpm_peak_metrics": (_I, [_P, _P, _P, _P, _P, _P, _I, _D, _P, _P, _P, _P]),
    "bpm_peak_trough_noise": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _D, _D, _P, _P, _P, _P, _P]),
    "bpm_bpm_series": (_I, [_P, _P, _P, _I, _I, _L, _P, _P, _P, _P, _P, _P]),
    "bpm_steepest_slope_workspace_bytes": (_Z, [_L, _I]),
    "bpm_steepest_slope": (_I, [_P, _P, _P, _P, _P, _I, _I, _D, _P, _P, _Z, _P]),
    "bpm_windowed_hrv": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "bpm_cast_f32": (_I, [_P, _P, _L, _P]),
    "bpm_key_histogram": (_I, [_P, _L, _I, _I, C.c_uint64, _P, _P, _P]),
    "bpm_key_collect": (_I, [_P, _L, _I, C.c_uint64, _P, _L, _P, _P, _P]),
    "bpm_key_pick": (_I, [_P, _I, _P, _P]),
    "bpm_key_finish": (_I, [_P, _I, _L, _P, _D, _P, _P, _P]),
    "bpm_chunk_proof": (_I, [_P, _P, _P, _P, _P, _P, _L, _L, _L, _L, _L, _I, _I, _L, _I, _I, _P, _P]),
    "bpm_find_peaks_chunk": (_I, [_P, _I, _P, _P, _I, _P, _P, _L, _L, _I, _I, _P, _P, _P, _P, _P, _Z, _P]),
    "bpm_noise_floor_chunk_workspace_bytes": (_Z, [_L]),
    "bpm_noise_floor_chunk": (_I, [_P, _P, _P, _I, _P, _D, _I, _D, _L, _L, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
    "bpm_chunk_pack": (_I, [_P, _P, _P, _P, _L, _L, _L, _P, _P]),
    "bpm_chunk_unpack": (_I, [_P, _P, _I, _L, _L, _P, _P, _P, _P]),
    "bpm_peak_strength": (_I, [_P, _P, _P, _P, _P, _P, _I, _P, _P]),
    "bpm_deviation_series": (_I, [_P, _P, _P, _P, _I, _D, _P, _P, _P]),
    "bpm_stage_a_workspace_bytes": (_Z, [_L, _I]),
    "bpm_stage_a": (_I, [_P, _P, _P, _I, _P, _P, _L, C.POINTER(StageAConfig), C.POINTER(StageAOutputs), _P, _Z, _P]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lock = threading.Lock()
_lib = None


def load_library(path: str = LIB_PATH):
    """Load (once) and return the ctypes handle; raises NativeLibraryError if unavailable."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(path):
            raise NativeLibraryError(
                f"{path} not found: build it with `python -m bpm_analysis_b200.build` "
                "(there is no CPU fallback for the front end)")
        try:
            lib = C.CDLL(path)
        except OSError as e:                                   # e.g. libcudart missing
            raise NativeLibraryError(f"cannot load {path}: {e}") from e
        for name, (res, args) in _SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as e:
                raise NativeLibraryError(f"{path} does not export {name}") from e
            fn.restype, fn.argtypes = res, args
        if lib.bpm_abi_version() != ABI_VERSION:
            raise NativeLibraryError(f"ABI mismatch: library {lib.bpm_abi_version()}, binding {ABI_VERSION}")
        _lib = lib
        return lib


def check(code: int) -> None:
    if code != 0:
        raise BpmError(code, load_library().bpm_error_string(code).decode())


def launch_count() -> int:
    return int(load_library().bpm_launch_count())
