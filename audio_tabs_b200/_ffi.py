"""ctypes binding of libb200spec.so (include/b200spec.h).

The library is the product path; there is no Python or CPU fallback.  If the shared object is
missing, or a call returns a negative status, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

ABI_VERSION = 7
MAX_RES = 4
MAX_DIFF_FRAMES = 16

OK, ERR_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_ARCH = 0, -1, -2, -3, -4
F32, I16 = 0, 1
CLIP_NONFINITE = 1
END_NORMAL, END_EXTEND = 0, 1

c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class ResDesc(C.Structure):
    _fields_ = [
        ("frame_size", C.c_int32),
        ("hop_size", C.c_double),
        ("origin", C.c_int32),
        ("window", c_float_p),
        ("num_bands", C.c_int32),
        ("band_start", c_int32_p),
        ("band_len", c_int32_p),
        ("band_woff", c_int32_p),
        ("weights", c_float_p),
        ("log_enabled", C.c_int32),
        ("mul", C.c_float),
        ("add", C.c_float),
        ("diff_frames", C.c_int32),
        ("positive_diffs", C.c_int32),
        ("num_classes", C.c_int32),
        ("proj_off", c_int32_p),
        ("proj_band", c_int32_p),
        ("proj_weight", c_float_p),
        ("diff_max_bins", C.c_int32),
        ("power", C.c_int32),
        ("log_scale", C.c_float),
        ("log_floor", C.c_float),
        ("circular_shift", C.c_int32),
        ("include_nyquist", C.c_int32),
    ]


class PlanDesc(C.Structure):
    _fields_ = [
        ("device", C.c_int32),
        ("dtype", C.c_int32),
        ("channels", C.c_int32),
        ("num_res", C.c_int32),
        ("res", ResDesc * MAX_RES),
    ]


class OutDesc(C.Structure):
    _fields_ = [
        ("d_out", C.c_void_p),
        ("ld_out", C.c_int64),
        ("col_spec", C.c_int32),
        ("col_diff", C.c_int32),
        ("d_flux", C.c_void_p),
        ("d_proj", C.c_void_p),
        ("ld_proj", C.c_int64),
        ("d_clip_scale", C.c_void_p),
        ("d_clip_status", C.c_void_p),
    ]


# every symbol include/b200spec.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("b200spec_abi_version", C.c_int, []),
    ("b200spec_last_error", C.c_char_p, []),
    ("b200spec_num_frames", C.c_int, [C.c_int64, C.c_double, C.c_int, c_int64_p]),
    ("b200spec_frame_start", C.c_int, [C.c_int64, C.c_double, C.c_int32, C.c_int32, c_int64_p]),
    ("b200spec_plan_create", C.c_int, [C.POINTER(PlanDesc), C.POINTER(C.c_void_p)]),
    ("b200spec_plan_destroy", C.c_int, [C.c_void_p]),
    ("b200spec_workspace_bytes", C.c_size_t, [C.c_int32]),
    ("b200spec_stft", C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    ("b200spec_spectrogram", C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                       C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    ("b200spec_logfilt", C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                   C.c_int64, C.POINTER(OutDesc), C.c_void_p, C.c_size_t, C.c_void_p]),
    ("b200spec_logfilt_multi", C.c_int, [C.c_void_p, C.c_int32, c_int32_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                         C.c_int64, C.POINTER(OutDesc), C.c_void_p, C.c_size_t, C.c_void_p]),
    ("b200spec_logfilt_multi_supported", C.c_int, [C.c_void_p, C.c_int32, c_int32_p]),
    ("b200spec_clip_peak", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_int32,
                                     C.c_void_p, C.c_void_p]),
    ("b200spec_context_stack", C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_int64,
                                         C.c_int32, C.c_void_p, C.c_void_p]),
    ("b200spec_onset_envelope", C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_int64,
                                          C.c_int32, C.c_float, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                          C.c_void_p]),
    ("b200spec_magnitude", C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    ("b200spec_filter_log", C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int64, C.c_int32,
                                      C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    ("b200spec_diff_flux_chroma", C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                                            C.c_int64, C.POINTER(OutDesc), C.c_void_p]),
    ("b200spec_plan_num_res", C.c_int, [C.c_void_p]),
    ("b200spec_plan_num_bands", C.c_int, [C.c_void_p, C.c_int32]),
    ("b200spec_plan_filterbank_layout", C.c_int, [C.c_void_p, C.c_int32, c_int32_p]),
    ("b200spec_launch_count", C.c_int64, []),
    ("b200spec_task_plan", C.c_int, [C.c_int32, C.c_int32, C.c_double, C.c_int64, C.c_int32, C.c_int32, c_int32_p]),
]

_LIB = None


def library_path() -> Path:
    env = os.environ.get("B200SPEC_LIB")
    if env:
        return Path(env)
    return Path(__file__).resolve().parent / "lib" / "libb200spec.so"


def lib():
    """Load libb200spec.so once; raise if it is not built (no fallback exists)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not path.exists():
        raise RuntimeError(
            f"{path} is missing: build it with `python -m audio_tabs_b200.build` "
            "(the CUDA library is the only implementation of this path; there is no CPU fallback)")
    handle = C.CDLL(str(path))
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(handle, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    ver = handle.b200spec_abi_version()
    if ver != ABI_VERSION:
        raise RuntimeError(f"libb200spec ABI {ver} != binding ABI {ABI_VERSION}")
    _LIB = handle
    return handle


class B200SpecError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libb200spec error {code}: {message}")
        self.code = code


def check(code: int) -> None:
    """Map a negative status to the exception type madmom would raise for the same mistake."""
    if code >= 0:
        return
    msg = lib().b200spec_last_error().decode("utf-8", "replace")
    if code in (ERR_ARG, ERR_UNSUPPORTED):
        raise ValueError(msg)
    raise B200SpecError(code, msg)


def num_frames(n_samples: int, hop_size: float, end: str = "normal") -> int:
    if end == "normal":
        mode = END_NORMAL
    elif end == "extend":
        mode = END_EXTEND
    else:
        raise ValueError("end of signal handling '%s' unknown" % end)
    out = C.c_int64(0)
    check(lib().b200spec_num_frames(int(n_samples), float(hop_size), mode, C.byref(out)))
    return int(out.value)


def frame_start(index: int, hop_size: float, frame_size: int, origin: int = 0) -> int:
    out = C.c_int64(0)
    check(lib().b200spec_frame_start(int(index), float(hop_size), int(frame_size), int(origin), C.byref(out)))
    return int(out.value)


def launch_count() -> int:
    return int(lib().b200spec_launch_count())
