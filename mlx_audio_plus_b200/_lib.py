"""ctypes binding of include/b200audio.h (the C-ABI drop-in boundary).

There is deliberately NO fallback: if libb200audio.so is missing the import fails loudly, and if no CUDA
device is present every compute call raises (B2A_ERR_CUDA).  Status codes are mapped back to the Python
exceptions the reference raises at the same places (dsp.py:109,126,132-136,175).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2A_LIB", os.path.join(_HERE, "lib", "libb200audio.so"))

# enums (mirrors include/b200audio.h)
WIN_HANN, WIN_HAMMING, WIN_BLACKMAN, WIN_BARTLETT = 0, 1, 2, 3
PAD_REFLECT, PAD_CONSTANT = 0, 1
SPEC_COMPLEX, SPEC_POWER, SPEC_MAGNITUDE, SPEC_SQRT_POWER_EPS = 0, 1, 2, 3
LOG_NONE, LOG_LOG10, LOG_LN = 0, 1, 2
GUARD_NONE, GUARD_MAX, GUARD_ADD = 0, 1, 2
CLAMP_NONE, CLAMP_CLIP_MAX, CLAMP_BATCH_MAX, CLAMP_FIXED = 0, 1, 2, 3
NORM_NONE, NORM_PER_FEATURE, NORM_GLOBAL = 0, 1, 2
LAYOUT_TM, LAYOUT_MT = 0, 1
ISTFT_NORM_WINDOW, ISTFT_NORM_WINDOW_SQ = 0, 1
ISTFT_DIV_WHERE, ISTFT_DIV_CLAMP = 0, 1
ISTFT_INPUT_COMPLEX, ISTFT_INPUT_POLAR = 0, 1

OK, ERR_INVALID_ARG, ERR_UNKNOWN_WINDOW, ERR_PAD_MODE, ERR_TOO_SHORT = 0, -1, -2, -3, -4
ERR_SHAPE, ERR_CUDA, ERR_UNSUPPORTED, ERR_NOMEM = -5, -6, -7, -8


class FrontendDesc(C.Structure):
    _fields_ = [
        ("n_fft", C.c_int32), ("hop", C.c_int32), ("center", C.c_int32), ("pad_mode", C.c_int32),
        ("window_len", C.c_int32), ("preemph", C.c_float), ("drop_last", C.c_int32),
        ("spec_kind", C.c_int32), ("spec_eps", C.c_float), ("n_mels", C.c_int32),
        ("log_kind", C.c_int32), ("guard_kind", C.c_int32), ("guard_eps", C.c_float),
        ("clamp_kind", C.c_int32), ("clamp_value", C.c_float), ("affine_add", C.c_float),
        ("affine_div", C.c_float), ("norm_kind", C.c_int32), ("norm_ddof", C.c_int32),
        ("norm_eps", C.c_float), ("out_layout", C.c_int32),
        ("frame_len", C.c_int32), ("frame_dc", C.c_int32), ("frame_preemph", C.c_float), ("dither", C.c_float),
        ("out_dtype", C.c_int32),
    ]


class ForwardArgs(C.Structure):
    _fields_ = [
        ("audio", C.c_void_p), ("clip_stride", C.c_int64), ("length", C.c_int64),
        ("valid_length", C.c_int64), ("pad_value", C.c_float), ("batch", C.c_int32),
        ("sample_offset", C.c_int64), ("frame_begin", C.c_int64), ("frame_count", C.c_int64),
        ("out", C.c_void_p), ("out_clip_stride", C.c_int64), ("clip_max", C.c_void_p),
        ("feat_sums", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("seed", C.c_uint64),
        ("audio_kind", C.c_int32), ("reserved0", C.c_int32),
    ]


class IstftDesc(C.Structure):
    _fields_ = [
        ("n_fft", C.c_int32), ("hop", C.c_int32), ("window_len", C.c_int32), ("center", C.c_int32),
        ("norm_kind", C.c_int32), ("div_kind", C.c_int32), ("trim_tail", C.c_int32),
        ("div_eps", C.c_float), ("input_form", C.c_int32), ("mag_clip_max", C.c_float), ("mag_clip_min_zero", C.c_int32),
        ("mag_log", C.c_int32),
    ]


class InverseArgs(C.Structure):
    _fields_ = [
        ("spec", C.c_void_p), ("spec_imag", C.c_void_p), ("clip_stride", C.c_int64),
        ("num_frames", C.c_int64), ("batch", C.c_int32), ("length", C.c_int64),
        ("out", C.c_void_p), ("out_clip_stride", C.c_int64),
    ]


class ResampleArgs(C.Structure):
    _fields_ = [
        ("inp", C.c_void_p), ("out", C.c_void_p), ("n_in", C.c_int64), ("in_clip_stride", C.c_int64),
        ("out_clip_stride", C.c_int64), ("batch", C.c_int32), ("channels", C.c_int32), ("in_kind", C.c_int32),
        ("mono", C.c_int32),
    ]


PCM_F32, PCM_I16 = 0, 1
DTYPE_F32, DTYPE_F16, DTYPE_BF16 = 0, 1, 2

# every symbol include/b200audio.h declares; tests/test_abi.py checks the library exports them all
SYMBOLS = {
    "b2a_version": (C.c_int, []),
    "b2a_last_error": (C.c_char_p, []),
    "b2a_device_count": (C.c_int, []),
    "b2a_window": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "b2a_mel_filters": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_void_p]),
    "b2a_stft_geometry": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "b2a_frame_source_index": (C.c_int64, [C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int]),
    "b2a_istft_geometry": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "b2a_frontend_create": (C.c_int, [C.POINTER(FrontendDesc), C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "b2a_launch_count": (C.c_ulonglong, []),
    "b2a_plan_destroy": (C.c_int, [C.c_void_p]),
    "b2a_frontend_out_frames": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "b2a_frontend_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int32]),
    "b2a_frontend_call_workspace_bytes": (C.c_size_t, [C.c_void_p, C.POINTER(ForwardArgs)]),
    "b2a_frontend_forward": (C.c_int, [C.c_void_p, C.POINTER(ForwardArgs), C.c_void_p]),
    "b2a_frontend_partial": (C.c_int, [C.c_void_p, C.POINTER(ForwardArgs), C.c_void_p]),
    "b2a_frontend_finalize": (C.c_int, [C.c_void_p, C.POINTER(ForwardArgs), C.c_int64, C.c_void_p]),
    "b2a_frontend_forward_host": (C.c_int, [C.c_void_p, C.POINTER(ForwardArgs)]),
    "b2a_frontend_dump_frames": (C.c_int, [C.c_void_p, C.POINTER(ForwardArgs), C.c_int, C.c_void_p]),
    "b2a_deltas": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "b2a_plan_kernel_name": (C.c_char_p, [C.c_void_p]),
    "b2a_istft_create": (C.c_int, [C.POINTER(IstftDesc), C.c_void_p, C.POINTER(C.c_void_p)]),
    "b2a_istft_out_len": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_int64)]),
    "b2a_istft_inverse": (C.c_int, [C.c_void_p, C.POINTER(InverseArgs), C.c_void_p]),
    "b2a_istft_inverse_host": (C.c_int, [C.c_void_p, C.POINTER(InverseArgs)]),
    "b2a_resampler_create": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.POINTER(C.c_void_p)]),
    "b2a_resampler_destroy": (C.c_int, [C.c_void_p]),
    "b2a_resampler_out_len": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "b2a_resample": (C.c_int, [C.c_void_p, C.POINTER(ResampleArgs), C.c_void_p]),
    "b2a_rows_pad_cast": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_int32,
                                    C.c_int32, C.c_void_p]),
    "b2a_lfr": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                          C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "b2a_transpose_pad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_void_p]),
    "b2a_rows_normalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_void_p]),
    "b2a_unwrap": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_float, C.c_float, C.c_void_p]),
    "b2a_cmvn_utterance": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_float, C.c_void_p, C.c_int32,
                                     C.c_void_p]),
    "b2a_measure_fp32_tflops": (C.c_int, [C.POINTER(C.c_double), C.c_void_p]),
    "b2a_measure_copy_gbs": (C.c_int, [C.POINTER(C.c_double), C.c_void_p]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"b200audio: {LIB_PATH} not found. Build it with `python -m mlx_audio_plus_b200.csrc.build` "
            "(or __graft_entry__.build()). There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class B2AError(RuntimeError):
    pass


def last_error() -> str:
    return (lib.b2a_last_error() or b"").decode()


def check(rc: int):
    """Map a b2a_status to the exception the reference raises at the same point."""
    if rc == OK:
        return
    msg = last_error()
    if rc in (ERR_UNKNOWN_WINDOW, ERR_PAD_MODE, ERR_TOO_SHORT, ERR_SHAPE, ERR_INVALID_ARG):
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == ERR_NOMEM:
        raise MemoryError(msg)
    raise B2AError(f"b200audio status {rc}: {msg}")
