"""Drop-in for mlx_audio/stt/models/voxtral_realtime/audio.py: periodic Hann evaluated in float32,
reflect pad, power, Slaney filterbank 0-8 kHz, log10, FIXED floor (global_log_mel_max - 8), (x+4)/4,
output (n_mels, frames)."""
from __future__ import annotations

import math

import numpy as np

from ...._arrays import emit, host_window
from ...._wrap import as_batch, run_frontend
from .... import _lib as L
from ....dsp import mel_filters


def compute_mel_filters(num_mel_bins: int = 128, window_size: int = 400, sample_rate: int = 16000) -> np.ndarray:
    """[freq, mel] filterbank (reference audio.py:19-38)."""
    fb = mel_filters(sample_rate=sample_rate, n_fft=window_size, n_mels=num_mel_bins, f_min=0, f_max=8000,
                     norm="slaney", mel_scale="slaney")
    return np.array(fb).T


def compute_mel_spectrogram(audio, mel_filters, window_size: int = 400, hop_length: int = 160,
                            global_log_mel_max: float = 1.5):
    """(L,) -> (mel_bins, frames) (reference audio.py:41-96)."""
    n = np.arange(window_size, dtype=np.float32)  # audio.py:60-61: float32 cosine
    window = (np.float32(0.5) * (np.float32(1.0) - np.cos(np.float32(2.0 * math.pi) * n / np.float32(window_size)))).astype(np.float32)
    fb = np.ascontiguousarray(host_window(np.asarray(mel_filters)).reshape(np.asarray(mel_filters).shape).T)  # -> (M, F)
    ing, was_1d = as_batch(audio)
    out = run_frontend(
        ing, window, fb, n_fft=window_size, hop=hop_length, center=True, pad_mode="reflect", drop_last=True,
        spec_kind=L.SPEC_POWER, log_kind=L.LOG_LOG10, guard_kind=L.GUARD_MAX, guard_eps=1e-10,
        clamp_kind=L.CLAMP_FIXED, clamp_value=float(global_log_mel_max - 8.0), affine_add=4.0, affine_div=4.0,
        out_layout=L.LAYOUT_MT)
    return emit(ing, out[0] if was_1d else out)
