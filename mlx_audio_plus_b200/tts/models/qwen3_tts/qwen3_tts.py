"""Drop-in for mel_spectrogram of mlx_audio/tts/models/qwen3_tts/qwen3_tts.py:33-90 — the one front-end
the reference pins with numeric golden vectors (tts/tests/test_qwen3_tts.py:157-329)."""
from __future__ import annotations

from ...._arrays import emit
from ...._wrap import as_batch, reflect_pad_rows, run_frontend
from .... import _lib as L
from ....dsp import hanning, mel_filters


def mel_spectrogram(audio, n_fft: int = 1024, num_mels: int = 128, sample_rate: int = 24000,
                    hop_size: int = 256, win_size: int = 1024, fmin: float = 0.0, fmax: float = 12000.0):
    ing, _ = as_batch(audio)
    fb = mel_filters(sample_rate=sample_rate, n_fft=n_fft, n_mels=num_mels, f_min=fmin, f_max=fmax,
                     norm="slaney", mel_scale="slaney")
    ing = reflect_pad_rows(ing, (n_fft - hop_size) // 2)  # manual reflect pad, then center=False
    out = run_frontend(
        ing, hanning(win_size), fb, n_fft=n_fft, hop=hop_size, center=False,
        spec_kind=L.SPEC_SQRT_POWER_EPS, spec_eps=1e-9, log_kind=L.LOG_LN, guard_kind=L.GUARD_MAX, guard_eps=1e-5)
    return emit(ing, out)  # (B, T, M)
