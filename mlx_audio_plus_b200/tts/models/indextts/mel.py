"""Drop-in for mlx_audio/tts/models/indextts/mel.py:6-37: symmetric "hann" of n_fft taps, hop forwarded (unlike the Vocos
twin), NO frame drop, magnitude, HTK filterbank without normalisation, ln(max(., 1e-5)), (1, T, n_mels)."""
from __future__ import annotations

from ...._arrays import emit
from ...._wrap import as_batch, run_frontend
from .... import _lib as L
from ....dsp import hanning, mel_filters


def log_mel_spectrogram(audio, sample_rate: int = 24_000, n_mels: int = 100, n_fft: int = 1024,
                        hop_length: int = 256, padding: int = 0):
    ing, _ = as_batch(audio)
    fb = mel_filters(sample_rate=sample_rate, n_fft=n_fft, n_mels=n_mels, norm=None, mel_scale="htk")
    out = run_frontend(
        ing, hanning(n_fft), fb, length=ing.data.shape[1] + max(int(padding), 0), n_fft=n_fft, hop=hop_length,
        center=True, pad_mode="reflect", spec_kind=L.SPEC_MAGNITUDE, log_kind=L.LOG_LN, guard_kind=L.GUARD_MAX,
        guard_eps=1e-5)
    return emit(ing, out)  # (1, T, M) / (B, T, M)
